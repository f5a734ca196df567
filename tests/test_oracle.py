"""Pin the oracle (oracle/numga_oracle.py) to the reference: integer tables exact, float outputs
bit-exact against vectors produced by the reference's numpy backend (tests/golden/vectors.npz)."""
import numpy as np
import pytest

import golden_io
from cases import CASES
from oracle import numga_oracle as O

BUILDERS = {
    'geometric_product': lambda a, l, r: O.product(a, l, r),
    'wedge_product': lambda a, l, r: O.product(a, l, r, 'wedge'),
    'inner_product': lambda a, l, r: O.product(a, l, r, 'inner'),
    'scalar_product': lambda a, l, r: O.product(a, l, r, 'scalar_product'),
    'left_contraction_product': lambda a, l, r: O.product(a, l, r, 'left_contraction'),
    'right_contraction_product': lambda a, l, r: O.product(a, l, r, 'right_contraction'),
    'reverse_product': O.reverse_product, 'regressive_product': O.regressive, 'commutator_product': O.commutator,
    'bivector_product': O.bivector_product, 'sandwich': O.sandwich, 'reverse_sandwich': O.reverse_sandwich,
    'full_sandwich': O.full_sandwich, 'inertia': O.inertia, 'select': O.select, 'restrict': O.restrict,
    'reverse': O.reverse, 'involute': O.involute, 'conjugate': O.conjugate, 'right_hodge': O.right_hodge,
    'right_hodge_inverse': O.right_hodge_inverse, 'scalar_negation': O.scalar_negation,
    'pseudoscalar_negation': O.pseudoscalar_negation, 'squared': O.squared,
    'symmetric_reverse_product': lambda a, x: O.symmetric_product(a, x, 'reverse'),
    'symmetric_scalar_negation_product': lambda a, x: O.symmetric_product(a, x, 'scalar_negation'),
    'symmetric_conjugate_product': lambda a, x: O.symmetric_product(a, x, 'conjugate'),
}
ALGEBRAS = sorted({k.split('/')[0] for k in golden_io.tables()})


@pytest.mark.parametrize('tag', ALGEBRAS)
def test_oracle_tables_match_reference(tag):
    alg = O.Algebra(tuple(int(c) for c in tag))
    checked = 0
    for key, (axes, index, coeff) in golden_io.tables().items():
        t, name, subs = key.split('/')
        if t != tag or name not in BUILDERS:
            continue
        args = [alg.subspace(s) for s in subs.split(',')]
        op = BUILDERS[name](alg, *args)
        assert [list(a) for a in op.axes] == [a.tolist() for a in axes], key
        assert np.array_equal(op.kernel.astype(np.int64), golden_io.dense_of(axes, index, coeff)), key
        checked += 1
    assert checked > 500


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_oracle_values_bit_exact(case, dtype):
    name, pqr, specs, scale, fn = case
    ins, want, want_blades = golden_io.case_vectors(name, dtype)
    ctx = O.context(pqr, dtype)
    mvs = [ctx.multivector(sub, a) for (sub, _), a in zip(specs, ins)]
    with np.errstate(all='ignore'):
        got = fn(*mvs)
    assert list(got.blades) == want_blades.tolist()
    assert got.values.dtype == dtype
    assert np.array_equal(np.broadcast_to(got.values, want.shape), want, equal_nan=True), \
        f'{name}: max abs diff {np.nanmax(np.abs(got.values - want))}'


def test_known_integer_answers():
    """exact small-integer results every backend must reproduce."""
    ctx = O.context((3, 0, 1), np.float64)
    a = ctx.multivector('even_grade', np.arange(1, 9.0))
    b = ctx.multivector('even_grade', np.arange(8, 0, -1.0))
    p = ctx.multivector('antivector', [1, 2, 3, 4.0])
    v = ctx.multivector('vector', [1, 2, 3, 4.0])
    bv = ctx.multivector('bivector', np.arange(1, 7.0))
    assert (a * b).values.tolist() == [-44, 32, 12, 46, -72, 102, 36, 114]
    assert a.reverse_product(b).values.tolist() == [60, 0, 36, 18, 144, 0, 72, 16]
    assert a.squared().values.tolist() == [-28, 4, 6, 8, -54, 60, -18, 48]
    assert a.symmetric_reverse_product().values.tolist() == [30, -16]
    assert (a >> p).values.tolist() == [30, 92, 90, -20]
    assert (a << p).values.tolist() == [30, 184, 162, 128]
    assert (a >> v).values.tolist() == [30, -60, -90, -116]
    assert (a >> bv).values.tolist() == [50, 68, 74, 160, 90, 164]
    assert (v ^ ctx.multivector('vector', [4, 3, 2, 1.0])).values.tolist() == [-5, -10, -5, -15, -10, -5]
    assert (v | bv).values.tolist() == [-8, -8, 8, 32]
    assert (p & ctx.multivector('antivector', [4, 3, 2, 1.0])).values.tolist() == [-5, -10, -15, -5, -10, -5]
    assert bv.commutator(ctx.multivector('bivector', np.arange(6, 0, -1.0))).values.tolist() == [7, -14, 7, -56, 0, 28]
    assert bv.dual().values.tolist() == [6, -5, 4, 3, -2, 1]
