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


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_oracle_sparse_operator_variant(dtype):
    """`O.OTYPE = 'sparse'` restates NumpySparseOperator (backend/numpy/operator.py:143-160), the second CPU baseline
    of bench.py.  Checked here against the golden vectors of the reference: within rounding of the einsum operator on
    every case (the two reference operators sum the same terms in different association), and bit-identical to
    the reference's own NumpySparseOperator when generated in the build container (see tests/golden/make_golden.py
    docstring: `python /tmp/mk_sparse.py`-style spot check recorded in DESIGN.md)."""
    O.OTYPE = 'sparse'
    try:
        exact = 0
        for name, pqr, specs, scale, fn in CASES:
            ins, want, want_blades = golden_io.case_vectors(name, dtype)
            ctx = O.context(pqr, dtype)
            mvs = [ctx.multivector(sub, a) for (sub, _), a in zip(specs, ins)]
            with np.errstate(all='ignore'):
                got = fn(*mvs)
            assert list(got.blades) == want_blades.tolist()
            g = np.broadcast_to(got.values, want.shape)
            ok = np.isfinite(want) & np.isfinite(g)
            scale_ = np.abs(want[ok]).max() if ok.any() else 1.0
            tol = 5e-3 if dtype == np.float32 else 1e-9
            assert np.abs(g[ok] - want[ok]).max() <= tol * scale_, name
            exact += bool(np.array_equal(g, want, equal_nan=True))
        assert exact >= len(CASES) // 3
    finally:
        O.OTYPE = 'einsum'


@pytest.mark.parametrize('tag,dtype', [('f64', np.float64), ('f32', np.float32)])
def test_oracle_physics_bit_exact(tag, dtype):
    """oracle/physics_oracle.py (the CPU restatement of numga/examples/physics/core.py used as config 5's CPU baseline
    and parity checker) against the stage-by-stage states of the unmodified reference: bit-exact, both dtypes."""
    import os
    from oracle import physics_oracle as P
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'physics.npz'))
    arrays = {k[len(tag) + 1:]: G[k] for k in G.files if k.startswith(tag + '/')}
    bodies, sets = P.from_arrays(arrays, dtype, rate=arrays['start/rate'])
    dt = 1 / 200
    new = bodies.pre_integrate(dt)
    assert np.array_equal(new.motor.values, arrays['pre/motor']) and np.array_equal(new.rate.values, arrays['pre/rate'])
    for k, c in enumerate(sets):
        new = c.apply(new, dt)
        assert np.array_equal(new.motor.values, arrays[f'apply{k}/motor'])
    new = new.post_integrate(bodies, dt)
    assert np.array_equal(new.motor.values, arrays['post/motor']) and np.array_equal(new.rate.values, arrays['post/rate'])
    for k, c in enumerate(sets):
        new = c.v_apply(new, dt)
        assert np.array_equal(new.rate.values, arrays[f'vapply{k}/rate'])
    b = bodies
    for step in range(1, 6):
        b = b.integrate(dt, sets)
        if step in (1, 2, 5):
            assert np.array_equal(b.motor.values, arrays[f'step{step}/motor'])
            assert np.array_equal(b.rate.values, arrays[f'step{step}/rate'])
    # replicated chains evolve alike (block-diagonal constraints; numpy's einsum may pick another inner loop for
    # another batch size, so this is equality to rounding, not bitwise)
    big, bsets = P.from_arrays(arrays, dtype, n_chains=3, rate=arrays['start/rate'])
    many = big.integrate(dt, bsets)
    assert np.allclose(many.motor.values.reshape(3, 12, 8), arrays['step1/motor'][None], rtol=0, atol=1e-5 if tag == 'f32' else 1e-13)
