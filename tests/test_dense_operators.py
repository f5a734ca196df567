"""Partially bound operators: `partial`, `sandwich_map`, `inertia_map`, kernel `sum`, `inverse` and
their application, against outputs of the UNMODIFIED reference (`tests/golden/make_golden_dense.py` ->
`tests/golden/dense.npz`).  On the B200 backend each of these is one generated kernel: `partial` computes only the
structurally non-zero kernel entries, `inverse` is an unrolled Gauss-Jordan with partial pivoting, application is a
masked mat-vec.  CPU tier: numpy interpretation of the traced IR; GPU tier: the real kernels."""
import contextlib
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dense.npz'))
TORCH = {'f32': torch.float32, 'f64': torch.float64}
TOL = {'f32': 1e-5, 'f64': 1e-12}
INV_TOL = {'f32': 1e-2, 'f64': 1e-9}          # inverse of a sum of 3 rank-deficient maps: condition numbers 9 ... 1e5,
                                              # i.e. up to cond * eps = 6e-3 of legitimate fp32 rounding difference


@contextlib.contextmanager
def _context(where, tag, mode='fast'):
    if where == 'gpu':
        from numga_b200 import B200Context
        yield B200Context('x+y+z+w0', dtype=TORCH[tag], device='cuda:0', mode=mode)
    else:
        from cpu_emulation import emulate_launches, emulated_context
        with emulate_launches():
            yield emulated_context('x+y+z+w0', TORCH[tag], mode=mode)


def _err(got, want):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


WHERE = [pytest.param('gpu', marks=pytest.mark.gpu), 'cpu']


@pytest.mark.parametrize('where', WHERE)
@pytest.mark.parametrize('mode', ['fast', 'faithful'])
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_partial_operators_vs_reference(tag, mode, where):
    with _context(where, tag, mode) as ga:
        S = ga.subspace
        Q, V, P, B = S.even_grade(), S.vector(), S.antivector(), S.bivector()
        q, q1 = ga.multivector(Q, G[f'{tag}/q']), ga.multivector(Q, G[f'{tag}/q1'])
        v, p, b = ga.multivector(V, G[f'{tag}/v']), ga.multivector(P, G[f'{tag}/p']), ga.multivector(B, G[f'{tag}/b'])
        tol = TOL[tag]
        for name, sub, x in (('vector', V, v), ('bivector', B, b)):
            m = q.sandwich_map(sub)
            assert m.shape == (5,) and m.kshape == (len(sub), len(sub)) and m.output is sub
            assert _err(m.kernel, G[f'{tag}/sandwich_map_{name}/kernel']) < tol
            assert _err(m(x).values, G[f'{tag}/sandwich_map_{name}/apply']) < tol
            m1 = q1.sandwich_map(sub)
            assert m1.shape == ()
            assert _err(m1.kernel, G[f'{tag}/sandwich_map1_{name}/kernel']) < tol
            assert _err(m1(x).values, G[f'{tag}/sandwich_map1_{name}/apply']) < tol
            # the same thing spelled through the operator API (backend/test/test_numpy.py:19-22)
            m2 = ga.operator.sandwich(Q, sub).partial({0: q, 2: q})
            assert _err(m2.kernel, G[f'{tag}/sandwich_map_{name}/kernel']) < tol
        op = ga.operator.product(Q, Q).partial({0: q})
        assert _err(op.kernel, G[f'{tag}/product_partial/kernel']) < tol
        assert _err(op(q).values, G[f'{tag}/product_partial/apply']) < tol
        # passing a bare subspace for an argument means "leave it unbound" (backend/numpy/operator.py:118-121)
        op2 = ga.operator.product(Q, Q)(q, Q)
        assert _err(op2.kernel, G[f'{tag}/product_partial/kernel']) < tol
        inertia = p.inertia_map()
        assert inertia.shape == (5, 3)
        assert _err(inertia.kernel, G[f'{tag}/inertia_map/kernel']) < tol
        total = inertia.sum(axis=-3)
        assert total.shape == (5,)
        assert _err(total.kernel, G[f'{tag}/inertia_sum/kernel']) < tol
        assert _err(total(b).values, G[f'{tag}/inertia_sum/apply']) < tol
        inv = total.inverse()
        assert inv.output.blades.astype(np.int64).tolist() == G[f'{tag}/inertia_inv/blades'].tolist()
        assert _err(inv.kernel, G[f'{tag}/inertia_inv/kernel']) < INV_TOL[tag]
        assert _err(inv(total(b)).values, G[f'{tag}/b']) < INV_TOL[tag]


@pytest.mark.parametrize('where', WHERE)
def test_inverse_kernel_pivots_and_flags_singular_matrices(where):
    """The generated Gauss-Jordan inverse: partial pivoting (a tiny leading entry must not destroy accuracy),
    agreement with LAPACK on random matrices, inf / nan (never an exception) for a singular matrix."""
    from numga_b200.backend import B200DenseOperator
    rng = np.random.default_rng(5)
    with _context(where, 'f64') as ga:
        S = ga.subspace
        K = rng.standard_normal((64, 6, 6))
        K[::2, 0, 0] = 1e-12                       # forces a row swap in the first column
        K[1, :, :] = K[1, :, ::-1].copy()
        op = B200DenseOperator(ga, (S.bivector(), S.antibivector()), ga.coerce_array(K.reshape(64, 36)))
        inv = op.inverse()
        assert inv.inputs[0] is S.antibivector() and inv.output is S.bivector()
        assert _err(inv.kernel, np.linalg.inv(K)) < 1e-9
        single = B200DenseOperator(ga, (S.vector(), S.vector()), ga.coerce_array(K[3, :4, :4].reshape(16)))
        assert _err(single.inverse().kernel, np.linalg.inv(K[3, :4, :4])) < 1e-11
        sing = np.ones((2, 4, 4))
        bad = B200DenseOperator(ga, (S.vector(), S.vector()), ga.coerce_array(sing.reshape(2, 16))).inverse()
        assert not np.isfinite(bad.kernel.detach().cpu().numpy()).all()


@pytest.mark.parametrize('where', WHERE)
def test_inertia_map_equals_direct_expression(where):
    """numga/operator/test/test_operator.py:118-135: composed ternary operator == its direct expression."""
    rng = np.random.RandomState(0)
    with _context(where, 'f64') as ga:
        p = ga.multivector(ga.subspace.antivector(), rng.normal(size=(10, 4))).normalized()
        b = ga.multivector(ga.subspace.bivector(), rng.normal(size=(10, 6)))
        direct = p.regressive(p.commutator(b))
        assert _err(p.inertia_map()(b).values, direct.values.detach().cpu().numpy()) < 1e-12
