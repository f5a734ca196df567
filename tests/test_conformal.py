"""2D conformal GA the way `numga/examples/conformal.py` builds it (algebra product 'x+y+' * 'p+n-', null basis ni / no,
point embedding, circles as wedges of three points, circle-circle intersection, point-pair splitting), against values
produced by the unmodified reference (`tests/golden/make_golden_conformal.py`, from examples/test_conformal.py:55-80).
Exercises API surface the hot-path cases do not: algebra products, basis-vector multivectors, subspace set algebra,
mixed scalar / array arithmetic, mixed-grade combine."""
import contextlib
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'conformal.npz'))


@contextlib.contextmanager
def _context(where):
    from numga_b200.algebra import Algebra
    algebra = Algebra('x+y+') * Algebra('p+n-')
    if where == 'gpu':
        from numga_b200 import B200Context
        yield B200Context(algebra, dtype=torch.float64, device='cuda:0')
    else:
        from cpu_emulation import emulate_launches, emulated_context
        with emulate_launches():
            yield emulated_context(algebra, torch.float64, mode='fast')


class Conformal:
    """Same formulas as numga/examples/conformal.py:145-224."""

    def __init__(self, context):
        self.context = context
        *X, p, n = context.multivector.basis()
        self.ni, self.no = p + n, (n - p) / 2
        self.N = self.ni ^ self.no
        S = context.subspace
        self.euclidian = S.vector().difference(p.subspace.union(n.subspace))

    def embed_point(self, P, r=0):
        p = self.context.multivector(self.euclidian, np.asarray(P, dtype=np.float64))
        return (self.ni * (p.norm_squared() - r * r) / 2 + self.no) + p

    def point_position(self, p):
        return (p ^ self.N) * self.N

    def split_points(self, T):
        b = (T.norm_squared() * -1).sqrt()
        F = T / b
        S = self.context.multivector.scalar([[1], [-1]])
        P = (F * S + 1) / 2
        return P.inner(T.inner(self.ni))

    def normalize(self, o):
        return o / (-o.reverse().inner(self.ni))


def _check(name, mv, atol=1e-12):
    want = G[name]
    assert mv.subspace.blades.astype(np.int64).tolist() == G[name + '/blades'].tolist(), name
    got = mv.values.detach().cpu().numpy()
    assert np.allclose(got, want, atol=atol * max(1.0, np.abs(want).max())), (name, np.abs(got - want).max())


@pytest.mark.parametrize('where', [pytest.param('gpu', marks=pytest.mark.gpu), 'cpu'])
def test_conformal_circles(where):
    with _context(where) as ctx:
        cga = Conformal(ctx)
        _check('ni', cga.ni); _check('no', cga.no); _check('N', cga.N)
        assert np.allclose(cga.ni.norm().values.cpu().numpy(), 0) and np.allclose((cga.N * cga.N).values.cpu().numpy(), 1)
        a, b, c = (cga.embed_point(p) for p in [(-1.5, -0.5), (1, -0.5), (0, 1.5)])
        _check('a', a); _check('b', b); _check('c', c)
        grid = cga.embed_point(G['grid/in'], r=0.1)
        _check('grid', grid)
        _check('b_inner_ni', b.inner(cga.ni)); _check('bb', b * b); _check('pos_b', cga.point_position(b))
        C = a ^ b ^ c
        D = cga.embed_point((-1, -1), r=1).dual()
        T = C & D
        _check('C', C); _check('D', D); _check('T', T)
        _check('split', cga.split_points(T), atol=1e-10)
        _check('C_dual_normalized', cga.normalize(C.dual()))
        _check('reflect', D >> grid)
