"""Golden values for 2D conformal GA built the way numga/examples/conformal.py does (`Conformal` wrapper around
NumpyContext(Algebra('x+y+') * Algebra('p+n-'))) — circle through three points, circle-circle intersection, point-pair
splitting (examples/test_conformal.py:55-80).  Writes tests/golden/conformal.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

load_reference()
from numga.backend.numpy.context import NumpyContext  # noqa: E402
from numga.examples.conformal import Conformal  # noqa: E402

cga = Conformal(NumpyContext(Conformal.algebra('x+y+')))
out = {}


def put(name, mv):
    out[name] = np.array(mv.values)
    out[name + '/blades'] = mv.subspace.blades.astype(np.int64)


put('ni', cga.ni); put('no', cga.no); put('N', cga.N)
pts = np.array([(-1.5, -0.5), (1, -0.5), (0, 1.5)])
a, b, c = (cga.embed_point(p) for p in pts)
put('a', a); put('b', b); put('c', c)
grid = np.random.default_rng(0).standard_normal((50, 2))
put('grid', cga.embed_point(grid, r=0.1))
out['grid/in'] = grid
put('b_inner_ni', b.inner(cga.ni)); put('bb', b * b); put('pos_b', cga.point_position(b))
C = a ^ b ^ c
D = cga.embed_point((-1, -1), r=1).dual()
T = C & D
put('C', C); put('D', D); put('T', T)
put('split', cga.split_points(T))
put('C_dual_normalized', cga.normalize(C.dual()))
put('reflect', D >> cga.embed_point(grid, r=0.1))
np.savez_compressed(os.path.join(HERE, 'conformal.npz'), **out)
print('wrote', len(out), 'arrays')
