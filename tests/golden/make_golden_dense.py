"""Golden data for partially bound (dense-kernel) operators from the UNMODIFIED reference: `partial`, `sandwich_map`,
`inertia_map`, kernel `sum`, `inverse` and their application (numga/backend/numpy/operator.py:56-132,
multivector/multivector.py:362-395).  Writes tests/golden/dense.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

load_reference()
from numga.backend.numpy.context import NumpyContext  # noqa: E402

out = {}
for tag, dtype in (('f32', np.float32), ('f64', np.float64)):
    rng = np.random.default_rng(11)
    ga = NumpyContext('x+y+z+w0', dtype=dtype)
    S = ga.subspace
    Q, V, P, B = S.even_grade(), S.vector(), S.antivector(), S.bivector()
    q = ga.multivector(Q, rng.standard_normal((5, 8)).astype(dtype))
    q1 = ga.multivector(Q, rng.standard_normal(8).astype(dtype))
    v = ga.multivector(V, rng.standard_normal((5, 4)).astype(dtype))
    p = ga.multivector(P, rng.standard_normal((5, 3, 4)).astype(dtype))          # 5 bodies x 3 points
    b = ga.multivector(B, rng.standard_normal((5, 6)).astype(dtype))
    for name, arr in (('q', q), ('q1', q1), ('v', v), ('p', p), ('b', b)):
        out[f'{tag}/{name}'] = np.array(arr.values)
    for sub_name, sub, x in (('vector', V, v), ('bivector', B, b)):
        m = q.sandwich_map(sub)
        out[f'{tag}/sandwich_map_{sub_name}/kernel'] = np.array(m.kernel)
        out[f'{tag}/sandwich_map_{sub_name}/apply'] = np.array(m(x).values)
        m1 = q1.sandwich_map(sub)                                                   # one motor, batch of vectors
        out[f'{tag}/sandwich_map1_{sub_name}/kernel'] = np.array(m1.kernel)
        out[f'{tag}/sandwich_map1_{sub_name}/apply'] = np.array(m1(x).values)
    op = ga.operator.product(Q, Q).partial({0: q})                                   # left multiplication matrix
    out[f'{tag}/product_partial/kernel'] = np.array(op.kernel)
    out[f'{tag}/product_partial/apply'] = np.array(op(q).values)
    I = p.inertia_map()
    out[f'{tag}/inertia_map/kernel'] = np.array(I.kernel)
    Isum = I.sum(axis=-3)
    out[f'{tag}/inertia_sum/kernel'] = np.array(Isum.kernel)
    out[f'{tag}/inertia_sum/apply'] = np.array(Isum(b).values)
    Iinv = Isum.inverse()
    out[f'{tag}/inertia_inv/kernel'] = np.array(Iinv.kernel)
    out[f'{tag}/inertia_inv/apply'] = np.array(Iinv(Isum(b)).values)
    out[f'{tag}/inertia_inv/blades'] = Iinv.operator.output.blades.astype(np.int64)
np.savez_compressed(os.path.join(HERE, 'dense.npz'), **out)
print('wrote', len(out), 'arrays')
