"""Golden outputs of the UNMODIFIED reference (numpy backend) on the edge-case inputs of tests/edge_cases.py, incl.
where it produces nan / inf.  Writes tests/golden/edges.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from ref_loader import load_reference  # noqa: E402

load_reference()
from numga.backend.numpy.context import NumpyContext  # noqa: E402
from edge_cases import EDGE_CASES, PGA3  # noqa: E402

out = {}
for name, sub, rows, fn in EDGE_CASES:
    for dtype in (np.float32, np.float64):
        ctx = NumpyContext(PGA3, dtype=dtype)
        x = rows().astype(dtype)
        with np.errstate(all='ignore'):
            res = fn(ctx.multivector(values=x, subspace=getattr(ctx.subspace, sub)()))
        tag = f'{name}/{np.dtype(dtype).name}'
        out[f'{tag}/in'] = x
        out[f'{tag}/out'] = np.ascontiguousarray(res.values)
        out[f'{tag}/out_blades'] = res.subspace.blades.astype(np.int64)
np.savez_compressed(os.path.join(HERE, 'edges.npz'), **out)
print('wrote', len(out), 'arrays')
