"""Measured error of the fast kernels on the edge-case fixtures (tests/golden/edges.npz) vs the reference's outputs,
per case and dtype (rows where the reference is finite; fp32 rows with subnormal squared norms excluded, as the test does).
Prints one line per case; run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from edge_cases import EDGE_CASES, PGA3
from numga_b200 import B200Context
Z = np.load(os.path.join(ROOT, 'tests', 'golden', 'edges.npz'))
for name, sub, _, fn in EDGE_CASES:
    row = {}
    for dtype, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
        tag = f'{name}/{np.dtype(dtype).name}'
        x, want = Z[f'{tag}/in'], Z[f'{tag}/out']
        ctx = B200Context(PGA3, dtype=tdt, device='cuda:0')
        fast = ctx.jit(fn)(ctx.multivector(values=x, subspace=getattr(ctx.subspace, sub)())).numpy()
        ok = np.isfinite(want).all(axis=-1)
        if dtype == np.float32:
            ok &= np.abs(x).max(axis=-1) > 1e-18
        scale = np.maximum(np.abs(want[ok]).max(axis=-1, keepdims=True), 1e-30)
        row[np.dtype(dtype).name] = float((np.abs(fast[ok] - want[ok]) / scale).max()) if ok.any() else 0.0
    print('edge', name, row, flush=True)
