"""Measure, per case of the catalogue (tests/cases.py) and for the physics sub-steps, the error of the benchmarked
(fast) GPU kernels and of the reference's own arithmetic at the same precision, both against the fp64 oracle:

    e_gpu_fast   = max_item ||gpu_fast(fp32) - oracle(fp64)||inf / ||oracle(fp64)||inf
    e_ref_fp32   = the same for the oracle port run in fp32 (= the reference's numpy backend, bit for bit)
    e_vs_ref     = gpu_fast(fp32) against the fp32 oracle (the quantity the parity tests gate)
    (and the fp64 analogues against an fp64 oracle; there e_ref is 0 by construction)

Writes profiles/r02_tolerances.json + a markdown table; the gates in tests/ are derived from it (no constant looser than
2x the reference's own error at that precision, with a floor of the north-star 1e-5 / 1e-12).  Run on the GPU box."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch

from cases import CASES, case_inputs
from numga_b200 import B200Context
from oracle import numga_oracle as O

N = 1 << 14


def rel(got, want):
    got, want = np.asarray(got, np.float64), np.broadcast_to(np.asarray(want, np.float64), np.shape(got))
    scale = np.maximum(np.abs(want).max(axis=-1), 1e-300)
    with np.errstate(all='ignore'):
        e = np.abs(got - want).max(axis=-1) / scale
    e = e[np.isfinite(e)]
    return (float(e.max()), float(np.quantile(e, 0.999))) if e.size else (0.0, 0.0)


rows = []
for case in CASES:
    name, pqr, specs, scale, fn = case
    widths = [len(O.Algebra(pqr).subspace(s)) for s, _ in specs]
    arrays64 = case_inputs(case, np.float64, batch=N, seed=12345)(widths)
    with np.errstate(all='ignore'):
        want64 = fn(*[O.context(pqr, np.float64).multivector(s, a) for (s, _), a in zip(specs, arrays64)]).values
        ref32 = fn(*[O.context(pqr, np.float32).multivector(s, a.astype(np.float32)) for (s, _), a in zip(specs, arrays64)]).values
    row = {'case': name}
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        ctx = B200Context(pqr, dtype=dt, device='cuda:0')
        npd = np.float32 if tag == 'f32' else np.float64
        mvs = [ctx.multivector(values=a.astype(npd), subspace=getattr(ctx.subspace, s)()) for (s, _), a in zip(specs, arrays64)]
        got = ctx.jit(fn)(*mvs).numpy()
        row[f'e_gpu_fast_{tag}'], row[f'e_gpu_fast_{tag}_q999'] = rel(got, want64)
        if tag == 'f32':
            row['e_ref_f32'], row['e_ref_f32_q999'] = rel(ref32, want64)
            row['e_vs_ref_f32'], _ = rel(got, ref32)
        else:
            row['e_vs_ref_f64'], _ = rel(got, want64)
    rows.append(row)
    print(row, flush=True)

# physics: 1, 2, 5 sub-steps of the chain-resident kernel and the six-kernel arrangement on 2^14 bodies
from bench import physics_arrays, physics_state, PHYSICS_DT
from numga_b200.examples.physics import ChainStep, FusedStep
from oracle import physics_oracle as P
pc = N // 12
hrate = 0.3 * np.random.default_rng(105).standard_normal((12 * pc, 6))
a64 = physics_arrays('f64')
ob64, os64 = P.from_arrays(a64, np.float64, n_chains=pc, rate=np.tile(a64['init/rate'], (pc, 1)) + hrate)
a32 = physics_arrays('f32')
ob32, os32 = P.from_arrays(a32, np.float32, n_chains=pc, rate=(np.tile(a32['init/rate'], (pc, 1)) + hrate).astype(np.float32))
phys = []
states = {}
for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
    ctx = B200Context('x+y+z+w0', dtype=dt, device='cuda:0')
    small, ssets = physics_state(ctx, pc, 0, host_rate=hrate.astype(np.float32 if tag == 'f32' else np.float64))
    states[tag] = (small, ssets, ChainStep(small, ssets, unroll=1), FusedStep(small, ssets))
w64, w32 = ob64, ob32
g = {tag: {'chain': states[tag][0], 'fused': states[tag][0]} for tag in states}
for step in range(1, 6):
    with np.errstate(all='ignore'):
        w64, w32 = w64.integrate(PHYSICS_DT, os64), w32.integrate(PHYSICS_DT, os32)
    for tag in states:
        g[tag]['chain'] = states[tag][2].integrate(g[tag]['chain'], PHYSICS_DT)
        g[tag]['fused'] = states[tag][3].integrate(g[tag]['fused'], PHYSICS_DT)
    if step in (1, 2, 5):
        row = {'case': f'physics sub-step x{step}'}
        for field in ('motor', 'rate'):
            row[f'e_ref_f32_{field}'], _ = rel(getattr(w32, field).values, getattr(w64, field).values)
            for tag in states:
                for k in ('chain', 'fused'):
                    row[f'e_{k}_{tag}_{field}'], _ = rel(getattr(g[tag][k], field).numpy(), getattr(w64, field).values)
            row[f'e_chain_vs_ref_f32_{field}'], _ = rel(getattr(g['f32']['chain'], field).numpy(), getattr(w32, field).values)
        phys.append(row)
        print(row, flush=True)

out = {'n_items': N, 'cases': rows, 'physics': phys, 'how': __doc__}
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r02_tolerances.json'), 'w'), indent=1)
