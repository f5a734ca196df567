"""compute-sanitizer target: the chain-resident kernel on a batch with a PARTIAL last tile, fp32 and fp64, several tile
sizes, with and without exploit_zeros; plus a gather / batch-sum / fast-path call each.  Run as
    compute-sanitizer --tool memcheck python scripts/sanitize_chain.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numga_b200 import B200Context
from numga_b200.examples.physics import ChainStep, FusedStep, replicate_chains, setup_bodies

for dt in (torch.float32, torch.float64):
    ctx = B200Context('x+y+z+w0', dtype=dt)
    bodies, sets = setup_bodies(ctx, n_bodies=12)
    big, bsets = replicate_chains(bodies, sets, 45)                    # 45 chains: 32 + 13 (partial tile at 384 bodies/tile)
    ref = FusedStep(big, bsets).integrate(FusedStep(big, bsets).integrate(big, 1 / 200), 1 / 200)
    for bpt, zeros in ((None, False), (84, False), (None, True)):
        got = ChainStep(big, bsets, unroll=2, bodies_per_tile=bpt, exploit_zeros=zeros).integrate(big, 1 / 200)
        err = float((got.motor.values - ref.motor.values).abs().max())
        print('chain', dt, bpt, zeros, 'max abs diff vs six kernels', err, flush=True)
        assert err < 1e-5
ctx = B200Context((3, 0, 1))
p = ctx.multivector.antivector(torch.randn(1001, 4, device='cuda'))
idx = torch.randint(0, 1001, (777,), device='cuda')
assert torch.equal(p[idx].values, p.values[idx])
q = ctx.multivector.antivector(torch.randn(7, 333, 4, device='cuda'))
assert torch.allclose(q.sum(0).values, q.values.sum(0), atol=1e-5) and torch.allclose(p.sum(0).values, p.values.sum(0), atol=1e-3)
R = ctx.multivector.motor(torch.randn(8, device='cuda'))
for _ in range(3):
    out = R >> p
torch.cuda.synchronize()
print('sanitize ok')
