#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
# correctness first (small, with a watchdog), then timings
NGB200_STAGE=1 timeout 120 python - <<'PY'
import numpy as np, torch
from numga_b200 import B200Context
ctx = B200Context((4, 1, 0)); F = ctx.subspace.full()
n = 128 * 37 + 5
a = ctx.multivector(values=torch.randn(n, 32, device='cuda'), subspace=F); b = ctx.multivector(values=torch.randn(n, 32, device='cuda'), subspace=F)
got = (a * b).values
import os
os.environ['NGB200_STAGE'] = '0'
ctx2 = B200Context((4, 1, 0))
a2 = ctx2.multivector(values=a.values, subspace=ctx2.subspace.full()); b2 = ctx2.multivector(values=b.values, subspace=ctx2.subspace.full())
want = (a2 * b2).values
torch.cuda.synchronize()
print('staged == plain:', torch.equal(got, want), float((got - want).abs().max()))
PY
echo "check rc=$?"
for st in 2 3; do for tp in 128 256; do
NGB200_STAGE=1 NGB200_STAGES=$st NGB200_STAGE_TPB=$tp timeout 120 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
done; done
NGB200_STAGE=1 NGB200_STAGES=4 NGB200_STAGE_TPB=64 timeout 120 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_STAGE=1 timeout 120 python scripts/profile_kernel.py physics_time 2>&1 | tail -1
NGB200_STAGE=1 timeout 120 python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1
