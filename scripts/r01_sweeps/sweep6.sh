#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
for io in 1 0; do for fv in 1 2; do for mb in 1 2 3; do
NGB200_IR_ORDER=$io NGB200_FORCE_VEC=$fv NGB200_MINBLOCKS=$mb python scripts/profile_kernel.py cga_time 2>&1 | tail -1
done; done; done
NGB200_FORCE_VEC=1 NGB200_MINBLOCKS=4 NGB200_TPB=128 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_FORCE_VEC=1 NGB200_MINBLOCKS=5 NGB200_TPB=128 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
python scripts/profile_kernel.py physics_time 2>&1 | tail -1
python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1
python scripts/profile_kernel.py roundtrip_time 2>&1 | tail -1
