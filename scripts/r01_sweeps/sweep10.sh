#!/bin/bash
# CGA full x full and friends: contiguous one-item kernel (batch stride compiled in), run-time blockDim.x in the index
# arithmetic (NGB200_BDIM), short (10 steps) and sustained (100 steps) timing
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
for bd in 0 1; do
for st in 10 100; do
NGB200_BDIM=$bd STEPS=$st timeout 300 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_BDIM=$bd STEPS=$st NGB200_FORCE_VEC=2 timeout 300 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_BDIM=$bd STEPS=$st timeout 300 python scripts/profile_kernel.py roundtrip_time 2>&1 | tail -1
done
NGB200_BDIM=$bd timeout 300 python scripts/profile_kernel.py physics_time 2>&1 | tail -1
NGB200_BDIM=$bd timeout 300 python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1
done
