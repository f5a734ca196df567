#!/bin/bash
# gathered physics kernels: batch stride of the scalar kernel compiled in (experiment knob NGB200_UNIT_BS) x blockDim form
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
for ub in 0 1; do for bd in 0 1; do
NGB200_UNIT_BS=$ub NGB200_BDIM=$bd timeout 300 python scripts/profile_kernel.py physics_time 2>&1 | tail -1
NGB200_UNIT_BS=$ub NGB200_BDIM=$bd timeout 300 python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1
done; done
