#!/bin/bash
# physics sub-step: grid waves and CTA size (final kernels)
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
for w in 8 2 4 16; do
NGB200_WAVES=$w timeout 300 python scripts/profile_kernel.py physics_time 2>&1 | tail -1
done
NGB200_TPB=128 timeout 300 python scripts/profile_kernel.py physics_time 2>&1 | tail -1
NGB200_WAVES=2 timeout 300 python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1
NGB200_WAVES=16 timeout 300 python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1
NGB200_WAVES=2 timeout 300 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_WAVES=16 timeout 300 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_WAVES=1 timeout 300 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
