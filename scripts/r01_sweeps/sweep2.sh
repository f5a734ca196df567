#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
for mb in 1 2 3; do NGB200_MINBLOCKS=$mb python scripts/profile_kernel.py physics_time 2>&1 | tail -1; done
for mb in 1 2; do NGB200_MINBLOCKS=$mb python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1; done
for pm in 0 1; do NGB200_PRECISE_MATH=$pm python scripts/profile_kernel.py roundtrip_time 2>&1 | tail -1; done
for mb in 2 3; do NGB200_MINBLOCKS=$mb python scripts/profile_kernel.py roundtrip_time 2>&1 | tail -1; done
python scripts/profile_kernel.py cga_pitch 2>&1 | tail -8
for mb in 1 2; do for fv in 1 2; do NGB200_MINBLOCKS=$mb NGB200_PAIR=0 NGB200_FORCE_VEC=$fv python scripts/profile_kernel.py cga_time 2>&1 | tail -1; done; done
