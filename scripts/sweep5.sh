#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
for pf in 0 1; do NGB200_PREFETCH=$pf python scripts/profile_kernel.py cga_time 2>&1 | tail -1; done
NGB200_PREFETCH=1 NGB200_WAVES=1 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_PREFETCH=0 NGB200_WAVES=1 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_PREFETCH=1 NGB200_FORCE_VEC=1 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
python scripts/profile_kernel.py roundtrip_time 2>&1 | tail -1
for pf in 0 1; do NGB200_PREFETCH=$pf python scripts/profile_kernel.py physics_time 2>&1 | tail -1; done
