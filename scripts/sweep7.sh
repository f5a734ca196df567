#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
for pv in last first; do for io in 1 0; do for fv in 1 2; do
NGB200_PIVOT=$pv NGB200_IR_ORDER=$io NGB200_FORCE_VEC=$fv python scripts/profile_kernel.py cga_time 2>&1 | tail -1
done; done; done
NGB200_FORCE_VEC=1 python scripts/dev_probe.py quick 2>&1 | grep "cga"
NGB200_FORCE_VEC=1 NGB200_IR_ORDER=1 python scripts/dev_probe.py quick 2>&1 | grep "cga"
NGB200_FORCE_VEC=2 NGB200_IR_ORDER=1 python scripts/dev_probe.py quick 2>&1 | grep "cga"
NGB200_FORCE_VEC=2 NGB200_IR_ORDER=0 python scripts/dev_probe.py quick 2>&1 | grep "cga"
