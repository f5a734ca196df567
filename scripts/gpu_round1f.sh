#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export NGB200_CACHE_DIR=/tmp/kc
python scripts/e2e_sweep.py 2>&1 | tail -16
python scripts/profile_kernel.py cga_time 2>&1 | tail -1
NGB200_FORCE_VEC=1 python scripts/profile_kernel.py cga_time 2>&1 | tail -1
python scripts/profile_kernel.py roundtrip_time 2>&1 | tail -1
python scripts/profile_kernel.py physics_time 2>&1 | tail -1
NGB200_FORCE_VEC=1 python scripts/profile_kernel.py physics_time 2>&1 | tail -1
python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1
for which in cga roundtrip; do
python scripts/profile_kernel.py $which 3 > gpurun_out/plain_$which.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:ngb_(vec|scalar)' -s 1 -c 1 -o gpurun_out/prof_$which -f python scripts/profile_kernel.py $which 3 > gpurun_out/ncu_$which.log 2>&1
echo "ncu $which rc=$?"
done
python scripts/profile_kernel.py physics_time > gpurun_out/plain_physics.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:ngb_(vec|scalar)' -s 140 -c 6 -o gpurun_out/prof_physics -f python scripts/profile_kernel.py physics_time > gpurun_out/ncu_physics_full.log 2>&1
echo "ncu physics rc=$?"; tail -3 gpurun_out/ncu_physics_full.log
ls -la gpurun_out/*.ncu-rep
