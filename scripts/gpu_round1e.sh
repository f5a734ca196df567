#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python scripts/e2e_sweep.py 2>&1 | tail -16
# full ncu captures (plain run first)
for which in cga roundtrip gp; do
python scripts/profile_kernel.py $which 3 > gpurun_out/plain_$which.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ngb_ -s 1 -c 1 -o gpurun_out/prof_$which -f python scripts/profile_kernel.py $which 3 > gpurun_out/ncu_$which.log 2>&1
echo "ncu $which rc=$?"
done
python scripts/profile_kernel.py physics_time > gpurun_out/plain_physics.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ngb_ -s 160 -c 6 -o gpurun_out/prof_physics -f python scripts/profile_kernel.py physics_time > gpurun_out/ncu_physics_full.log 2>&1
echo "ncu physics rc=$?"
ls -la gpurun_out/*.ncu-rep
