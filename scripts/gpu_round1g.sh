#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python -m pytest tests/test_reference_identities.py -m gpu -q -x ) > gpurun_out/pytest_identities.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_identities.log
python scripts/profile_kernel.py physics_time > gpurun_out/plain_physics.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:ngb_scalar' -s 30 -c 5 -o gpurun_out/prof_physics -f python scripts/profile_kernel.py physics_time > gpurun_out/ncu_physics_full.log 2>&1
echo "ncu physics rc=$?"; tail -3 gpurun_out/ncu_physics_full.log
ls -la gpurun_out/*.ncu-rep
