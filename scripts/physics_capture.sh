#!/bin/bash
# ncu --set full of the six kernels of one physics sub-step (fp32): a launch-list pass finds where the timed steps
# begin (the last 6 generated-kernel launches of the run are one whole sub-step), then the full capture skips there.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/profile_kernel.py physics_time > gpurun_out/plain_physics.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:ngb_(vec|scalar|scalar_u)' --csv --log-file gpurun_out/launches_physics.csv python scripts/profile_kernel.py physics_time > /dev/null 2>&1
n=$(grep -c '"ngb_' gpurun_out/launches_physics.csv)
skip=$((n - 6))
echo "generated-kernel launches: $n, capturing from $skip"
ncu --set full --clock-control none --import-source on -k 'regex:ngb_(vec|scalar|scalar_u)' -s $skip -c 6 -o gpurun_out/prof_physics -f python scripts/profile_kernel.py physics_time > gpurun_out/ncu_physics_full.log 2>&1
echo "ncu physics rc=$?"
