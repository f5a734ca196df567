#!/bin/bash
# Round evidence: bench (all configs), reference arm, ncu launch list of the bench command, full ncu captures of the
# dominant kernels.  Every ncu pass follows a plain run of the same command that exited 0.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --configs > gpurun_out/bench_configs.json 2> gpurun_out/bench_configs.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu list rc=$?"
for which in sandwich cga roundtrip gp; do
python scripts/profile_kernel.py $which 3 > gpurun_out/plain_$which.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:ngb_(vec|scalar)' -s 1 -c 1 -o gpurun_out/prof_$which -f python scripts/profile_kernel.py $which 3 > gpurun_out/ncu_$which.log 2>&1
echo "ncu $which rc=$?"
done
python scripts/profile_kernel.py physics_time > gpurun_out/plain_physics.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:ngb_(vec|scalar|scalar_u)' -s 30 -c 6 -o gpurun_out/prof_physics -f python scripts/profile_kernel.py physics_time > gpurun_out/ncu_physics_full.log 2>&1
echo "ncu physics rc=$?"
ls -la gpurun_out/*.ncu-rep
