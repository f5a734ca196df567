#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --configs > gpurun_out/bench_configs.json 2> gpurun_out/bench_configs.err; echo "bench rc=$?"
cat gpurun_out/bench_configs.json | head -c 6000; tail -3 gpurun_out/bench_configs.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; echo "ref rc=$?"; cat gpurun_out/bench_reference.json | head -c 1500
nproc; free -g | head -2
# ncu: launch list of the bench command (plain run first), then one full capture of the top kernel
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu list rc=$?"
python scripts/profile_kernel.py sandwich 4 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ngb_vec -s 1 -c 2 -o gpurun_out/prof_sandwich -f python scripts/profile_kernel.py sandwich 4 > gpurun_out/ncu_prof.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_prof.log
