#!/bin/bash
# first GPU probe: correctness + load/store/grid sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python scripts/dev_probe.py check
for ld in 0 1 2 3; do for st in 0 1 2; do NGB200_LDMODE=$ld NGB200_STMODE=$st python scripts/dev_probe.py quick | grep -E 'RESULT|bench'; done; done
for w in 2 4 8; do NGB200_WAVES=$w python scripts/dev_probe.py quick | grep -E 'RESULT|bench'; done
for t in 128 512; do NGB200_TPB=$t python scripts/dev_probe.py quick | grep -E 'RESULT|bench'; done
} > gpurun_out/sweep1.log 2>&1
tail -5 gpurun_out/sweep1.log
