#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
mkdir -p gpurun_out
for io in 1 0; do
export NGB200_IR_ORDER=$io
python scripts/profile_kernel.py physics_time f64 | tail -1
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__block_size --clock-control none -c 3000 --csv --log-file gpurun_out/physics_launches64_$io.csv python scripts/profile_kernel.py physics_time f64 > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/physics_launches64_$io.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
by = {}
for r in rows[hdr + 1:]:
    d = dict(zip(H, r)); by.setdefault(d['ID'], {'name': d['Kernel Name'][:20]})[d['Metric Name']] = d['Metric Value']
seq = [v for v in by.values() if v['name'].startswith('ngb_')][-6:]
for v in seq: print('  ', v)
PY
done
