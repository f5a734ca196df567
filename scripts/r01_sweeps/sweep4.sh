#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
mkdir -p gpurun_out
NGB200_MINBLOCKS=2 python scripts/profile_kernel.py physics_time > /dev/null 2>&1
NGB200_MINBLOCKS=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/physics_launches.csv python scripts/profile_kernel.py physics_time > gpurun_out/ncu_physics.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/physics_launches.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
seq = [(dict(zip(H, r))['Kernel Name'][:60], float(dict(zip(H, r))['Metric Value'])) for r in rows[hdr + 1:]]
print(len(seq), 'launches')
for name, ns in seq[-40:]:
    print(f'{ns/1e3:10.1f} us  {name}')
PY
