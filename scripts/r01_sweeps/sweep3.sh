#!/bin/bash
cd "$(dirname "$0")/.."
export NGB200_CACHE_DIR=/tmp/kc
mkdir -p gpurun_out
for mb in 4 5 6 8; do NGB200_MINBLOCKS=$mb python scripts/profile_kernel.py roundtrip_time 2>&1 | tail -1; done
for mb in 3 4 6; do NGB200_TPB=128 NGB200_MINBLOCKS=$mb python scripts/profile_kernel.py physics_time 2>&1 | tail -1; done
for mb in 1 2 3; do NGB200_MINBLOCKS=$mb python scripts/profile_kernel.py physics_time f64 2>&1 | tail -1; done
# per-kernel times of one physics sub-step + full metrics of the constraint kernels
NGB200_MINBLOCKS=2 python scripts/profile_kernel.py physics_time > /dev/null 2>&1 && \
NGB200_MINBLOCKS=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,launch__registers_per_thread,smsp__inst_executed.sum --clock-control none -k regex:ngb_ -s 120 -c 24 --csv --log-file gpurun_out/physics_kernels.csv python scripts/profile_kernel.py physics_time > gpurun_out/ncu_physics.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/physics_kernels.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
by = {}
for r in rows[hdr + 1:]:
    d = dict(zip(H, r))
    by.setdefault(d['ID'], {'name': d['Kernel Name'], 'grid': d['Grid Size']})[d['Metric Name']] = d['Metric Value']
for k, v in list(by.items())[-12:]:
    print(k, v)
PY
