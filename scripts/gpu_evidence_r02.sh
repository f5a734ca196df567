#!/bin/bash
# Round-2 GPU evidence (run under gpurun, ONE GPU): ncu --set full captures of the dominant kernels + launch list of bench.py.
# Outputs land in gpurun_out/ (scratch); scripts/ncu_summary.py condenses them into profiles/.
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:ngb_vec -c 2 -o $O/r02_sandwich python scripts/profile_kernel.py sandwich 3 > $O/r02_ncu_sandwich.log 2>&1
$NCU -k regex:ngb_vec -c 2 -o $O/r02_gp python scripts/profile_kernel.py gp 3 > $O/r02_ncu_gp.log 2>&1
$NCU -k regex:ngb_vec -c 2 -o $O/r02_roundtrip python scripts/profile_kernel.py roundtrip 3 > $O/r02_ncu_roundtrip.log 2>&1
$NCU -k regex:ngb_vec -c 2 -o $O/r02_roundtrip64 python scripts/profile_kernel.py roundtrip64 3 > $O/r02_ncu_roundtrip64.log 2>&1
$NCU -k regex:ngb_vec -c 2 -o $O/r02_cga python scripts/profile_kernel.py cga 3 > $O/r02_ncu_cga.log 2>&1
NGB200_STAGE=1 NGB200_NO_CACHE=1 $NCU -k regex:ngb_stage -c 2 -o $O/r02_cga_stage python scripts/profile_kernel.py cga 3 > $O/r02_ncu_cga_stage.log 2>&1
$NCU -k regex:ngb_pipeline -c 2 -o $O/r02_chain_f32_u1 python scripts/profile_chain.py f32 22 1 > $O/r02_ncu_chain1.log 2>&1
$NCU -k regex:ngb_pipeline -c 2 -o $O/r02_chain_f32_u10 python scripts/profile_chain.py f32 22 10 > $O/r02_ncu_chain10.log 2>&1
$NCU -k regex:ngb_pipeline -c 2 -o $O/r02_chain_f64_u1 python scripts/profile_chain.py f64 22 1 > $O/r02_ncu_chain64.log 2>&1
# launch list of the default bench command (cold-cache, serialised times: shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_bench_for_launchlist.json 2> $O/r02_bench_for_launchlist.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_ncu_bench.log 2>&1
ls -la $O | tail -30
