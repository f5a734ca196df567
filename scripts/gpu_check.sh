#!/bin/bash
# GPU box: parity tests, then the bench with all configurations; summaries into gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --configs > gpurun_out/bench_configs.json 2> gpurun_out/bench_configs.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_configs.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'clocks', d['clocks'])
print('e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline'))
for c in d['configs']: print(f"{c['config'][:90]:90s} {c['ms']:9.3f} ms {c['items_per_s']/1e9:8.2f} G/s {c['achieved_gbs']:7.0f} GB/s frac {c['frac_of_hbm_peak']:.3f} {c.get('tflops', 0):.1f} TF")
PY
tail -3 gpurun_out/bench_configs.err
