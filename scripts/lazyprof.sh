cd /root/repo
MODES=lazy python scripts/profile_kernel.py physics_generic_time > /dev/null 2>&1
MODES=lazy ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/lazy_launches.csv python scripts/profile_kernel.py physics_generic_time > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/lazy_launches.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
seq = [(dict(zip(H, r))['Kernel Name'][:300], float(dict(zip(H, r))['Metric Value'])) for r in rows[hdr + 1:]]
last = seq[-60:]   # roughly the last sub-steps
agg = collections.Counter(); cnt = collections.Counter()
for n, t in last: agg[n] += t; cnt[n] += 1
for n, t in agg.most_common(8): print(f'{t/1e3:10.1f} us total {cnt[n]:4d} x  {n[:46] + ' ' + n[60:150]}')
PY
