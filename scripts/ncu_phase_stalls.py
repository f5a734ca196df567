"""Attribute the warp-stall samples of an ngb_pipeline capture (ncu --set full --import-source on) to its barrier-delimited phases.
Usage: ncu_phase_stalls.py capture.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
lines = raw.splitlines()
# first kernel only
rows = list(csv.reader(lines))
hdr_i = [i for i,r in enumerate(rows) if r and r[0]=='Address']
H = rows[hdr_i[0]]
end = hdr_i[1]-1 if len(hdr_i)>1 else len(rows)
body = rows[hdr_i[0]+1:end]
iS = H.index('# Samples'); iI = H.index('Instructions Executed'); iSrc=H.index('Source')
stall_cols = [i for i,h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
# segment by BAR.SYNC
seg=0; segs=[]
cur={'n':0,'samples':0,'inst':0,'stalls':{}, 'ldg':0,'lds':0,'sts':0,'stg':0}
for r in body:
    src=r[iSrc]
    cur['n']+=1; cur['samples']+=int(r[iS] or 0); cur['inst']+=int(r[iI] or 0)
    for c in stall_cols:
        v=int(r[c] or 0)
        if v: cur['stalls'][H[c]]=cur['stalls'].get(H[c],0)+v
    for k,m in (('ldg','LDG'),('lds','LDS'),('sts','STS'),('stg','STG')):
        if m in src: cur[k]+=1
    if 'BAR.SYNC' in src or 'EXIT' in src:
        segs.append(cur); cur={'n':0,'samples':0,'inst':0,'stalls':{}, 'ldg':0,'lds':0,'sts':0,'stg':0}
segs.append(cur)
tot=sum(s['samples'] for s in segs)
for i,s in enumerate(segs):
    top=sorted(s['stalls'].items(), key=lambda kv:-kv[1])[:4]
    print(f"seg {i}: sass {s['n']:5d} ldg {s['ldg']:3d} lds {s['lds']:3d} sts {s['sts']:3d} stg {s['stg']:3d} warp-inst {s['inst']:>10d} samples {s['samples']:6d} ({100*s['samples']/tot:4.1f}%)  {top}")
