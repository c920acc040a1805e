#!/bin/bash
set -u
mkdir -p gpurun_out
PTB_MESH_PIPELINE=1 PTB_MP_CTAS_PER_SM=16 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02f_pipe_launches.csv python tools/profile_kernel.py C4_1M 16 1 > gpurun_out/r02f_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02f_pipe_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}
seq=[]
for r in rows[1:]:
    n=r[ki].split('(')[0][:40]; v=float(r[vi].replace(',',''))
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v; seq.append((n,v))
for n,a in agg.items(): print(n, a[0], 'launches', round(a[1]/1e3,1), 'us total', round(a[1]/a[0]/1e3,1), 'us avg')
print([ (n[:12], round(v/1e3)) for n,v in seq[2:42]])
P
