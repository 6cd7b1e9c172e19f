#!/bin/bash
# per-kernel device times of one fwd+bwd step (ncu, cold caches, serialised) -> gpurun_out/launches_$1.csv
python tools/profile_step.py 1000000 ours 3 > gpurun_out/plain_$1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$1.csv python tools/profile_step.py 1000000 ours 3 > gpurun_out/ncu_$1.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_$1.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0][-34:],float(r[vi].replace(',',''))/1000) for r in rows[h+1:] if len(r)>vi]
n=len(seq)//3
tot=0
for s in seq[-n:]:
    if "Functor" in s[0]: continue
    print("%-42s %8.1f us"%s); tot+=s[1]
print("total", tot)
PY
