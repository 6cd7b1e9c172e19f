#!/bin/bash
# ncu --set full with source of one fwd+bwd step -> per-CUDA-line listings + warp-state details of the per-Gaussian kernels
mkdir -p gpurun_out
K=${1:-'regex:preprocess'}
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" -o /tmp/nb -f python tools/profile_step.py 1000000 ours 1 > /tmp/ps_ncu.log 2>&1
tail -1 /tmp/ps_ncu.log
ncu -i /tmp/nb.ncu-rep --page source --print-source cuda --csv > /tmp/nb_cuda.csv 2>/dev/null
for k in preprocess_kernel preprocess_backward_kernel scatter_pairs tile_sort_kernel tile_scan_kernel; do
python tools/ncu_lines_cuda.py /tmp/nb_cuda.csv $k 0.7 > gpurun_out/r2d_lines_$k.txt
done
ncu -i /tmp/nb.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); si=h.index('Section Name')
for r in rows[1:]:
    if r[si] in ('Warp State Statistics','Scheduler Statistics','Occupancy','Launch Statistics','Memory Workload Analysis','GPU Speed Of Light Throughput'): print(r[ki][:30], '|', r[mi], '|', r[vi])
" > gpurun_out/r2d_details.txt
ncu -i /tmp/nb.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keep=[i for i,c in enumerate(h) if 'issue_stalled' in c and 'pct' not in c or c=='Kernel Name']
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:40])
    vals=sorted(((float(r[i].replace(',','')) if r[i] else 0.0, h[i]) for i in keep if h[i]!='Kernel Name'), reverse=True)[:10]
    for v,n in vals: print('   %10.3f %s'%(v,n))
" > gpurun_out/r2d_stalls.txt
ls -la gpurun_out | grep r2d_
