#!/bin/bash
# ncu --set full with source of one fwd+bwd step -> per-CUDA-line listings of the non-blend kernels
mkdir -p gpurun_out
K='regex:preprocess|scatter|tile_'
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" -o /tmp/nb -f python tools/profile_step.py 1000000 ours 1 > /tmp/ps_ncu.log 2>&1
tail -1 /tmp/ps_ncu.log
python tools/ncu_summary.py /tmp/nb.ncu-rep gpurun_out/r2c_ncu_summary.csv
ncu -i /tmp/nb.ncu-rep --page source --print-source cuda --csv > /tmp/nb_cuda.csv 2>/dev/null
head -c 3000 /tmp/nb_cuda.csv > gpurun_out/r2c_cuda_head.txt
for k in tile_sort_kernel scatter_pairs preprocess_kernel preprocess_backward_kernel tile_scan_kernel; do
python tools/ncu_lines_cuda.py /tmp/nb_cuda.csv $k 0.4 > gpurun_out/r2c_lines_$k.txt
done
ncu -i /tmp/nb.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); si=h.index('Section Name')
for r in rows[1:]:
    if r[si] in ('Warp State Statistics','Scheduler Statistics','Occupancy','Launch Statistics','Memory Workload Analysis'): print(r[ki][:30], '|', r[mi], '|', r[vi])
" > gpurun_out/r2c_details.txt
ls -la gpurun_out | grep r2c_
