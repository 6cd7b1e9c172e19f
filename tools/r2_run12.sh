#!/bin/bash
# parity tests, bench, then ncu source-level profile of the two blend kernels (hot-spot listings)
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_ops_gpu.py tests/test_trainer_gpu.py tests/test_rasterizer_vs_reference_gpu.py tests/test_configs_gpu.py -x -q --timeout 400 > gpurun_out/r2_t12.log 2>&1
tail -4 gpurun_out/r2_t12.log
for L in ${LIBS:-liblgdwt_b200.so}; do
export LGDWT_LIBNAME=$L
echo "== $L"
timeout -s KILL 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-iteration --no-cfg5 > gpurun_out/r2_bench12_$L.json 2> gpurun_out/r2_bench12_$L.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench12_$L.json')); print(d['value'], d['ms_per_view'], d['e2e']['value'], d['gpu_launches'], d['dropin']['value']); print({k:v['ms'] for k,v in d['stages'].items()})"
done
unset LGDWT_LIBNAME
K='regex:blend_'
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -o /tmp/blend -f python tools/profile_step.py 1000000 ours 1 > /tmp/ps_ncu.log 2>&1
tail -1 /tmp/ps_ncu.log
python tools/ncu_summary.py /tmp/blend.ncu-rep gpurun_out/r2_blend_ncu_summary.csv
ncu -i /tmp/blend.ncu-rep --page source --csv > /tmp/blend_src.csv 2>/dev/null
python tools/ncu_sass.py /tmp/blend_src.csv blend_backward_kernel 0.0 > gpurun_out/r2b_blend_backward_sass.txt
python tools/ncu_sass.py /tmp/blend_src.csv blend_forward_kernel 0.0 > gpurun_out/r2b_blend_forward_sass.txt
ncu -i /tmp/blend.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|Warp Cycles Per\|No Eligible\|Eligible Warps\|Issued Warp" | cut -c1-300 > gpurun_out/r2b_blend_details.txt
ls -la gpurun_out | grep r2b_
