mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_trainer_gpu.py tests/test_rasterizer_vs_reference_gpu.py tests/test_configs_gpu.py -x -q --timeout 300 > gpurun_out/r2_t9.log 2>&1
tail -4 gpurun_out/r2_t9.log
for L in ${LIBS:-liblgdwt_b200.so libv_nobulk.so}; do
export LGDWT_LIBNAME=$L
echo "== $L"
timeout -s KILL 150 bash tools/launches.sh r2_$L > gpurun_out/r2_launchsum_$L.txt 2>&1; grep -v "Functor\|tor<\|nctor" gpurun_out/r2_launchsum_$L.txt | tail -10
timeout -s KILL 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-iteration --no-cfg5 > gpurun_out/r2_bench9_$L.json 2> gpurun_out/r2_bench9_$L.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench9_$L.json')); print(d['value'], d['ms_per_view'], d['e2e']['value'], d['gpu_launches'], d['dropin']['value']); print({k:v['ms'] for k,v in d['stages'].items()})"
done
