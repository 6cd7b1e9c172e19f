mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rasterizer_vs_reference_gpu.py tests/test_ops_gpu.py -x -q --timeout 600 > gpurun_out/r2_t3.log 2>&1
tail -25 gpurun_out/r2_t3.log
for L in liblgdwt_b200.so libv_ts2.so; do
  export LGDWT_LIBNAME=$L
  timeout 600 bash tools/launches.sh r2_$L > gpurun_out/r2_launchsum_$L.txt 2>&1; tail -22 gpurun_out/r2_launchsum_$L.txt
done
unset LGDWT_LIBNAME
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-iteration > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench3.json')); print(d['value'], d['ms_per_view'], d['e2e']['value']); print(json.dumps(d['stages']))"
