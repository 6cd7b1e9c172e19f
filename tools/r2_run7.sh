mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_rasterizer_vs_reference_gpu.py tests/test_ops_gpu.py tests/test_trainer_gpu.py -x -q --timeout 900 > gpurun_out/r2_t7.log 2>&1
tail -5 gpurun_out/r2_t7.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-iteration > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; tail -3 gpurun_out/r2_bench7.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench7.json')); print(d['value'], d['ms_per_view'], d['e2e']['value'], d['gpu_launches']); print({k:v['ms'] for k,v in d['stages'].items()})"
