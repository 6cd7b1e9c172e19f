mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_ops_gpu.py tests/test_trainer_gpu.py tests/test_rasterizer_vs_reference_gpu.py -x -q --timeout 900 > gpurun_out/r2_t8.log 2>&1
tail -8 gpurun_out/r2_t8.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench8_ref.json 2> gpurun_out/r2_bench8_ref.err; tail -2 gpurun_out/r2_bench8_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; tail -3 gpurun_out/r2_bench8.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2_bench8_ref.json')); d=json.load(open('gpurun_out/r2_bench8.json'))
print("ref  value %.1f e2e %.1f  (%s)"%(r['value'], r['e2e']['value'], r['reference_kind'][:60]))
print("ours value %.1f e2e %.1f launches %d  dropin %s"%(d['value'], d['e2e']['value'], d['gpu_launches'], d.get('dropin')))
print("ratio value %.2f e2e %.2f dropin-e2e %.2f"%(d['value']/r['value'], d['e2e']['value']/r['e2e']['value'], d['dropin']['e2e_value']/r['e2e']['value']))
print({k:v['ms'] for k,v in d['stages'].items()}); print(d['roofline']); print(d.get('cfg5')); print(d.get('image_loss')); print(d['train_iteration']['ms_per_iteration'], r['train_iteration'].get('ms_per_iteration'))
PY
