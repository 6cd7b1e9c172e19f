mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_ops_gpu.py tests/test_trainer_gpu.py -x -q --timeout 300 > gpurun_out/r2_t13.log 2>&1
tail -4 gpurun_out/r2_t13.log
timeout -s KILL 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench13_ref.json 2> gpurun_out/r2_bench13_ref.err; tail -2 gpurun_out/r2_bench13_ref.err
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; tail -3 gpurun_out/r2_bench13.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2_bench13_ref.json')); d=json.load(open('gpurun_out/r2_bench13.json'))
print("ref  value %.1f e2e %.1f"%(r['value'], r['e2e']['value']), r.get('image_loss'))
print("ours value %.1f e2e %.1f launches %d  dropin %.1f / %.1f"%(d['value'], d['e2e']['value'], d['gpu_launches'], d['dropin']['value'], d['dropin']['e2e_value']))
print("ratio value %.2f e2e %.2f dropin-e2e %.2f"%(d['value']/r['value'], d['e2e']['value']/r['e2e']['value'], d['dropin']['e2e_value']/r['e2e']['value']))
print({k:v['ms'] for k,v in d['stages'].items()}); print(d.get('cfg5')); print(d.get('image_loss')); print(d['train_iteration']['ms_per_iteration'], d['train_iteration']['breakdown_ms'], r['train_iteration'].get('ms_per_iteration'))
PY
