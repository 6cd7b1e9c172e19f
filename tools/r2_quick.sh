#!/bin/bash
# quick GPU check of a change: selected tests, then the bench without its slow sections
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "${1:-loss or trainer or two_stream or photometric or dwt}" 2>&1 | grep -E "^E |passed|failed" | head -30
timeout 400 python bench.py --no-cfg5 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err || tail -5 gpurun_out/quick_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "dropin", d.get("dropin",{}).get("value"), d.get("dropin",{}).get("e2e_value"))
print({k:v["ms"] for k,v in d["stages"].items()})
print("image_loss", {k:d["image_loss"][k] for k in ("gpu_fused_ms","gpu_launches","photometric_fwd_bwd_ms")})
print("train_iteration", d.get("train_iteration",{}).get("ms_per_iteration"), d.get("train_iteration",{}).get("breakdown_ms"))
PY
