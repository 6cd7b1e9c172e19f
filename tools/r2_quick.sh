#!/bin/bash
# quick GPU check of a change: selected tests ($1 = pytest -k expression, "all" = whole GPU tier), then the bench
# without its slow sections
mkdir -p gpurun_out
if [ "${1:-all}" = "all" ]; then
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E |passed|failed" | head -30
else
timeout 600 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | grep -E "^E |passed|failed" | head -30
fi
timeout 400 python bench.py --no-cfg5 ${BENCH_FLAGS} > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err || tail -5 gpurun_out/quick_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "dropin", d.get("dropin",{}).get("value"), d.get("dropin",{}).get("e2e_value"), "overlap", d.get("view_overlap",{}).get("value"), d.get("view_overlap",{}).get("e2e_value"))
print({k:v["ms"] for k,v in d["stages"].items()})
if "image_loss" in d: print("image_loss", {k:d["image_loss"].get(k) for k in ("gpu_fused_ms","gpu_fused_device_ms","gpu_launches","photometric_fwd_bwd_ms","photometric_fwd_bwd_device_ms")})
print("train_iteration", d.get("train_iteration",{}).get("ms_per_iteration"), d.get("train_iteration",{}).get("breakdown_ms"))
PY
