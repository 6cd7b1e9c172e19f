#!/bin/bash
# two-stream view overlap: the new trainer test, then the bench with one and with two view streams (same box)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -x -q -k "two_stream" 2>&1 | tail -5
for n in 1 2 1 2; do
LGDWT_VIEW_STREAMS=$n timeout 300 python bench.py --no-cpu-baseline --no-train-iteration --no-cfg5 > gpurun_out/r2b_bench_vs$n.json 2> gpurun_out/r2b_bench_vs$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_bench_vs$n.json").read().strip().splitlines()[-1])
print("streams $n: value", d["value"], "e2e", d["e2e"]["value"], "step_ms", d["step_ms"], "dropin", d.get("dropin",{}).get("value"))
print({k:v["ms"] for k,v in d["stages"].items()})
PY
done
