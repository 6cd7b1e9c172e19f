#!/bin/bash
# round-2 evidence run: ncu --set full of one fwd+bwd step, launch list of the bench command
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# (compute-sanitizer is closed on this GPU pool: the attempt of this round printed its refusal, profiles/r2_sanitizer_closed.txt)
K='regex:preprocess|scatter|tile_|blend|mark_visible'
python tools/profile_step.py 1000000 ours 2 > /tmp/ps_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -o /tmp/step -f python tools/profile_step.py 1000000 ours 2 > /tmp/ps_ncu.log 2>&1
tail -1 /tmp/ps_ncu.log
python tools/ncu_summary.py /tmp/step.ncu-rep gpurun_out/r2_step_ncu_summary.csv
python tools/ncu_traffic.py /tmp/step.ncu-rep gpurun_out/r2_ncu_traffic.json 2
ncu -i /tmp/step.ncu-rep --page source --csv > /tmp/step_src.csv 2>/dev/null
python tools/ncu_sass.py /tmp/step_src.csv blend_backward_kernel 0.3 > gpurun_out/r2_blend_backward_sass_hotspots.txt
python tools/ncu_sass.py /tmp/step_src.csv blend_forward_kernel 0.3 > gpurun_out/r2_blend_forward_sass_hotspots.txt
python tools/ncu_sass.py /tmp/step_src.csv tile_sort_kernel 0.3 > gpurun_out/r2_tile_sort_sass_hotspots.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-iteration --no-cfg5 > /tmp/bl_plain.log 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 200 -c 200 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-iteration --no-cfg5 > /tmp/bl_ncu.log 2>&1
ls -la gpurun_out/ | grep r2_ | tail -12
