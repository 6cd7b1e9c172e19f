#!/bin/bash
# Final ncu evidence, computed on the GPU box so that only small summaries travel back (the .ncu-rep stays in /tmp).
set -x
K='regex:preprocess|rs_|emit|tile_|blend|mark_visible'
python tools/profile_step.py 1000000 ours 2 > /tmp/ps_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -o /tmp/step -f python tools/profile_step.py 1000000 ours 2 > /tmp/ps_ncu.log 2>&1
tail -1 /tmp/ps_ncu.log
python tools/ncu_summary.py /tmp/step.ncu-rep gpurun_out/r1_step_ncu_summary_all.csv
python tools/ncu_traffic.py /tmp/step.ncu-rep gpurun_out/r1_ncu_traffic.json 2
ncu -i /tmp/step.ncu-rep --page source --csv > /tmp/step_src.csv 2>/dev/null
python tools/ncu_sass.py /tmp/step_src.csv blend_backward_kernel 0.3 > gpurun_out/r1_blend_backward_sass_hotspots.txt
python tools/ncu_sass.py /tmp/step_src.csv blend_forward_kernel 0.3 > gpurun_out/r1_blend_forward_sass_hotspots.txt
# launch list of the bench command (steady state: skip set-up + warm-up launches)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-iteration > /tmp/bl_plain.log 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 450 -c 400 --csv --log-file gpurun_out/r1_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-iteration > /tmp/bl_ncu.log 2>&1
ls -la gpurun_out/
