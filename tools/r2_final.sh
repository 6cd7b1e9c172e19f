#!/bin/bash
# Round-2 evidence in one GPU call: the -m gpu tier, both bench arms, the ncu launch list of the bench command and one
# `ncu --set full` capture of a fwd+bwd step (summaries only travel back; the .ncu-rep stays in /tmp).
# Usage: tools/r2_final.sh [tag]      -> gpurun_out/<tag>_*
TAG=${1:-r2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "bench reference rc=$?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err
echo "bench ours rc=$?"; head -c 1500 gpurun_out/${TAG}_bench_ours.json
if [ -z "$SKIP_NCU" ]; then
# launch list of the same bench command, steady state only (the library counts its own launches: skip everything before
# the timed resident loop is impractical, so the list holds the LAST 400 launches before exit = cfg5 excluded by flags)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 450 -c 400 --csv \
  --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train-iteration --no-cfg5 \
  > /tmp/bl_ncu.log 2>&1
echo "launch list rc=$?"
python tools/profile_step.py 1000000 ours 2 > /tmp/ps_plain.log 2>&1 || { echo "profile_step failed"; tail -5 /tmp/ps_plain.log; exit 1; }
K='regex:preprocess|scatter|tile_|blend'
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -o /tmp/step -f python tools/profile_step.py 1000000 ours 2 > /tmp/ps_ncu.log 2>&1
tail -1 /tmp/ps_ncu.log
python tools/ncu_summary.py /tmp/step.ncu-rep gpurun_out/${TAG}_step_ncu_summary.csv > /dev/null
python tools/ncu_traffic.py /tmp/step.ncu-rep gpurun_out/${TAG}_ncu_traffic.json 2
ncu -i /tmp/step.ncu-rep --page source --csv > /tmp/step_src.csv 2>/dev/null
python tools/ncu_sass.py /tmp/step_src.csv blend_backward_kernel 0.3 > gpurun_out/${TAG}_blend_backward_sass_hotspots.txt
python tools/ncu_sass.py /tmp/step_src.csv blend_forward_kernel 0.3 > gpurun_out/${TAG}_blend_forward_sass_hotspots.txt
python tools/ncu_sass.py /tmp/step_src.csv tile_sort_kernel 0.3 > gpurun_out/${TAG}_tile_sort_sass_hotspots.txt
fi
ls -la gpurun_out | grep ${TAG}_
