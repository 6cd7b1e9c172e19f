#!/bin/bash
# ncu --set full of the two blend kernels of one fwd+bwd step -> one summary line each
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:blend -o /tmp/bl -f python tools/profile_step.py 1000000 ours 1 > /tmp/bl_ncu.log 2>&1
tail -1 /tmp/bl_ncu.log
python tools/ncu_summary.py /tmp/bl.ncu-rep gpurun_out/blend_ncu_summary.csv > /dev/null
cat gpurun_out/blend_ncu_summary.csv
