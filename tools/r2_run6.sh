mkdir -p gpurun_out
python tools/profile_step.py 1000000 ours 3 > gpurun_out/plain_r2_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-tile_sort}" -s ${KSKIP:-2} -c ${KCOUNT:-2} -o gpurun_out/prof_r2_binning -f python tools/profile_step.py 1000000 ours 3 > gpurun_out/ncu_r2_prof.log 2>&1
tail -3 gpurun_out/ncu_r2_prof.log; ls -la gpurun_out/*.ncu-rep | tail -3
