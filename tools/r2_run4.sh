mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rasterizer_vs_reference_gpu.py -x -q --timeout 600 > gpurun_out/r2_t4.log 2>&1
tail -5 gpurun_out/r2_t4.log
for L in ${LIBS:-liblgdwt_b200.so}; do
  export LGDWT_LIBNAME=$L
  timeout 600 bash tools/launches.sh r2_$L > gpurun_out/r2_launchsum_$L.txt 2>&1; grep -v "Functor\|tor<\|nctor" gpurun_out/r2_launchsum_$L.txt | tail -14
done
