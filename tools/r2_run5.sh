mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rasterizer_vs_reference_gpu.py -x -q --timeout 600 > gpurun_out/r2_t5.log 2>&1
tail -5 gpurun_out/r2_t5.log
for L in ${LIBS:-liblgdwt_b200.so}; do
export LGDWT_LIBNAME=$L
echo "== $L"
python tools/tile_stats.py metric cfg3 cfg5 2>&1 | tee gpurun_out/r2_tile_stats_$L.txt
timeout 600 bash tools/launches.sh r2_$L > gpurun_out/r2_launchsum_$L.txt 2>&1; grep -v "Functor\|tor<\|nctor" gpurun_out/r2_launchsum_$L.txt | tail -12
done
