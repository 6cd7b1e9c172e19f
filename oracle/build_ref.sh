#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the reference's own CUDA sources, in place under /root/reference, together
# with oracle/ref_shim.cu into oracle/_ref/libref_dgr.so (git-ignored, still shipped to the GPU box by gpurun).
# Nothing is copied from the reference into this repository.  `-include cstdint` is needed because
# DGR/cuda_rasterizer/rasterizer_impl.h uses std::uintptr_t without including <cstdint> (gcc 13).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
DGR="$REF/fs3dgs_benchmark/gaussian-splatting/submodules/diff-gaussian-rasterization"
KNN="$REF/fs3dgs_benchmark/gaussian-splatting/submodules/simple-knn"
OUT="$HERE/_ref"
if [ ! -d "$DGR/cuda_rasterizer" ]; then
  echo "reference sources not found under $REF; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
FLAGS=(-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -include cstdint -Xcompiler -fPIC
       -I"$DGR/third_party/glm" -I"$DGR/cuda_rasterizer" -I"$KNN" -diag-suppress 177)
pids=()
for src in "$DGR/cuda_rasterizer/forward.cu" "$DGR/cuda_rasterizer/backward.cu" \
           "$DGR/cuda_rasterizer/rasterizer_impl.cu" "$KNN/simple_knn.cu" "$HERE/ref_shim.cu"; do
  obj="$OUT/obj/$(basename "${src%.cu}").o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then
    nvcc "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libref_dgr.so" "$OUT"/obj/*.o -lcudart
echo "$OUT/libref_dgr.so"
