#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun):
#   * the stock torch extensions `dgr_3dgs` (diff-gaussian-rasterization) and `sknn_3dgs` (simple-knn), built by the
#     reference's own setup.py (plus `fused_ssim` from submodules/fused-ssim, the optional import of LG/train.py:36-40),
#     each built by its own setup.py from a scratch copy under /tmp (the source tree is read-only), for sm_100a;
#   * the reference's Python applications LGDWT-GS/ (train.py, render.py, scene/, utils/ ...) and LGDWT-GS/mult-dwtgs/
#     (train_nir.py), copied as they are;
#   * two alias packages `diff_gaussian_rasterization` / `simple_knn` that re-export the stock extensions under the
#     names the LGDWT-GS callers import (what fs3dgs_benchmark/post_install does on a user's machine, readme.md:163-164).
# Nothing here is product source: bench.py --impl reference and tests/test_callers_gpu.py use it as the reference arm.
# `-include cstdint`: DGR/cuda_rasterizer/rasterizer_impl.h uses std::uintptr_t without <cstdint> (gcc 13).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
SUB="$REF/fs3dgs_benchmark/gaussian-splatting/submodules"
if [ ! -d "$SUB/diff-gaussian-rasterization" ]; then
  echo "reference sources not found under $REF; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/site"
SCRATCH="$(mktemp -d /tmp/refbuild.XXXXXX)"
trap 'rm -rf "$SCRATCH"' EXIT
export TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS="${MAX_JOBS:-8}" NVCC_APPEND_FLAGS="-include cstdint"
build_ext() {  # $1 = submodule dir, $2 = package name
  if ls "$OUT/site/$2"/_C*.so >/dev/null 2>&1; then return; fi
  cp -r "$SUB/$1" "$SCRATCH/$1"
  (cd "$SCRATCH/$1" && python setup.py build_ext --inplace >"$SCRATCH/$2.log" 2>&1) || { tail -30 "$SCRATCH/$2.log"; exit 1; }
  mkdir -p "$OUT/site/$2"
  cp "$SCRATCH/$1/$2"/*.py "$SCRATCH/$1/$2"/_C*.so "$OUT/site/$2/"
}
build_fused_ssim() {  # top-level extension module fused_ssim_cuda + package fused_ssim (SSIM/setup.py)
  if ls "$OUT/site"/fused_ssim_cuda*.so >/dev/null 2>&1; then return; fi
  cp -r "$SUB/fused-ssim" "$SCRATCH/fused-ssim"
  (cd "$SCRATCH/fused-ssim" && python setup.py build_ext --inplace >"$SCRATCH/fused_ssim.log" 2>&1) || { tail -30 "$SCRATCH/fused_ssim.log"; exit 1; }
  mkdir -p "$OUT/site/fused_ssim"
  cp "$SCRATCH/fused-ssim/fused_ssim"/*.py "$OUT/site/fused_ssim/"
  cp "$SCRATCH/fused-ssim"/fused_ssim_cuda*.so "$OUT/site/"
}
build_ext diff-gaussian-rasterization dgr_3dgs &
build_ext simple-knn sknn_3dgs &
build_fused_ssim &
wait
mkdir -p "$OUT/site/diff_gaussian_rasterization" "$OUT/site/simple_knn"
cat > "$OUT/site/diff_gaussian_rasterization/__init__.py" <<'PY'
# alias written by baseline/install_reference.sh: the stock reference rasterizer under the name LGDWT-GS imports
from dgr_3dgs import GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians, rasterize_gaussians  # noqa: F401
from dgr_3dgs import _C  # noqa: F401
PY
cat > "$OUT/site/simple_knn/__init__.py" <<'PY'
PY
cat > "$OUT/site/simple_knn/_C.py" <<'PY'
# alias written by baseline/install_reference.sh
from sknn_3dgs._C import distCUDA2  # noqa: F401
PY
rm -rf "$OUT/LGDWT-GS"
cp -r "$REF/fs3dgs_benchmark/LGDWT-GS" "$OUT/LGDWT-GS"
echo "$OUT"
