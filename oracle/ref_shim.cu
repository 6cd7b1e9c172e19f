// ref_shim.cu — TEST INFRASTRUCTURE ONLY.  A thin extern "C" door onto the UNMODIFIED reference CUDA rasterizer and
// simple-knn, compiled by oracle/build_ref.sh from the reference's own source files where they lie under
// /root/reference (nothing is copied into this repo) into oracle/_ref/libref_dgr.so.
//
// It exists so that tests/ and bench.py can (a) compare this repo's kernels with the reference's on the same GPU
// ("oracle by execution": the reference has no CPU path and no tests of its own) and (b) time
// the reference arm.  The product library never links or loads it.
//
// Entry points mirror CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (DGR/cuda_rasterizer/rasterizer.h:24-97) and SimpleKNN::knn (KNN/simple_knn.h); ref_state_read exposes the
// reference's opaque buffers through its own fromChunk (DGR/cuda_rasterizer/rasterizer_impl.cu:155-194).
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <cuda_runtime.h>
#include "rasterizer.h"
#include "rasterizer_impl.h"
#include "simple_knn.h"

typedef char* (*ref_alloc_fn)(void* ctx, size_t bytes);

static thread_local char g_ref_err[512] = "";

extern "C" {

const char* ref_last_error(void) { return g_ref_err; }

int ref_rasterize_forward(ref_alloc_fn geom_alloc, void* geom_ctx, ref_alloc_fn binning_alloc, void* binning_ctx,
                          ref_alloc_fn img_alloc, void* img_ctx, int P, int D, int M, const float* background,
                          int width, int height, const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                          const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
                          float* out_invdepth, int antialiasing, int* radii, int debug, int* num_rendered) {
    try {
        std::function<char*(size_t)> gf = [=](size_t n) { return geom_alloc(geom_ctx, n); };
        std::function<char*(size_t)> bf = [=](size_t n) { return binning_alloc(binning_ctx, n); };
        std::function<char*(size_t)> imf = [=](size_t n) { return img_alloc(img_ctx, n); };
        *num_rendered = CudaRasterizer::Rasterizer::forward(
            gf, bf, imf, P, D, M, background, width, height, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
            prefiltered != 0, out_color, out_invdepth, antialiasing != 0, radii, debug != 0);
    } catch (const std::exception& e) {
        snprintf(g_ref_err, sizeof(g_ref_err), "%s", e.what());
        return 1;
    }
    return 0;
}

int ref_rasterize_backward(int P, int D, int M, int R, const float* background, int width, int height,
                           const float* means3D, const float* shs, const float* colors_precomp,
                           const float* opacities, const float* scales, float scale_modifier, const float* rotations,
                           const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                           const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer,
                           char* binning_buffer, char* img_buffer, const float* dL_dpix, const float* dL_invdepths,
                           float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                           float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                           float* dL_drot, int antialiasing, int debug) {
    try {
        CudaRasterizer::Rasterizer::backward(P, D, M, R, background, width, height, means3D, shs, colors_precomp,
                                             opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                                             projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer,
                                             binning_buffer, img_buffer, dL_dpix, dL_invdepths, dL_dmean2D, dL_dconic,
                                             dL_dopacity, dL_dcolor, dL_dinvdepth, dL_dmean3D, dL_dcov3D, dL_dsh,
                                             dL_dscale, dL_drot, antialiasing != 0, debug != 0);
    } catch (const std::exception& e) {
        snprintf(g_ref_err, sizeof(g_ref_err), "%s", e.what());
        return 1;
    }
    return 0;
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    return 0;
}

// name -> (device pointer, bytes) inside the reference's opaque buffers
int ref_state_read(const char* name, int P, int width, int height, int R, char* geom_buffer, char* binning_buffer,
                   char* img_buffer, void* dst, size_t dst_bytes) {
    const void* src = nullptr;
    size_t bytes = 0;
    const size_t N = (size_t)width * height;
    if (geom_buffer) {
        char* p = geom_buffer;
        CudaRasterizer::GeometryState g = CudaRasterizer::GeometryState::fromChunk(p, P);
        if (!strcmp(name, "depths")) { src = g.depths; bytes = 4 * (size_t)P; }
        else if (!strcmp(name, "means2D")) { src = g.means2D; bytes = 8 * (size_t)P; }
        else if (!strcmp(name, "cov3D")) { src = g.cov3D; bytes = 24 * (size_t)P; }
        else if (!strcmp(name, "conic_opacity")) { src = g.conic_opacity; bytes = 16 * (size_t)P; }
        else if (!strcmp(name, "rgb")) { src = g.rgb; bytes = 12 * (size_t)P; }
        else if (!strcmp(name, "clamped")) { src = g.clamped; bytes = 3 * (size_t)P; }
        else if (!strcmp(name, "tiles_touched")) { src = g.tiles_touched; bytes = 4 * (size_t)P; }
        else if (!strcmp(name, "point_offsets")) { src = g.point_offsets; bytes = 4 * (size_t)P; }
    }
    if (!src && img_buffer) {
        char* p = img_buffer;
        CudaRasterizer::ImageState s = CudaRasterizer::ImageState::fromChunk(p, N);
        if (!strcmp(name, "final_T")) { src = s.accum_alpha; bytes = 4 * N; }
        else if (!strcmp(name, "n_contrib")) { src = s.n_contrib; bytes = 4 * N; }
        else if (!strcmp(name, "ranges")) { src = s.ranges; bytes = 8 * (size_t)(((width + 15) / 16) * ((height + 15) / 16)); }
    }
    if (!src && binning_buffer) {
        char* p = binning_buffer;
        CudaRasterizer::BinningState b = CudaRasterizer::BinningState::fromChunk(p, R);
        if (!strcmp(name, "point_list")) { src = b.point_list; bytes = 4 * (size_t)R; }
        else if (!strcmp(name, "point_list_keys")) { src = b.point_list_keys; bytes = 8 * (size_t)R; }
    }
    if (!src) { snprintf(g_ref_err, sizeof(g_ref_err), "unknown array %s", name); return 1; }
    if (dst_bytes < bytes) { snprintf(g_ref_err, sizeof(g_ref_err), "dst too small for %s", name); return 1; }
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) { snprintf(g_ref_err, sizeof(g_ref_err), "%s", cudaGetErrorString(e)); return 2; }
    return 0;
}

int ref_knn_mean_dist2(int P, float* points, float* mean_dists) {
    SimpleKNN::knn(P, (float3*)points, mean_dists);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_ref_err, sizeof(g_ref_err), "%s", cudaGetErrorString(e)); return 2; }
    return 0;
}

}  // extern "C"
