// photometric.cu — fused photometric loss terms of the LGDWT-GS iteration: mean |pred - gt| (l1_loss,
// LG/utils/loss_utils.py:40-41) and the mean SSIM map (ssim / _ssim, LG/utils/loss_utils.py:58-86: 11x11 Gaussian
// window, sigma 1.5, zero "same" padding, C1 = 0.01^2, C2 = 0.03^2), forward and backward, in two launches instead of
// the ~25 PyTorch launches of the reference path (5 grouped conv2d + elementwise + 3 reductions, and their autograd).
// The reference's optional CUDA path for the same term is gaussian-splatting/submodules/fused-ssim/ssim.cu (bound as
// diff_gaussian_rasterization._C.fusedssim, LG/utils/loss_utils.py:16-37); semantics are identical.
//
// Forward: a CTA owns a PH_BX x PH_BY output tile of one channel plane, stages the (tile + 5-pixel halo) of both
// images in shared memory, runs the separable 11-tap filter over the five moment maps (x, y, x^2, y^2, x*y) and
// evaluates SSIM per pixel.  It also stores, per pixel, the three partial derivatives of the SSIM value with respect
// to the windowed moments (E[x], E[x^2], E[xy]); the backward kernel filters those three maps with the same (symmetric)
// window — the adjoint of the forward filter — and adds the L1 sign term:
//   dL/dx(p) = g_ssim/N * sum_q w(q-p) [ dm/dE[x](q) + 2 x(p) dm/dE[x^2](q) + y(p) dm/dE[xy](q) ] + g_l1/N * sign(x-y).
#include "common.cuh"
#include <math.h>

namespace lg {

#define PH_BX 32
#define PH_BY 16
#define PH_R 5                       // window radius (11 taps)
#define PH_SX (PH_BX + 2 * PH_R)     // staged tile width
#define PH_SY (PH_BY + 2 * PH_R)     // staged tile height
#define PH_THREADS (PH_BX * PH_BY)

__constant__ float c_gauss[2 * PH_R + 1];

struct PhotoWorkspace {
    double* sums;   // [0] sum |x-y|, [1] sum ssim_map
    float* dmaps;   // 3 * C * H * W: dm/dE[x], dm/dE[x^2], dm/dE[xy]
    static PhotoWorkspace from_chunk(char*& chunk, int C, int H, int W) {
        PhotoWorkspace w;
        carve(chunk, w.sums, 2);
        carve(chunk, w.dmaps, (size_t)3 * C * H * W);
        return w;
    }
};

__global__ void __launch_bounds__(PH_THREADS) photometric_forward_kernel(const float* __restrict__ pred,
                                                                         const float* __restrict__ gt, int H, int W,
                                                                         float C1, float C2, double* __restrict__ sums,
                                                                         float* __restrict__ dmaps, size_t plane_stride_maps) {
    __shared__ float s_x[PH_SY][PH_SX + 1];
    __shared__ float s_y[PH_SY][PH_SX + 1];
    __shared__ float s_h[5][PH_SY][PH_BX + 1];  // horizontally filtered moments
    __shared__ double s_red[2][PH_THREADS / 32];

    const int tx = threadIdx.x % PH_BX, ty = threadIdx.x / PH_BX;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * PH_BX, y0 = blockIdx.y * PH_BY;
    const float* px = pred + (size_t)plane * H * W;
    const float* py = gt + (size_t)plane * H * W;

    // stage tile + halo, zero outside the image (conv2d zero padding, loss_utils.py:59-60)
    for (int i = threadIdx.x; i < PH_SY * PH_SX; i += PH_THREADS) {
        const int ly = i / PH_SX, lx = i % PH_SX;
        const int gy = y0 + ly - PH_R, gx = x0 + lx - PH_R;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        s_x[ly][lx] = in ? px[(size_t)gy * W + gx] : 0.0f;
        s_y[ly][lx] = in ? py[(size_t)gy * W + gx] : 0.0f;
    }
    __syncthreads();
    // horizontal pass over all staged rows
    for (int i = threadIdx.x; i < PH_SY * PH_BX; i += PH_THREADS) {
        const int ly = i / PH_BX, lx = i % PH_BX;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * PH_R + 1; k++) {
            const float w = c_gauss[k];
            const float x = s_x[ly][lx + k], y = s_y[ly][lx + k];
            a0 = fmaf(w, x, a0);
            a1 = fmaf(w, y, a1);
            a2 = fmaf(w, x * x, a2);
            a3 = fmaf(w, y * y, a3);
            a4 = fmaf(w, x * y, a4);
        }
        s_h[0][ly][lx] = a0; s_h[1][ly][lx] = a1; s_h[2][ly][lx] = a2; s_h[3][ly][lx] = a3; s_h[4][ly][lx] = a4;
    }
    __syncthreads();
    // vertical pass + SSIM
    const int gx = x0 + tx, gy = y0 + ty;
    double l1 = 0.0, ss = 0.0;
    if (gx < W && gy < H) {
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 2 * PH_R + 1; k++) {
            const float w = c_gauss[k];
#pragma unroll
            for (int q = 0; q < 5; q++) m[q] = fmaf(w, s_h[q][ty + k][tx], m[q]);
        }
        const float mu1 = m[0], mu2 = m[1];
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float sigma1_sq = m[2] - mu1_sq, sigma2_sq = m[3] - mu2_sq, sigma12 = m[4] - mu1_mu2;
        const float A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
        const float Cc = 2.f * mu1_mu2 + C1, D = 2.f * sigma12 + C2;
        const float inv_AB = 1.0f / (A * B);
        const float ssim = Cc * D * inv_AB;
        ss = (double)ssim;
        const float x = s_x[ty + PH_R][tx + PH_R], y = s_y[ty + PH_R][tx + PH_R];
        l1 = (double)fabsf(x - y);
        if (dmaps) {
            // partials w.r.t. the windowed moments P = E[x], Q = E[x^2], R = E[xy] (sigma1_sq = Q - P^2,
            // sigma12 = R - P mu2)
            const float dm_dP = 2.f * mu2 * (D - Cc) * inv_AB - 2.f * mu1 * ssim / A + 2.f * mu1 * ssim / B;
            const float dm_dQ = -ssim / B;
            const float dm_dR = 2.f * Cc * inv_AB;
            const size_t o = (size_t)plane * H * W + (size_t)gy * W + gx;
            dmaps[o] = dm_dP;
            dmaps[o + plane_stride_maps] = dm_dQ;
            dmaps[o + 2 * plane_stride_maps] = dm_dR;
        }
    }
    // block reduction of the two sums (double accumulation: the result must not depend on the launch geometry to
    // more than fp64 rounding)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[0][warp] = l1; s_red[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < PH_THREADS / 32; w++) { a += s_red[0][w]; b += s_red[1][w]; }
        atomicAdd(&sums[0], a);
        atomicAdd(&sums[1], b);
    }
}

__global__ void photometric_finalize_kernel(const double* __restrict__ sums, double inv_n, float* __restrict__ out) {
    out[0] = (float)(sums[0] * inv_n);  // mean |pred - gt|
    out[1] = (float)(sums[1] * inv_n);  // mean SSIM
}

__global__ void __launch_bounds__(PH_THREADS) photometric_backward_kernel(const float* __restrict__ pred,
                                                                          const float* __restrict__ gt, int H, int W,
                                                                          const float* __restrict__ dmaps,
                                                                          size_t plane_stride_maps,
                                                                          const float* __restrict__ g_l1,
                                                                          const float* __restrict__ g_ssim,
                                                                          const float* __restrict__ g_up, float inv_n,
                                                                          float* __restrict__ dL_dpred) {
    __shared__ float s_m[3][PH_SY][PH_SX + 1];
    __shared__ float s_h[3][PH_SY][PH_BX + 1];
    const int tx = threadIdx.x % PH_BX, ty = threadIdx.x / PH_BX;
    const int plane = blockIdx.z;
    const int x0 = blockIdx.x * PH_BX, y0 = blockIdx.y * PH_BY;
    const size_t pbase = (size_t)plane * H * W;

    for (int i = threadIdx.x; i < PH_SY * PH_SX; i += PH_THREADS) {
        const int ly = i / PH_SX, lx = i % PH_SX;
        const int gy = y0 + ly - PH_R, gx = x0 + lx - PH_R;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const size_t o = pbase + (size_t)gy * W + gx;
#pragma unroll
        for (int q = 0; q < 3; q++) s_m[q][ly][lx] = in ? dmaps[o + q * plane_stride_maps] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PH_SY * PH_BX; i += PH_THREADS) {
        const int ly = i / PH_BX, lx = i % PH_BX;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * PH_R + 1; k++) {
            const float w = c_gauss[k];
            a0 = fmaf(w, s_m[0][ly][lx + k], a0);
            a1 = fmaf(w, s_m[1][ly][lx + k], a1);
            a2 = fmaf(w, s_m[2][ly][lx + k], a2);
        }
        s_h[0][ly][lx] = a0; s_h[1][ly][lx] = a1; s_h[2][ly][lx] = a2;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < W && gy < H) {
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * PH_R + 1; k++) {
            const float w = c_gauss[k];
            m0 = fmaf(w, s_h[0][ty + k][tx], m0);
            m1 = fmaf(w, s_h[1][ty + k][tx], m1);
            m2 = fmaf(w, s_h[2][ty + k][tx], m2);
        }
        const size_t o = pbase + (size_t)gy * W + gx;
        const float x = pred[o], y = gt[o];
        const float up = g_up ? *g_up : 1.0f;  // upstream gradient of the combined loss (device scalar)
        const float gs = g_ssim ? (*g_ssim * up) * inv_n : 0.0f;
        const float gl = g_l1 ? (*g_l1 * up) * inv_n : 0.0f;
        const float d = x - y;
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);  // torch.abs backward: sign(0) = 0
        dL_dpred[o] = gs * (m0 + 2.f * x * m1 + y * m2) + gl * sgn;
    }
}

static bool g_gauss_ready = false;
static int upload_window(cudaStream_t stream) {
    if (g_gauss_ready) return LG_OK;
    // gaussian(11, 1.5) of LG/utils/loss_utils.py:46-48: fp32 taps exp(-(i-5)^2 / (2 sigma^2)), normalised by their
    // fp32 sum
    float g[2 * PH_R + 1];
    float s = 0.f;
    for (int i = 0; i < 2 * PH_R + 1; i++) {
        g[i] = (float)exp(-(double)((i - PH_R) * (i - PH_R)) / (2.0 * 1.5 * 1.5));
        s += g[i];
    }
    for (int i = 0; i < 2 * PH_R + 1; i++) g[i] = g[i] / s;
    LG_CUDA(cudaMemcpyToSymbolAsync(c_gauss, g, sizeof(g), 0, cudaMemcpyHostToDevice, stream));
    LG_CUDA(cudaStreamSynchronize(stream));  // `g` lives on this stack frame
    g_gauss_ready = true;
    return LG_OK;
}

}  // namespace lg

using namespace lg;

extern "C" size_t lg_photometric_workspace_bytes(int C, int H, int W) {
    char* p = nullptr;
    PhotoWorkspace::from_chunk(p, C, H, W);
    return (size_t)p + 128;
}

extern "C" int lg_photometric_loss_forward(const float* pred, const float* gt, int C, int H, int W, float* out_losses,
                                           char* workspace, size_t workspace_bytes, int want_backward,
                                           void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!pred || !gt || !out_losses || !workspace || C <= 0 || H <= 0 || W <= 0 ||
        workspace_bytes < lg_photometric_workspace_bytes(C, H, W)) {
        set_error("lg_photometric_loss_forward: invalid arguments or workspace too small");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (C > 65535) {
        set_error("lg_photometric_loss_forward: at most 65535 channel planes");
        return LG_ERR_UNSUPPORTED;
    }
    int rc = upload_window(stream);
    if (rc != LG_OK) return rc;
    char* p = workspace;
    PhotoWorkspace w = PhotoWorkspace::from_chunk(p, C, H, W);
    LG_CUDA(cudaMemsetAsync(w.sums, 0, 2 * sizeof(double), stream));
    const dim3 grid((W + PH_BX - 1) / PH_BX, (H + PH_BY - 1) / PH_BY, C);
    photometric_forward_kernel<<<grid, PH_THREADS, 0, stream>>>(pred, gt, H, W, 0.01f * 0.01f, 0.03f * 0.03f, w.sums,
                                                                want_backward ? w.dmaps : nullptr, (size_t)C * H * W);
    LG_LAUNCH_CHECK(false, stream);
    photometric_finalize_kernel<<<1, 1, 0, stream>>>(w.sums, 1.0 / ((double)C * H * W), out_losses);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_photometric_loss_backward(const float* pred, const float* gt, int C, int H, int W,
                                            const char* workspace, const float* g_l1_dev, const float* g_ssim_dev,
                                            float* dL_dpred, void* stream_v) {
    return lg_photometric_loss_backward_scaled(pred, gt, C, H, W, workspace, g_l1_dev, g_ssim_dev, nullptr, dL_dpred,
                                               stream_v);
}

extern "C" int lg_photometric_loss_backward_scaled(const float* pred, const float* gt, int C, int H, int W,
                                                   const char* workspace, const float* g_l1_dev,
                                                   const float* g_ssim_dev, const float* g_up_dev, float* dL_dpred,
                                                   void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!pred || !gt || !workspace || !dL_dpred || C <= 0 || H <= 0 || W <= 0 || C > 65535) {
        set_error("lg_photometric_loss_backward: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    char* p = const_cast<char*>(workspace);
    PhotoWorkspace w = PhotoWorkspace::from_chunk(p, C, H, W);
    const dim3 grid((W + PH_BX - 1) / PH_BX, (H + PH_BY - 1) / PH_BY, C);
    photometric_backward_kernel<<<grid, PH_THREADS, 0, stream>>>(pred, gt, H, W, w.dmaps, (size_t)C * H * W, g_l1_dev,
                                                                 g_ssim_dev, g_up_dev, (float)(1.0 / ((double)C * H * W)),
                                                                 dL_dpred);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}
