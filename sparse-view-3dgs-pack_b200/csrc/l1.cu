// l1.cu — mean absolute error and its gradient in two launches.  Replaces l1_loss (LG/utils/loss_utils.py:40-41:
// torch.abs(network_output - gt).mean()) and its autograd backward (sub, abs, mean forward; expand, sign-mul,
// neg/identity backward = 7-9 elementwise PyTorch launches over the image) where only the L1 term is wanted;
// fused_photometric_loss (photometric.cu) is the op when SSIM is wanted too.  HBM-bound: 8 B/element forward,
// 12 B/element backward.
#include "common.cuh"

namespace lg {

// sums[0] = running sum (double), sums[1] = ticket counter (as double bits unused) — kept in a 16-byte scratch
__global__ void __launch_bounds__(256) l1_forward_kernel(const float4* __restrict__ a4, const float4* __restrict__ b4,
                                                         const float* __restrict__ a, const float* __restrict__ b,
                                                         long long n, double* __restrict__ sum,
                                                         unsigned* __restrict__ ticket, float* __restrict__ out) {
    __shared__ double s_part[8];
    const long long n4 = n / 4, stride = (long long)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        const float4 x = a4[q], y = b4[q];
        acc += (double)(fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w));
    }
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)(n - 4 * n4)) {
        const long long i = 4 * n4 + threadIdx.x;
        acc += (double)fabsf(a[i] - b[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += s_part[w];
        atomicAdd(sum, t);
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {  // last block: every partial sum is in
            __threadfence();
            out[0] = (float)(*(volatile double*)sum / (double)n);
        }
    }
}

// dL/dpred = g * sign(pred - gt) / n   (torch's sign: 0 at 0)
__global__ void __launch_bounds__(256) l1_backward_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          long long n, const float* __restrict__ g, float inv_n,
                                                          float* __restrict__ out) {
    const float s = g[0] * inv_n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = a[i] - b[i];
        out[i] = d > 0.0f ? s : (d < 0.0f ? -s : 0.0f);
    }
}

}  // namespace lg

using namespace lg;

extern "C" int lg_l1_loss_forward(const float* pred, const float* gt, long long n, float* out_loss, char* scratch16,
                                  void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!pred || !gt || !out_loss || !scratch16 || n <= 0 || (((uintptr_t)pred | (uintptr_t)gt) & 15u) ||
        ((uintptr_t)scratch16 & 7u)) {
        set_error("lg_l1_loss_forward: invalid arguments (16-byte aligned inputs, 8-byte aligned 16-byte scratch)");
        return LG_ERR_INVALID_ARGUMENT;
    }
    LG_CUDA(cudaMemsetAsync(scratch16, 0, 16, stream));
    const long long want = (n / 4 + 255) / 256;
    const int blocks = (int)(want < (long long)LG_NUM_SMS * 4 ? (want > 0 ? want : 1) : (long long)LG_NUM_SMS * 4);
    l1_forward_kernel<<<blocks, 256, 0, stream>>>((const float4*)pred, (const float4*)gt, pred, gt, n, (double*)scratch16,
                                                  (unsigned*)(scratch16 + 8), out_loss);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_l1_loss_backward(const float* pred, const float* gt, long long n, const float* g_dev,
                                   float* dL_dpred, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!pred || !gt || !g_dev || !dL_dpred || n <= 0) {
        set_error("lg_l1_loss_backward: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const long long want = (n + 255) / 256;
    const int blocks = (int)(want < (long long)LG_NUM_SMS * 8 ? want : (long long)LG_NUM_SMS * 8);
    l1_backward_kernel<<<blocks, 256, 0, stream>>>(pred, gt, n, g_dev, (float)(1.0 / (double)n), dL_dpred);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

// ---------------------------------------------------------------- loss assembly (LG/train.py:188-202)
namespace lg {
// One thread: base = (1 - lambda) * L1 + lambda * (1 - SSIM); the running-mean ratio that rescales the DWT term
// (rm <- 0.95 rm + 0.05 base / (dwt + 1e-8), used clamped to [0.1, 10] in the same iteration); loss = base + scale * dwt
// + patch_weight * patch.  Also leaves d(loss)/d(l1, ssim, dwt, patch) for the backward.  Replaces a dozen scalar
// PyTorch kernels and the reference's `.item()` host round trip (LG/train.py:193).
__global__ void image_loss_combine_kernel(const float* __restrict__ photometric, const float* __restrict__ dwt_out,
                                          float* running_mean, float lambda_dssim, float patch_weight,
                                          int update_running_mean, float* __restrict__ loss_out,
                                          float* __restrict__ coef_out) {
    const float l1 = photometric[0], ssim = photometric[1], dwt = dwt_out[0], patch = dwt_out[1];
    const float base = (1.0f - lambda_dssim) * l1 + lambda_dssim * (1.0f - ssim);
    float rm = running_mean[0];
    if (update_running_mean) {
        rm = 0.95f * rm + 0.05f * (base / (dwt + 1e-8f));
        running_mean[0] = rm;
    }
    const float scale = fminf(fmaxf(rm, 0.1f), 10.0f);
    loss_out[0] = base + scale * dwt + patch_weight * patch;
    loss_out[1] = base;
    coef_out[0] = 1.0f - lambda_dssim;
    coef_out[1] = -lambda_dssim;
    coef_out[2] = scale;
    coef_out[3] = patch_weight;
}
// coefficients of the backward: d(loss)/d(term) * upstream gradient (device scalar)
__global__ void image_loss_scale_kernel(const float* __restrict__ coef, const float* __restrict__ g, float* __restrict__ out) {
    if (threadIdx.x < 4) out[threadIdx.x] = coef[threadIdx.x] * g[0];
}
// dL_dpred = a + b over n floats (the two backward kernels write one image each)
__global__ void __launch_bounds__(256) image_loss_add_kernel(float4* __restrict__ a, const float4* __restrict__ b, long long n4,
                                                             float* a_tail, const float* b_tail, int tail) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 x = a[i];
        const float4 y = b[i];
        x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
        a[i] = x;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) a_tail[threadIdx.x] += b_tail[threadIdx.x];
}
}  // namespace lg

extern "C" int lg_image_loss_combine(const float* photometric_out, const float* dwt_out, float* running_mean,
                                     float lambda_dssim, float patch_weight, int update_running_mean, float* loss_out,
                                     float* coef_out, void* stream_v) {
    if (!photometric_out || !dwt_out || !running_mean || !loss_out || !coef_out) {
        set_error("lg_image_loss_combine: null pointer");
        return LG_ERR_INVALID_ARGUMENT;
    }
    image_loss_combine_kernel<<<1, 1, 0, (cudaStream_t)stream_v>>>(photometric_out, dwt_out, running_mean, lambda_dssim,
                                                                  patch_weight, update_running_mean, loss_out, coef_out);
    LG_LAUNCH_CHECK(false, (cudaStream_t)stream_v);
    return LG_OK;
}

extern "C" int lg_image_loss_backward_coefs(const float* coef, const float* g, float* out4, void* stream_v) {
    if (!coef || !g || !out4) {
        set_error("lg_image_loss_backward_coefs: null pointer");
        return LG_ERR_INVALID_ARGUMENT;
    }
    image_loss_scale_kernel<<<1, 32, 0, (cudaStream_t)stream_v>>>(coef, g, out4);
    LG_LAUNCH_CHECK(false, (cudaStream_t)stream_v);
    return LG_OK;
}

extern "C" int lg_image_loss_add(float* a, const float* b, long long n, void* stream_v) {
    if (!a || !b || n < 0 || (((uintptr_t)a | (uintptr_t)b) & 15u)) {
        set_error("lg_image_loss_add: invalid arguments (16-byte aligned buffers)");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return LG_OK;
    const long long n4 = n / 4;
    const long long want = (n4 + 255) / 256;
    const int blocks = (int)(want < (long long)LG_NUM_SMS * 8 ? (want > 0 ? want : 1) : (long long)LG_NUM_SMS * 8);
    image_loss_add_kernel<<<blocks, 256, 0, (cudaStream_t)stream_v>>>((float4*)a, (const float4*)b, n4, a + 4 * n4, b + 4 * n4,
                                                                      (int)(n - 4 * n4));
    LG_LAUNCH_CHECK(false, (cudaStream_t)stream_v);
    return LG_OK;
}

// The whole iteration loss in ONE library call each way (LG/train.py:128-202): the host enters the library once per
// direction instead of seven times (at 3x800x800 the seven calls and their temporaries cost more host time than the
// kernels take on the GPU).  `terms` (24 floats): [0:2] L1, SSIM | [2:14] lg_dwt_loss_forward's out_losses |
// [14:16] loss, base | [16:20] d(loss)/d(L1, SSIM, dwt, patch) | [20:24] spare.
extern "C" int lg_image_loss_forward(const float* pred, const float* gt, int C, int H, int W,
                                     const float* band_weights_host, int patch_size, double percentile, float patch_w_lh,
                                     float patch_w_hl, float* running_mean, float lambda_dssim, float patch_weight,
                                     int update_running_mean, float* terms, uint8_t* patch_mask, size_t patch_mask_bytes,
                                     char* photometric_workspace, size_t photometric_bytes, char* dwt_workspace,
                                     size_t dwt_bytes, int want_backward, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!terms || !running_mean) {
        set_error("lg_image_loss_forward: null pointer");
        return LG_ERR_INVALID_ARGUMENT;
    }
    LG_CUDA(cudaMemsetAsync(terms, 0, 24 * sizeof(float), stream));
    if (patch_mask && patch_mask_bytes) LG_CUDA(cudaMemsetAsync(patch_mask, 0, patch_mask_bytes, stream));
    int rc = lg_photometric_loss_forward(pred, gt, C, H, W, terms, photometric_workspace, photometric_bytes, want_backward,
                                         stream_v);
    if (rc != LG_OK) return rc;
    rc = lg_dwt_loss_forward(pred, gt, C, H, W, band_weights_host, patch_size, percentile, patch_w_lh, patch_w_hl,
                             terms + 2, patch_mask, dwt_workspace, dwt_bytes, stream_v);
    if (rc != LG_OK) return rc;
    return lg_image_loss_combine(terms, terms + 2, running_mean, lambda_dssim, patch_weight, update_running_mean,
                                 terms + 14, terms + 16, stream_v);
}

// dL/dpred of the combined loss: the photometric kernel writes the image, the wavelet kernel adds its part in place;
// both scale their coefficient (terms[16:20]) by the upstream gradient `g_up` (device scalar) themselves.
extern "C" int lg_image_loss_backward(const float* pred, const float* gt, int C, int H, int W,
                                      const float* band_weights_host, int patch_size, float patch_w_lh, float patch_w_hl,
                                      const float* terms, const float* g_up, const uint8_t* patch_mask,
                                      const char* photometric_workspace, float* dL_dpred, void* stream_v) {
    if (!terms || !g_up) {
        set_error("lg_image_loss_backward: null pointer");
        return LG_ERR_INVALID_ARGUMENT;
    }
    int rc = lg_photometric_loss_backward_scaled(pred, gt, C, H, W, photometric_workspace, terms + 16, terms + 17, g_up,
                                                 dL_dpred, stream_v);
    if (rc != LG_OK) return rc;
    return lg_dwt_loss_backward_scaled(pred, gt, C, H, W, band_weights_host, patch_size, patch_w_lh, patch_w_hl, terms + 18,
                                       terms + 19, g_up, patch_mask, terms + 2, dL_dpred, 1, stream_v);
}
