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
