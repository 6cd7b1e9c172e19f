// microbench.cu — on-box SIMT peaks for the blend kernels' roofline ("the builder must measure
// an FFMA/MUFU microbenchmark on the box"; MEASURED_PEAKS.json only carries HBM and tensor-core numbers).
//
//   [0] FFMA        dense fp32 FMA issue rate, TFLOP/s (2 FLOP per lane-FMA; 8 independent chains per thread)
//   [1] MUFU.EX2    ex2.approx rate, G lane-ops/s (the blend kernels' exp goes through MUFU.EX2)
//   [2] MUFU.RCP    rcp.approx rate, G lane-ops/s (blend backward's 1/(1-alpha))
//   [3] LDS.128     warp-wide broadcast 16-byte shared loads (how the blend kernels fetch a staged entry),
//                   G warp-instructions/s
//   [4] SHFL.BFLY   warp shuffles (the backward's butterfly reduction), G warp-instructions/s
//   [5] ALU         integer add / logic ops (index arithmetic, bounds tests: the ALU pipe), G warp-instructions/s;
//                   the warp-instruction ISSUE ceiling itself is what [0] reaches: one FFMA per scheduler per clock
// Every kernel is timed with CUDA events on the caller's stream, best of 5, grid = 148 SMs x 4 blocks x 256 threads
// (8 warps per scheduler: enough to cover the 4-cycle dependent-issue latency with 8 chains).
#include "common.cuh"

namespace lg {

constexpr int MB_BLOCK = 256;
constexpr int MB_ITERS = 4096;

__global__ void __launch_bounds__(MB_BLOCK) mb_ffma_kernel(float* out, float a, float b) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < MB_ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = __fmaf_rn(x[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k];
    if (s == 123456.789f) out[0] = s;  // never true: keeps the chains alive
}

template <int OP>
__global__ void __launch_bounds__(MB_BLOCK) mb_mufu_kernel(float* out, float a) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = a + 0.001f * (float)(threadIdx.x + k);
    for (int i = 0; i < MB_ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float y;  // MUFU result + one FADD on the (otherwise idle) FMA pipe, so that rcp(rcp(x)) cannot be folded
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[k]));
            else asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[k]));
            x[k] = y + a;
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k];
    if (s == 123456.789f) out[0] = s;
}

__global__ void __launch_bounds__(MB_BLOCK) mb_lds_kernel(float* out, int stride) {
    __shared__ float4 s_buf[512];
    for (int i = threadIdx.x; i < 512; i += MB_BLOCK) s_buf[i] = make_float4((float)i, 1.0f, 2.0f, 3.0f);
    __syncthreads();
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    int j = (threadIdx.x >> 5) * 7;  // one address per warp: a broadcast
    for (int i = 0; i < MB_ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float4 v;
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&s_buf[(j + k * stride) & 511]);
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        j += 8 * stride;
    }
    if (acc.x + acc.y + acc.z + acc.w == 123456.789f) out[0] = acc.x;
}

__global__ void __launch_bounds__(MB_BLOCK) mb_shfl_kernel(float* out) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = (float)(threadIdx.x + k);
    for (int i = 0; i < MB_ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1 + (k & 3));
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k];
    if (s == 123456.789f) out[0] = s;
}

__global__ void __launch_bounds__(MB_BLOCK) mb_alu_kernel(float* out, int a) {
    int x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    for (int i = 0; i < MB_ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {  // alternate xor / add so that ptxas cannot merge consecutive operations
            if (k & 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
            else asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
            else asm volatile("add.s32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
        }
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    if (s == 0x7fffffff) out[0] = (float)s;
}

}  // namespace lg

using namespace lg;

extern "C" int lg_simt_peaks(float* peaks_out, int n, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!peaks_out || n < 6) {
        set_error("lg_simt_peaks: need room for 6 floats");
        return LG_ERR_INVALID_ARGUMENT;
    }
    float* d_out = nullptr;
    LG_CUDA(cudaMalloc((void**)&d_out, 64));
    cudaEvent_t e0, e1;
    LG_CUDA(cudaEventCreate(&e0));
    LG_CUDA(cudaEventCreate(&e1));
    const int blocks = LG_NUM_SMS * 4;
    const double lane_ops = (double)blocks * MB_BLOCK * MB_ITERS * 8.0;
    const double warp_ops = lane_ops / 32.0;
    for (int which = 0; which < 6; which++) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; rep++) {
            LG_CUDA(cudaEventRecord(e0, stream));
            switch (which) {
                case 0: mb_ffma_kernel<<<blocks, MB_BLOCK, 0, stream>>>(d_out, 0.999f, 0.001f); break;
                case 1: mb_mufu_kernel<0><<<blocks, MB_BLOCK, 0, stream>>>(d_out, -0.5f); break;
                case 2: mb_mufu_kernel<1><<<blocks, MB_BLOCK, 0, stream>>>(d_out, 1.5f); break;
                case 3: mb_lds_kernel<<<blocks, MB_BLOCK, 0, stream>>>(d_out, 3); break;
                case 4: mb_shfl_kernel<<<blocks, MB_BLOCK, 0, stream>>>(d_out); break;
                default: mb_alu_kernel<<<blocks, MB_BLOCK, 0, stream>>>(d_out, 3); break;
            }
            count_launch();
            LG_CUDA(cudaEventRecord(e1, stream));
            LG_CUDA(cudaEventSynchronize(e1));
            float ms = 0.0f;
            LG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;  // rep 0 warms up
        }
        const double s = best * 1e-3;
        if (which == 0) peaks_out[0] = (float)(2.0 * lane_ops / s / 1e12);
        else if (which == 1 || which == 2) peaks_out[which] = (float)(lane_ops / s / 1e9);
        else peaks_out[which] = (float)((which == 5 ? 2.0 : 1.0) * warp_ops / s / 1e9);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return LG_OK;
}
