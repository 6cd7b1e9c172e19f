// stand-alone probe of mbarrier + bulk-copy patterns (each variant run in its own process under a timeout)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../sparse-view-3dgs-pack_b200/csrc/common.cuh"
namespace lg { void set_error(const char*, ...) {} void count_launch() {} int cuda_fail(cudaError_t, const char*, const char*, int) { return 1; } void stage_begin(int, cudaStream_t) {} void stage_end(int, cudaStream_t) {} }

// variant 1: lane 0 of each warp: init count 1, expect_tx(6144) + ONE 6144-byte copy of the warp's 32 dense rows
// variant 2: count 32, all lanes arrive, lane 0 expect_tx + one 6144-byte copy
// variant 3: count 32, even lanes expect_tx(384) + 384-byte copy each (pairs, padded layout)
// variant 4: count 1, lane 0 expect_tx(6144), then EVERY even lane issues a 384-byte copy (no per-lane arrive)
__global__ void __launch_bounds__(256) probe_kernel(int P, const float* __restrict__ rows, float* __restrict__ out, int variant, float* __restrict__ out2) {
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ __align__(8) uint64_t s_bar[8];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const unsigned count = (variant == 1 || variant == 4) ? 1u : 32u;
    if (lane == 0) { lg_mbar_init(&s_bar[warp], count); lg_mbar_init_fence(); }
    __syncthreads();
    float* dense = s_dyn + (size_t)warp * 32 * 50 * 1;  // room: 32 rows x 48 (+pad)
    if (variant == 1 || variant == 2) {
        if (lane == 0) {
            lg_mbar_arrive_expect_tx(&s_bar[warp], 6144);
            lg_bulk_load(s_dyn + (size_t)warp * 32 * 48, rows + (size_t)(idx) * 48, 6144, &s_bar[warp]);
        } else if (variant == 2) {
            lg_mbar_arrive(&s_bar[warp]);
        }
    } else if (variant == 3 || variant >= 8) {
        if ((lane & 1u) == 0) {
            lg_mbar_arrive_expect_tx(&s_bar[warp], 384);
            lg_bulk_load(lg_sh_row(s_dyn, threadIdx.x), rows + (size_t)idx * 48, 384, &s_bar[warp]);
        } else {
            lg_mbar_arrive(&s_bar[warp]);
        }
    } else if (variant >= 5) {
        // 5: some pairs skip (plain arrive by the even lane); 6: ragged tail (P not a multiple of 256, last copy 192 B); 7: both
        const bool skip = (variant == 5 || variant == 7) && ((idx >> 1) % 4 == 0);
        const bool in_range = idx < P;
        if (in_range && !skip && (lane & 1u) == 0) {
            const unsigned bytes = (idx + 1 < P ? 2u : 1u) * 192u;
            lg_mbar_arrive_expect_tx(&s_bar[warp], bytes);
            lg_bulk_load(lg_sh_row(s_dyn, threadIdx.x), rows + (size_t)idx * 48, bytes, &s_bar[warp]);
        } else {
            lg_mbar_arrive(&s_bar[warp]);
        }
    } else {
        if (lane == 0) lg_mbar_arrive_expect_tx(&s_bar[warp], 6144);
        __syncwarp();
        if ((lane & 1u) == 0) lg_bulk_load(lg_sh_row(s_dyn, threadIdx.x), rows + (size_t)idx * 48, 384, &s_bar[warp]);
    }
    (void)dense;
    lg_mbar_wait(&s_bar[warp], 0);
    const float* row = (variant <= 2) ? s_dyn + (size_t)threadIdx.x * 48 : lg_sh_row(s_dyn, threadIdx.x);
    if (variant >= 8) {
        // 8: v3 + fence.proxy.async; 9: + write 2x the row back with a bulk store; 10: bulk reduce-add instead
        float4* row4 = reinterpret_cast<float4*>(lg_sh_row(s_dyn, threadIdx.x));
        if (variant >= 9)
            for (int j = 0; j < 12; j++) { float4 v = row4[j]; row4[j] = make_float4(2 * v.x, 2 * v.y, 2 * v.z, 2 * v.w); }
        lg_fence_proxy_async();
        __syncwarp();
        if (variant >= 9 && (lane & 1u) == 0) {
            if (variant == 10) lg_bulk_reduce_add_f32(out2 + (size_t)idx * 48, lg_sh_row(s_dyn, threadIdx.x), 384);
            else lg_bulk_store(out2 + (size_t)idx * 48, lg_sh_row(s_dyn, threadIdx.x), 384);
            lg_bulk_commit();
            lg_bulk_wait_read();
        }
        __syncwarp();
    }
    float acc = 0.f;
    const bool have = variant < 5 || (idx < P && !((variant == 5 || variant == 7) && ((idx >> 1) % 4 == 0)));
    if (have) for (int j = 0; j < 48; j++) acc += row[j];
    if (idx < P) out[idx] = acc;
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 1;
    const int P = (variant == 6 || variant == 7) ? 256 * 400 - 129 : 256 * 400;
    std::vector<float> h((size_t)P * 48);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)((i * 7) % 13) - 6.0f;
    float *d_rows, *d_out;
    cudaMalloc(&d_rows, h.size() * 4);
    cudaMalloc(&d_out, (size_t)P * 4);
    cudaMemcpy(d_rows, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 128 * 100 * 4;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float* d_out2;
    cudaMalloc(&d_out2, h.size() * 4);
    cudaMemcpy(d_out2, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    probe_kernel<<<(P + 255) / 256, 256, smem>>>(P, d_rows, d_out, variant, d_out2);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(P);
    cudaMemcpy(o.data(), d_out, (size_t)P * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < P; i++) {
        float want = 0; for (int k = 0; k < 48; k++) want += h[(size_t)i * 48 + k];
        if ((variant == 5 || variant == 7) && ((i >> 1) % 4 == 0)) want = 0;
        if (o[i] != want) bad++;
    }
    if (variant >= 9) {
        std::vector<float> o2(h.size());
        cudaMemcpy(o2.data(), d_out2, h.size() * 4, cudaMemcpyDeviceToHost);
        bad = 0;
        for (size_t i = 0; i < h.size(); i++) if (o2[i] != (variant == 10 ? 3 * h[i] : 2 * h[i])) bad++;
    }
    printf("variant %d: %s, mismatches %ld\n", variant, cudaGetErrorString(e), bad);
    return 0;
}
