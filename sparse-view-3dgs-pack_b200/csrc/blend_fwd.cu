// blend_fwd.cu — per-tile front-to-back alpha blending (forward).  Replaces renderCUDA<NUM_CHANNELS>
// (DGR/cuda_rasterizer/forward.cu:274-397), generalised to C = 1..4 channels.
//
// Bit-exact contract: the power / alpha / test_T chain decides n_contrib and final_T, so it is evaluated with the
// exact operation order of the reference's sm_100a build and the same libdevice expf (this file
// must never be compiled with --use_fast_math).  Only the thread->pixel mapping and the data staging differ:
//  - a warp covers an 8x4 pixel patch (better hit coherence than the reference's 16x2 rows);
//  - while a batch is staged, every entry gets an 8-bit mask of the patches it can reach at all (conservative
//    alpha >= 1/255 radius); each warp compacts its own order-preserving list with ballots and only evaluates those
//    entries — a skipped entry would have failed the reference's alpha test on all 32 pixels, so results are unchanged;
//  - colours and 1/depth are staged in shared memory with the geometry, so a hit never touches global memory
//    (the reference gathers features[] / depths[] per hit, forward.cu:372,375).
#include "common.cuh"

namespace lg {

#ifndef FWD_UNROLL
#define FWD_UNROLL 4
#endif
#ifndef FWD_MIN_BLOCKS
#define FWD_MIN_BLOCKS 5
#endif

#ifndef FWD_CP_ASYNC
#define FWD_CP_ASYNC 0
#endif
#ifndef BLEND_BATCH
#define BLEND_BATCH 512  // list entries staged per round (a multiple of the 256 threads)
#endif

template <int C>
__global__ void __launch_bounds__(LG_TILE_PIX, FWD_MIN_BLOCKS) blend_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float2* __restrict__ means2D, const float* __restrict__ features, const float4* __restrict__ conic_opacity,
    const float* __restrict__ depths, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ out_invdepth,
    const uint32_t* __restrict__ tile_order, uint32_t* __restrict__ tile_neff) {
    // one staged entry = three float4: (mean.x, mean.y, -, 1/depth) (conic a, b, c, opacity) (colours, C <= 4), read
    // as warp-wide broadcasts from a single base address
    __shared__ float4 s_ent[BLEND_BATCH * 3];
    __shared__ uint8_t s_mask[BLEND_BATCH];                   // per staged entry: which of the 8 patches it can touch
    __shared__ lg_slot_t s_list[LG_TILE_PIX / 32][BLEND_BATCH]; // per warp: compacted slots of the entries it must evaluate
    __shared__ uint32_t s_neff;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = tile_order[blockIdx.x];  // heaviest tiles first
    const uint32_t tile_x = tile % (uint32_t)grid_x, tile_y = tile / (uint32_t)grid_x;
    if (tid == 0) s_neff = 0;
    const uint32_t pix_x = tile_x * LG_TILE_X + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t pix_y = tile_y * LG_TILE_Y + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pix_y + pix_x;
    const float pixf_x = (float)pix_x, pixf_y = (float)pix_y;
    const float tile_x0 = (float)(tile_x * LG_TILE_X), tile_y0 = (float)(tile_y * LG_TILE_Y);

    const uint2 range = ranges[tile];
    const int rounds = (int)((range.y - range.x + BLEND_BATCH - 1) / BLEND_BATCH);
    int to_do = (int)(range.y - range.x);

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.0f;
    float acc_invd = 0.0f;

    for (int i = 0; i < rounds; i++, to_do -= BLEND_BATCH) {
        if (__syncthreads_count(done) == LG_TILE_PIX) break;
        // ---- stage one batch of the tile's depth-sorted list, with the per-patch reach mask of every entry
#pragma unroll
        for (int u = 0; u < BLEND_BATCH / LG_TILE_PIX; u++) {
            const unsigned slot = u * LG_TILE_PIX + tid;
            const uint32_t progress = (uint32_t)i * BLEND_BATCH + slot;
            unsigned mask = 0;
            if (range.x + progress < range.y) {
                const uint32_t id = point_list[range.x + progress];
                const float2 m = means2D[id];
                const float4 co = conic_opacity[id];
                mask = lg_patch_mask(m.x, m.y, co, tile_x0, tile_y0);
                float fv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int c = 0; c < C; c++) fv[c] = features[(size_t)id * C + c];
                s_ent[slot * 3 + 0] = make_float4(m.x, m.y, 0.0f, F_RCP(depths[id]));
                s_ent[slot * 3 + 1] = co;
                s_ent[slot * 3 + 2] = make_float4(fv[0], fv[1], fv[2], fv[3]);
            }
            s_mask[slot] = (uint8_t)mask;
        }
        __syncthreads();
        const int batch = min(BLEND_BATCH, to_do);
        // ---- each warp keeps only the entries that can reach its 8x4 patch (order preserved)
        const int cnt = lg_compact_patch_list(s_mask, s_list[warp], warp, lane, 0, batch);
        const uint32_t batch_base = (uint32_t)i * BLEND_BATCH;
        // Two list entries per trip: their power / exp / alpha chains are independent (only T couples the entries of
        // a pixel), so evaluating both before applying either doubles the instruction-level parallelism of the loop
        // and halves its overhead.  The decisions are the reference's, in the reference's order (forward.cu:349-381).
        for (int k = 0; !done && k < cnt; k += FWD_UNROLL) {
            int js[FWD_UNROLL];
            float4 xys[FWD_UNROLL];
            float alphas[FWD_UNROLL];
            bool pass[FWD_UNROLL];
#pragma unroll
            for (int u = 0; u < FWD_UNROLL; u++) {
                const bool has = k + u < cnt;
                const int j = s_list[warp][has ? k + u : k];
                const float4 xy = s_ent[j * 3 + 0];
                const float4 co = s_ent[j * 3 + 1];
                const float dx = F_SUB(xy.x, pixf_x), dy = F_SUB(xy.y, pixf_y);
                // power = -0.5f * (a*dx*dx + c*dy*dy) - b*dx*dy, reference contraction order
                const float q = F_FMA(dx, F_MUL(dx, co.x), F_MUL(dy, F_MUL(dy, co.z)));
                const float power = F_FMA(q, -0.5f, -F_MUL(dy, F_MUL(dx, co.y)));  // reference SASS: FFMA(q, -0.5, -m)
                const float alpha = fminf(F_MUL(co.w, expf(power)), 0.99f);
                js[u] = j;
                xys[u] = xy;
                alphas[u] = alpha;
                pass[u] = has && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            }
#pragma unroll
            for (int u = 0; u < FWD_UNROLL; u++) {
                if (pass[u] && !done) {
                    const float alpha = alphas[u];
                    const float test_T = F_MUL(T, F_SUB(1.0f, alpha));
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
                        const float4 f = s_ent[js[u] * 3 + 2];
                        const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                        for (int c = 0; c < C; c++) acc[c] = F_FMA(T, F_MUL(alpha, fv[c]), acc[c]);
                        acc_invd = F_FMA(T, F_MUL(alpha, xys[u].w), acc_invd);
                        T = test_T;
                        last_contributor = batch_base + (uint32_t)js[u] + 1u;  // 1-based list position (forward.cu:345,381)
                    }
                }
            }
        }
    }

    // the backward's launch order key: how deep into the list this tile's pixels reached
    const uint32_t warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
    __syncthreads();  // s_neff initialised (every thread passes here: the loop above has no early return)
    if (lane == 0 && warp_max) atomicMax(&s_neff, warp_max);
    __syncthreads();
    if (tid == 0) tile_neff[tile] = s_neff;
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
#pragma unroll
        for (int c = 0; c < C; c++) out_color[(size_t)c * H * W + pix_id] = F_FMA(T, bg_color[c], acc[c]);
        if (out_invdepth) out_invdepth[pix_id] = acc_invd;
    }
}

#if FWD_CP_ASYNC
// EXPERIMENT (north_star: "Gaussians staged into shared memory by TMA/cp.async in batches"): the same kernel with the
// gathers of batch i+1 issued as cp.async (LDGSTS) into the other half of a double buffer while batch i is blended.
// The per-entry reach mask and 1/depth need the staged values, so every thread post-processes its own slot after
// cp.async.wait_group.  Results are identical; see DESIGN.md §4.3 for the measurement.
#ifndef FWD_CPA_BATCH
#define FWD_CPA_BATCH 256
#endif
__device__ __forceinline__ void cpa4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cpa8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cpa16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}

template <int C>
__global__ void __launch_bounds__(LG_TILE_PIX, FWD_MIN_BLOCKS) blend_forward_cpa_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float2* __restrict__ means2D, const float* __restrict__ features, const float4* __restrict__ conic_opacity,
    const float* __restrict__ depths, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
    const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ out_invdepth,
    const uint32_t* __restrict__ tile_order, uint32_t* __restrict__ tile_neff) {
    constexpr int NB = FWD_CPA_BATCH, PER = NB / LG_TILE_PIX;
    extern __shared__ __align__(16) unsigned char cpa_smem[];
    float4* s_ent = reinterpret_cast<float4*>(cpa_smem);                       // [2][NB * 3]
    lg_slot_t* s_list_all = reinterpret_cast<lg_slot_t*>(s_ent + 2 * NB * 3);  // [8][NB]
    uint8_t* s_mask = reinterpret_cast<uint8_t*>(s_list_all + (LG_TILE_PIX / 32) * NB);
    __shared__ uint32_t s_neff;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = tile_order[blockIdx.x];
    const uint32_t tile_x = tile % (uint32_t)grid_x, tile_y = tile / (uint32_t)grid_x;
    if (tid == 0) s_neff = 0;
    const uint32_t pix_x = tile_x * LG_TILE_X + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t pix_y = tile_y * LG_TILE_Y + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pix_y + pix_x;
    const float pixf_x = (float)pix_x, pixf_y = (float)pix_y;
    const float tile_x0 = (float)(tile_x * LG_TILE_X), tile_y0 = (float)(tile_y * LG_TILE_Y);
    lg_slot_t* s_list = s_list_all + warp * NB;

    const uint2 range = ranges[tile];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + NB - 1) / NB;

    auto issue = [&](int r) {
        float4* buf = s_ent + (r & 1) * NB * 3;
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const unsigned slot = u * LG_TILE_PIX + tid;
            const uint32_t progress = (uint32_t)r * NB + slot;
            if (range.x + progress < range.y) {
                const uint32_t id = point_list[range.x + progress];
                float* e0 = reinterpret_cast<float*>(buf + slot * 3);
                cpa8(e0, means2D + id);
                cpa4(e0 + 3, depths + id);
                cpa16(buf + slot * 3 + 1, conic_opacity + id);
#pragma unroll
                for (int c = 0; c < C; c++) cpa4(e0 + 8 + c, features + (size_t)id * C + c);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; c++) acc[c] = 0.0f;
    float acc_invd = 0.0f;

    if (rounds > 0) issue(0);
    for (int i = 0; i < rounds; i++) {
        if (__syncthreads_count(done) == LG_TILE_PIX) break;
        const bool more = i + 1 < rounds;
        if (more) issue(i + 1);
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        float4* buf = s_ent + (i & 1) * NB * 3;
        const int batch = min(NB, total - i * NB);
        // every thread finishes the slots it copied itself (its own cp.async results are visible to it)
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int slot = u * LG_TILE_PIX + (int)tid;
            unsigned mask = 0;
            if (slot < batch) {
                const float4 e0 = buf[slot * 3 + 0];
                const float4 co = buf[slot * 3 + 1];
                mask = lg_patch_mask(e0.x, e0.y, co, tile_x0, tile_y0);
                reinterpret_cast<float*>(buf + slot * 3)[3] = F_RCP(e0.w);
            }
            s_mask[slot] = (uint8_t)mask;
        }
        __syncthreads();
        const int cnt = lg_compact_patch_list(s_mask, s_list, warp, lane, 0, batch);
        const uint32_t batch_base = (uint32_t)i * NB;
        for (int k = 0; !done && k < cnt; k += FWD_UNROLL) {
            int js[FWD_UNROLL];
            float4 xys[FWD_UNROLL];
            float alphas[FWD_UNROLL];
            bool pass[FWD_UNROLL];
#pragma unroll
            for (int u = 0; u < FWD_UNROLL; u++) {
                const bool has = k + u < cnt;
                const int j = s_list[has ? k + u : k];
                const float4 xy = buf[j * 3 + 0];
                const float4 co = buf[j * 3 + 1];
                const float dx = F_SUB(xy.x, pixf_x), dy = F_SUB(xy.y, pixf_y);
                const float q = F_FMA(dx, F_MUL(dx, co.x), F_MUL(dy, F_MUL(dy, co.z)));
                const float power = F_FMA(q, -0.5f, -F_MUL(dy, F_MUL(dx, co.y)));
                const float alpha = fminf(F_MUL(co.w, expf(power)), 0.99f);
                js[u] = j;
                xys[u] = xy;
                alphas[u] = alpha;
                pass[u] = has && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            }
#pragma unroll
            for (int u = 0; u < FWD_UNROLL; u++) {
                if (pass[u] && !done) {
                    const float alpha = alphas[u];
                    const float test_T = F_MUL(T, F_SUB(1.0f, alpha));
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
                        const float4 f = buf[js[u] * 3 + 2];
                        const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                        for (int c = 0; c < C; c++) acc[c] = F_FMA(T, F_MUL(alpha, fv[c]), acc[c]);
                        acc_invd = F_FMA(T, F_MUL(alpha, xys[u].w), acc_invd);
                        T = test_T;
                        last_contributor = batch_base + (uint32_t)js[u] + 1u;
                    }
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");  // a prefetched batch may still be in flight after an early exit

    const uint32_t warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
    __syncthreads();
    if (lane == 0 && warp_max) atomicMax(&s_neff, warp_max);
    __syncthreads();
    if (tid == 0) tile_neff[tile] = s_neff;
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
#pragma unroll
        for (int c = 0; c < C; c++) out_color[(size_t)c * H * W + pix_id] = F_FMA(T, bg_color[c], acc[c]);
        if (out_invdepth) out_invdepth[pix_id] = acc_invd;
    }
}
#endif  // FWD_CP_ASYNC

// Instrumentation: replays the blend decisions and counts, over all pixels, the list entries evaluated before early
// termination (N_eval) and the (pixel, Gaussian) pairs that pass all three tests (N_hit) — the work terms of the
// blend roofline.  Not part of the rendering path.
__global__ void __launch_bounds__(LG_TILE_PIX) blend_count_kernel(const uint2* __restrict__ ranges,
                                                                  const uint32_t* __restrict__ point_list, int W, int H,
                                                                  int grid_x, const float2* __restrict__ means2D,
                                                                  const float4* __restrict__ conic_opacity,
                                                                  unsigned long long* __restrict__ counts) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t pix_x = blockIdx.x * LG_TILE_X + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t pix_y = blockIdx.y * LG_TILE_Y + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const float pixf_x = (float)pix_x, pixf_y = (float)pix_y;
    const uint2 range = ranges[blockIdx.y * (uint32_t)grid_x + blockIdx.x];
    unsigned long long n_eval = 0, n_hit = 0;
    if (inside) {
        float T = 1.0f;
        for (uint32_t e = range.x; e < range.y; e++) {
            n_eval++;
            const uint32_t id = point_list[e];
            const float2 xy = means2D[id];
            const float4 co = conic_opacity[id];
            const float dx = F_SUB(xy.x, pixf_x), dy = F_SUB(xy.y, pixf_y);
            const float q = F_FMA(dx, F_MUL(dx, co.x), F_MUL(dy, F_MUL(dy, co.z)));
            const float power = F_FMA(q, -0.5f, -F_MUL(dy, F_MUL(dx, co.y)));
            if (power > 0.0f) continue;
            const float alpha = fminf(F_MUL(co.w, expf(power)), 0.99f);
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = F_MUL(T, F_SUB(1.0f, alpha));
            if (test_T < 0.0001f) break;
            T = test_T;
            n_hit++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_eval += __shfl_xor_sync(0xffffffffu, n_eval, o);
        n_hit += __shfl_xor_sync(0xffffffffu, n_hit, o);
    }
    if (lane == 0) {
        atomicAdd(&counts[0], n_eval);
        atomicAdd(&counts[1], n_hit);
    }
}

int launch_blend_count(int W, int H, const GeometryState& g, const BinningState& b, const ImageState& img,
                       unsigned long long* counts, cudaStream_t stream) {
    LG_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), stream));
    const dim3 grid(num_tiles_x(W), num_tiles_y(H), 1);
    blend_count_kernel<<<grid, LG_TILE_PIX, 0, stream>>>(img.ranges, b.point_list, W, H, (int)grid.x, g.means2D,
                                                         g.conic_opacity, counts);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

int launch_blend_forward(int C, int W, int H, const GeometryState& g, const BinningState& b, ImageState& img,
                         const float* features, const float* background, float* out_color, float* out_invdepth,
                         bool debug, cudaStream_t stream) {
    const int gx = num_tiles_x(W);
    const dim3 grid(gx * num_tiles_y(H), 1, 1), block(LG_TILE_PIX, 1, 1);
#define LG_LAUNCH_FWD(CH)                                                                                          \
    blend_forward_kernel<CH><<<grid, block, 0, stream>>>(img.ranges, b.point_list, W, H, gx, g.means2D,              \
                                                         features, g.conic_opacity, g.depths, img.accum_alpha,       \
                                                         img.n_contrib, background, out_color, out_invdepth,         \
                                                         img.tile_order, img.tile_neff)
#if FWD_CP_ASYNC
    {
        const size_t smem = (size_t)2 * FWD_CPA_BATCH * 3 * sizeof(float4) + (LG_TILE_PIX / 32) * FWD_CPA_BATCH * sizeof(lg_slot_t) + FWD_CPA_BATCH;
        if (C == 3) {
            LG_CUDA(cudaFuncSetAttribute(blend_forward_cpa_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            blend_forward_cpa_kernel<3><<<grid, block, smem, stream>>>(img.ranges, b.point_list, W, H, gx, g.means2D, features,
                                                                       g.conic_opacity, g.depths, img.accum_alpha, img.n_contrib,
                                                                       background, out_color, out_invdepth, img.tile_order,
                                                                       img.tile_neff);
            LG_LAUNCH_CHECK(debug, stream);
            return LG_OK;
        }
    }
#endif
    switch (C) {
        case 1: LG_LAUNCH_FWD(1); break;
        case 2: LG_LAUNCH_FWD(2); break;
        case 3: LG_LAUNCH_FWD(3); break;
        case 4: LG_LAUNCH_FWD(4); break;
        default: set_error("blend forward: unsupported channel count %d", C); return LG_ERR_UNSUPPORTED;
    }
#undef LG_LAUNCH_FWD
    LG_LAUNCH_CHECK(debug, stream);
    return LG_OK;
}

}  // namespace lg
