// blend_fwd.cu — per-tile front-to-back alpha blending (forward).  Replaces renderCUDA<NUM_CHANNELS>
// (DGR/cuda_rasterizer/forward.cu:274-397), generalised to C = 1..4 channels.
//
// Bit-exact contract: the power / alpha / test_T chain decides n_contrib and final_T, so it is evaluated with the
// exact operation order of the reference's sm_100a build and the same libdevice expf (this file
// must never be compiled with --use_fast_math).  Only the thread->pixel mapping and the data staging differ:
//  - a warp covers an 8x4 pixel patch (better hit coherence than the reference's 16x2 rows);
//  - while a batch is staged, every entry gets a 16-bit mask of the 4x4 sub-patches it can reach at all (exact minimum
//    of the quadratic form over each box, common.cuh); neighbouring bits are ORed into the 8 patch bits and each warp
//    compacts its own order-preserving list with ballots and only evaluates those entries — a skipped entry would have
//    failed the reference's alpha test on all 32 pixels, so results are unchanged.  The 16-bit masks stay in global
//    memory for the backward pass, which walks one list per half-warp;
//  - colours and 1/depth are staged in shared memory with the geometry, so a hit never touches global memory
//    (the reference gathers features[] / depths[] per hit, forward.cu:372,375).
#include "common.cuh"

namespace lg {

#ifndef FWD_UNROLL
#define FWD_UNROLL 4
#endif
#ifndef FWD_MIN_BLOCKS
#define FWD_MIN_BLOCKS 4  // 64 registers: 5 blocks (48 registers) spills in the staging loop since the sub-patch masks: 206 vs 192 us
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
    const uint32_t* __restrict__ tile_order, uint32_t* __restrict__ tile_neff, uint32_t* __restrict__ counters,
    uint32_t capacity, uint32_t* __restrict__ tile_order_bwd, uint16_t* __restrict__ entry_masks) {
    // The list was only built if it fits the binning buffer the caller sized speculatively (abi.cu); otherwise this
    // launch is void and the host queues the tail of the forward again.
    if (counters[1] > capacity) return;
    // one staged entry = three float4: (mean.x, mean.y, -, 1/depth) (conic a, b, c, opacity) (colours, C <= 4), read
    // as warp-wide broadcasts from a single base address; entry BLEND_BATCH is a sentinel that never passes the alpha
    // test (opacity 0) and pads the per-warp lists to a multiple of FWD_UNROLL
    constexpr int ENT = 48;  // bytes per staged entry
    __shared__ float4 s_ent[(BLEND_BATCH + 1) * 3];
    // per 32 staged entries and patch: which of them can touch the patch at all (ballots of the staging warps)
    __shared__ uint32_t s_reach[BLEND_BATCH / 32][LG_TILE_PIX / 32];
    // per warp: byte offsets into s_ent of the entries it must evaluate, FWD_UNROLL of them fetched by one load
    __shared__ __align__(8) lg_slot_t s_list[LG_TILE_PIX / 32][BLEND_BATCH + 4];
    __shared__ uint32_t s_neff;
    static_assert(FWD_UNROLL == 4, "the list is read four offsets at a time");

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = tile_order[blockIdx.x];  // heaviest tiles first
    const uint32_t tile_x = tile % (uint32_t)grid_x, tile_y = tile / (uint32_t)grid_x;
    if (tid == 0) {
        s_neff = 0;
        s_ent[BLEND_BATCH * 3 + 0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        s_ent[BLEND_BATCH * 3 + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // opacity 0: alpha = 0 < 1/255
        s_ent[BLEND_BATCH * 3 + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
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
    const char* const ent_base = reinterpret_cast<const char*>(s_ent);

    for (int i = 0; i < rounds; i++, to_do -= BLEND_BATCH) {
        if (__syncthreads_count(done) == LG_TILE_PIX) break;
        // ---- stage one batch of the tile's depth-sorted list, with the per-patch reach mask of every entry
#pragma unroll
        for (int u = 0; u < BLEND_BATCH / LG_TILE_PIX; u++) {
            const unsigned slot = u * LG_TILE_PIX + tid;
            const uint32_t progress = (uint32_t)i * BLEND_BATCH + slot;
            unsigned mask = 0, mask16 = 0;
            if (range.x + progress < range.y) {
                const uint32_t id = point_list[range.x + progress];
                const float2 m = means2D[id];
                const float4 co = conic_opacity[id];
                mask16 = lg_subpatch_mask16(m.x, m.y, co, tile_x0, tile_y0);
                // this kernel walks one list per warp (8x4 patch = two sub-patches side by side: bits 2j, 2j + 1)
                unsigned t = (mask16 | (mask16 >> 1)) & 0x5555u;
                t = (t | (t >> 1)) & 0x3333u;
                t = (t | (t >> 2)) & 0x0f0fu;
                mask = (t | (t >> 4)) & 0x00ffu;
                float fv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int c = 0; c < C; c++) fv[c] = features[(size_t)id * C + c];
                s_ent[slot * 3 + 0] = make_float4(m.x, m.y, 0.0f, F_RCP(depths[id]));
                s_ent[slot * 3 + 1] = co;
                s_ent[slot * 3 + 2] = make_float4(fv[0], fv[1], fv[2], fv[3]);
            }
            // the staging warp votes once per patch; lane b keeps the word of patch b.  The 16-bit sub-patch mask goes to
            // global memory for the backward pass, which stages the same entries (it never goes past a batch staged
            // here) and walks one list per half-warp
            if (range.x + progress < range.y) entry_masks[range.x + progress] = (uint16_t)mask16;
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < LG_TILE_PIX / 32; b++) {
                const uint32_t bal = __ballot_sync(0xffffffffu, (mask >> b) & 1u);
                word = lane == (unsigned)b ? bal : word;
            }
            if (lane < LG_TILE_PIX / 32) s_reach[slot >> 5][lane] = word;
        }
        __syncthreads();
        // ---- each warp keeps only the entries that can reach its 8x4 patch (order preserved)
        // (words of chunks past the end of the list are zero: the staging loop covers all BLEND_BATCH slots)
        int cnt = 0;
        {
            const uint32_t list_addr = (uint32_t)__cvta_generic_to_shared(&s_list[warp][0]);
            const unsigned lt = (1u << lane) - 1u;
            const uint32_t my_off = lane * ENT;
#pragma unroll
            for (int c = 0; c < BLEND_BATCH / 32; c++) {
                const uint32_t bal = s_reach[c][warp];
                if ((bal >> lane) & 1u)
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr + 2u * (uint32_t)(cnt + __popc(bal & lt))),
                                 "h"((unsigned short)(my_off + c * 32 * ENT)) : "memory");
                cnt += __popc(bal);
            }
            if (lane < 3)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr + 2u * (uint32_t)(cnt + (int)lane)),
                             "h"((unsigned short)(BLEND_BATCH * ENT)) : "memory");
            __syncwarp();
            cnt = (int)__reduce_max_sync(0xffffffffu, (unsigned)cnt);  // same in every lane; tells the compiler so
        }
        const uint32_t batch_base = (uint32_t)i * BLEND_BATCH;
        // Four list entries per trip: their power / exp / alpha chains are independent (only T couples the entries of
        // a pixel), so all are evaluated before any is applied.  The decisions are the reference's, in the reference's
        // order (forward.cu:349-381), applied without branches: an entry that does not pass, or a pixel that is done,
        // blends with alpha = 0, which leaves T and the accumulators bit-for-bit unchanged (T * (1 - 0) = T,
        // fma(T, 0 * f, acc) = acc), and since T never drops below 1e-4 such a step cannot trigger the stop test.
        int last_off = -1;  // byte offset of the last entry of this batch that contributed to this pixel
        for (int k = 0; k < cnt; k += FWD_UNROLL) {
            const uint2 offs = *reinterpret_cast<const uint2*>(&s_list[warp][k]);
            const int offv[FWD_UNROLL] = {(int)(offs.x & 0xffffu), (int)(offs.x >> 16), (int)(offs.y & 0xffffu),
                                          (int)(offs.y >> 16)};
            float invds[FWD_UNROLL], alphas[FWD_UNROLL];
            bool pass[FWD_UNROLL];
#pragma unroll
            for (int u = 0; u < FWD_UNROLL; u++) {
                const float4 xy = *reinterpret_cast<const float4*>(ent_base + offv[u]);
                const float4 co = *reinterpret_cast<const float4*>(ent_base + offv[u] + 16);
                const float dx = F_SUB(xy.x, pixf_x), dy = F_SUB(xy.y, pixf_y);
                // power = -0.5f * (a*dx*dx + c*dy*dy) - b*dx*dy, reference contraction order
                const float q = F_FMA(dx, F_MUL(dx, co.x), F_MUL(dy, F_MUL(dy, co.z)));
                const float power = F_FMA(q, -0.5f, -F_MUL(dy, F_MUL(dx, co.y)));  // reference SASS: FFMA(q, -0.5, -m)
                const float alpha = fminf(F_MUL(co.w, expf(power)), 0.99f);
                invds[u] = xy.w;
                alphas[u] = alpha;
                pass[u] = !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            }
#pragma unroll
            for (int u = 0; u < FWD_UNROLL; u++) {
                const bool live = pass[u] && !done;
                const float a = live ? alphas[u] : 0.0f;
                const float test_T = F_MUL(T, F_SUB(1.0f, a));
                const bool stop = test_T < 0.0001f;  // only a live step can get here: T >= 1e-4 at all times
                done = done || stop;
                const float a2 = stop ? 0.0f : a;
                const float4 f = *reinterpret_cast<const float4*>(ent_base + offv[u] + 32);
                const float fv[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int c = 0; c < C; c++) acc[c] = F_FMA(T, F_MUL(a2, fv[c]), acc[c]);
                acc_invd = F_FMA(T, F_MUL(a2, invds[u]), acc_invd);
                T = stop ? T : test_T;
                last_off = (live && !stop) ? offv[u] : last_off;  // 1-based list position below (forward.cu:345,381)
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
        if (last_off >= 0) last_contributor = batch_base + (uint32_t)last_off / (uint32_t)ENT + 1u;
    }

    // the backward's launch order key: how deep into the list this tile's pixels reached
    const uint32_t warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
    __syncthreads();  // s_neff initialised (every thread passes here: the loop above has no early return)
    if (lane == 0 && warp_max) atomicMax(&s_neff, warp_max);
    __syncthreads();
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
#pragma unroll
        for (int c = 0; c < C; c++) out_color[(size_t)c * H * W + pix_id] = F_FMA(T, bg_color[c], acc[c]);
        if (out_invdepth) out_invdepth[pix_id] = acc_invd;
    }
    // The block that finishes last turns tile_neff into the backward's launch order (deepest tiles first), which
    // used to be a kernel launch of its own at the head of the backward pass.
    if (tid == 0) {
        tile_neff[tile] = s_neff;
        __threadfence();
        s_neff = atomicAdd(&counters[3], 1u);
    }
    __syncthreads();
    if (s_neff == gridDim.x - 1) {
        __threadfence();
        uint32_t* scratch = reinterpret_cast<uint32_t*>(s_ent);  // 1024 + 33 words of the staging buffer
        lg_bucket_order<LG_TILE_PIX>((int)gridDim.x, tile_neff, tile_order_bwd, scratch, scratch + 1024);
    }
}

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

int launch_blend_forward(int C, int W, int H, int capacity, const GeometryState& g, const BinningState& b, ImageState& img,
                         const float* features, const float* background, float* out_color, float* out_invdepth,
                         bool debug, cudaStream_t stream) {
    const int gx = num_tiles_x(W);
    const dim3 grid(gx * num_tiles_y(H), 1, 1), block(LG_TILE_PIX, 1, 1);
#define LG_LAUNCH_FWD(CH)                                                                                          \
    blend_forward_kernel<CH><<<grid, block, 0, stream>>>(img.ranges, b.point_list, W, H, gx, g.means2D,              \
                                                         features, g.conic_opacity, g.depths, img.accum_alpha,       \
                                                         img.n_contrib, background, out_color, out_invdepth,         \
                                                         img.tile_order, img.tile_neff, img.counters,                \
                                                         (uint32_t)capacity, img.tile_order_bwd,                     \
                                                         reinterpret_cast<uint16_t*>(b.pairs))
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
