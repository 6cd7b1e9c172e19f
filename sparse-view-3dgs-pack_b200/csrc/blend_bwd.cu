// blend_bwd.cu — per-tile back-to-front replay of the alpha blend (backward).  Replaces renderCUDA<C> in
// DGR/cuda_rasterizer/backward.cu:452-638.
//
// The reference issues 9-10 global atomicAdds per (pixel, Gaussian) hit.  Here a half-warp owns a 4x4 pixel sub-patch
// and walks its own compacted list of the entries that can reach it (the 16-bit sub-patch masks the forward left per
// list entry); the two halves of a warp take a trip together, two entries each.  The 16 pixels of a half reduce the
// 7+C partial gradients of both entries with a halving butterfly (19 shuffles per trip for 2 x 2 x 9 values), after
// which every lane owns at most two totals and sends them with red.global.add into a packed 12-float per-Gaussian
// record.  Trips in which no pixel hits skip everything after a ballot, and list entries behind the last contributor of
// every pixel of the tile are never even staged.  (One list per warp over the 8x4 patch, round 2 until its last day:
// 21 % more trips for the same hits — 336 -> 315 us.  The forward keeps the per-warp lists: with two addresses per
// shared-memory load it becomes bound by the load pipe and loses 14 %.)
#include "common.cuh"

namespace lg {

#ifndef BWD_UNROLL
#define BWD_UNROLL 2
#endif
#ifndef BWD_MIN_BLOCKS
#define BWD_MIN_BLOCKS 4
#endif

#ifndef BWD_BATCH
#define BWD_BATCH 256  // list entries staged per round (a multiple of the 256 threads; 512: 315 vs 307 us)
#endif
#define LG_REC 12  // floats per packed gradient record: mean2D.xy, conic.xyw, opacity, invdepth, colour[C], pad

// One stage (lane ^ OFF) and all later ones of a halving butterfly that sums N per-lane values over a group of 2 * OFF
// lanes: the two partner lanes split the n live values between them (the upper lane keeps the upper half), so stage
// sizes are ceil(n/2): from 9 values at OFF = 4: 5, 3, 2 shuffles, after which a lane holds at most two totals, values
// base .. base + cnt - 1 of the N (cnt <= 0: none).  Sizes are template parameters so that every array index is a
// compile-time constant (a run-time `n` put the values into local memory).
template <int NA, int N, int OFF>
__device__ __forceinline__ void warp_multi_reduce_stage(float (&v)[NA], unsigned lane, int& base, int& cnt) {
    if constexpr (OFF >= 1) {
        const bool upper = (lane & OFF) != 0;
        if constexpr (N > 1) {
            constexpr int h = (N + 1) / 2;
#pragma unroll
            for (int k = 0; k < h; k++) {
                const float lo = v[k];
                const float hi = (k + h < N) ? v[k + h] : 0.0f;
                const float send = upper ? lo : hi;
                const float keep = upper ? hi : lo;
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
            }
            if (upper) { base += h; cnt -= h; } else { cnt = min(cnt, h); }
            warp_multi_reduce_stage<NA, h, OFF / 2>(v, lane, base, cnt);
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
            if (upper) cnt = 0;  // both partners now hold the same total: the lower lane owns it
            warp_multi_reduce_stage<NA, 1, OFF / 2>(v, lane, base, cnt);
        }
    }
}

// Per-Gaussian record accumulated here (LG_REC floats, consumed by preprocess_backward_kernel):
//   [0] sum w*dx   [1] sum w*dy   [2] sum w*dx^2   [3] sum w*dx*dy   [4] sum w*dy^2   [5] sum w
//   [6] sum alpha*T*dL/dinvdepth_pix   [7..7+C) sum alpha*T*dL/dpix_c
// with w = G * dL/dalpha and (dx, dy) = mean2D - pixel.  The reference's dL/dmean2D, dL/dconic and dL/dopacity
// (backward.cu:598-632) are linear in these moments with per-Gaussian coefficients (conic, opacity), so the
// coefficients are applied once per Gaussian in the per-Gaussian kernel instead of once per (pixel, Gaussian) hit.
//
// Pixel state.  The reference carries accum_rec[c], last_color[c], last_alpha per pixel and forms
//   dL/dalpha = T * sum_c (col_c - accum_rec_c) * g_c + bg_term / (1 - alpha)          (backward.cu:560-589, g = dL/dpixel)
// Everything in it is linear in g, so the pixel only needs the scalar B = sum_c accum_rec_c * g_c *as the next hit will
// see it*: after a hit with colour dot cg = sum_c col_c g_c,  B <- alpha * cg + (1 - alpha) * B.  Two registers of
// state (T, B) instead of 2C + 2; the inverse-depth term is one more "colour" (1/depth, dL/dinvdepth_pix).
//
// No branch around the per-hit arithmetic: lanes whose pixel is not hit run it with alpha = G = 0, which leaves B
// unchanged and makes every partial sum an exact zero (T is kept by a select); the 7+C zero-initialisations and the
// divergent region of the earlier version cost more issue slots than the arithmetic they skipped.
template <int C, bool INVD>
struct BwdPixel {
    float T, B;            // transmittance in front of the next hit; blended colour behind it, dotted with g
    float g[C];            // dL/dpixel
    float g_invd;          // dL/dinvdepth_pix
    float bg_term;         // -T_final * sum_c bg_c g_c
    float x, y;            // pixel centre
};

// One hit's per-lane scalars: everything its 6 + C (+1) partial sums are products of.
struct BwdHit {
    float w, dx, dy, aT;   // G * dL/dalpha, mean2D - pixel, alpha * T
};

// Advances the pixel state (T, B) over one list entry and returns the hit's scalars (all zero weights when the pixel is
// not hit: alpha = G = 0 leaves B unchanged, T is kept by a select).
template <int C, bool INVD>
__device__ __forceinline__ BwdHit bwd_hit_scalars(BwdPixel<C, INVD>& px, const char* ent, bool hit, float dx, float dy,
                                                  float G, float alpha, float invd) {
    const float alpha_e = hit ? alpha : 0.0f;
    const float G_e = hit ? G : 0.0f;
    float rinv;  // 1 - alpha lies in [0.01, 1]: the bare MUFU.RCP needs no range fix-up
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(1.0f - alpha_e));
    const float Tn = px.T * rinv;
    px.T = hit ? Tn : px.T;
    const float aT = alpha_e * px.T;
    const float4 f4 = *reinterpret_cast<const float4*>(ent + 32);
    const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
    float cg = fv[0] * px.g[0];
#pragma unroll
    for (int c = 1; c < C; c++) cg = fmaf(fv[c], px.g[c], cg);
    if (INVD) cg = fmaf(invd, px.g_invd, cg);
    const float d = cg - px.B;
    const float dL_dalpha = fmaf(d, px.T, px.bg_term * rinv);
    px.B = fmaf(alpha_e, d, px.B);
    BwdHit h;
    h.w = G_e * dL_dalpha;
    h.dx = dx;
    h.dy = dy;
    h.aT = aT;
    return h;
}

// the 6 + (INVD) + C partial sums of one (pixel, Gaussian) hit (layout: see the record description above)
template <int C, bool INVD, int NVT>
__device__ __forceinline__ void bwd_hit_products(const BwdPixel<C, INVD>& px, const BwdHit& h, float (&v)[NVT]) {
    constexpr int VC = 6 + (INVD ? 1 : 0);
    const float wx = h.w * h.dx, wy = h.w * h.dy;
    v[0] = wx;
    v[1] = wy;
    v[2] = wx * h.dx;
    v[3] = wx * h.dy;
    v[4] = wy * h.dy;
    v[5] = h.w;
    if (INVD) v[6] = h.aT * px.g_invd;
#pragma unroll
    for (int c = 0; c < C; c++) v[VC + c] = h.aT * px.g[c];
}

template <int C, bool INVD, int NVT>
__device__ __forceinline__ void bwd_hit_values(BwdPixel<C, INVD>& px, const char* ent, bool hit, float dx, float dy,
                                               float G, float alpha, float invd, float (&v)[NVT]) {
    const BwdHit h = bwd_hit_scalars<C, INVD>(px, ent, hit, dx, dy, G, alpha, invd);
    bwd_hit_products<C, INVD, NVT>(px, h, v);
}

template <int C, bool INVD>
__global__ void __launch_bounds__(LG_TILE_PIX, BWD_MIN_BLOCKS) blend_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float* __restrict__ bg_color, const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
    const float* __restrict__ colors, const float* __restrict__ depths, const float* __restrict__ final_Ts,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dinvdepth_pix, float* __restrict__ grad_rec,
    const uint32_t* __restrict__ tile_order, const uint16_t* __restrict__ entry_masks) {
    constexpr int NV = 6 + (INVD ? 1 : 0) + C;  // values reduced per (warp, Gaussian)
    constexpr int ENT = 48;                     // bytes per staged entry
    // one staged entry = three float4: (mean.x, mean.y, Gaussian id, 1/depth) (conic a, b, c, opacity) (colours, C <= 4),
    // read as warp-wide broadcasts from a single base address; entry BWD_BATCH is a sentinel that can never hit
    // (opacity 0), which pads odd-length lists
    __shared__ float4 s_ent[(BWD_BATCH + 1) * 3];
    // per 32 staged entries and patch: which of them can touch the patch at all (ballots of the staging warps over the
    // mask bytes the forward pass left per list entry)
    __shared__ __align__(8) uint32_t s_reach[BWD_BATCH / 32][16];
    // per half-warp (= 4x4 sub-patch, as in the forward): byte offsets into s_ent; the two lists of a warp are padded
    // with the sentinel to a common even length
    __shared__ __align__(4) lg_slot_t s_list[LG_TILE_PIX / 32][2][BWD_BATCH + 2];
    __shared__ uint32_t s_max;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = tile_order[blockIdx.x];  // deepest tiles first
    const uint32_t tile_x = tile % (uint32_t)grid_x, tile_y = tile / (uint32_t)grid_x;
    const unsigned half = lane >> 4;  // lanes 0-15: the left 4x4 sub-patch of the warp's 8x4 patch, 16-31: the right one
    const uint32_t pix_x = tile_x * LG_TILE_X + (warp & 1u) * 8u + half * 4u + (lane & 3u);
    const uint32_t pix_y = tile_y * LG_TILE_Y + (warp >> 1) * 4u + ((lane >> 2) & 3u);
    const unsigned sub0 = (warp >> 1) * 4u + (warp & 1u) * 2u;  // mask bit of the left sub-patch (right: + 1)
    const unsigned half_mask = half ? 0xffff0000u : 0x0000ffffu;
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pix_y + pix_x;
    const uint2 range = ranges[tile];

    const float T_final = inside ? final_Ts[pix_id] : 0.0f;
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0u;

    if (tid == 0) {
        s_max = 0;
        s_ent[BWD_BATCH * 3 + 0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        s_ent[BWD_BATCH * 3 + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // opacity 0: alpha = 0 < 1/255
        s_ent[BWD_BATCH * 3 + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncthreads();
    const uint32_t warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
    const uint32_t left_max = __reduce_max_sync(0xffffffffu, half ? 0u : last_contributor);
    const uint32_t right_max = __reduce_max_sync(0xffffffffu, half ? last_contributor : 0u);
    if (lane == 0 && warp_max) atomicMax(&s_max, warp_max);
    __syncthreads();
    const uint32_t n_eff = s_max;  // entries [0, n_eff) of this tile's list can contribute to some pixel
    if (n_eff == 0) return;

    BwdPixel<C, INVD> px;
    px.T = T_final;
    px.B = 0.0f;
    px.x = (float)pix_x;
    px.y = (float)pix_y;
    float bg_dot_dpixel = 0.0f;
#pragma unroll
    for (int c = 0; c < C; c++) {
        px.g[c] = inside ? dL_dpixels[(size_t)c * H * W + pix_id] : 0.0f;
        bg_dot_dpixel += bg_color[c] * px.g[c];
    }
    px.bg_term = -T_final * bg_dot_dpixel;
    px.g_invd = 0.0f;
    if (INVD) px.g_invd = inside ? dL_dinvdepth_pix[pix_id] : 0.0f;

    const char* const ent_base = reinterpret_cast<const char*>(s_ent);
    const int rounds = (int)((n_eff + BWD_BATCH - 1) / BWD_BATCH);
    for (int i = 0; i < rounds; i++) {
        __syncthreads();
        // ---- stage one batch, back to front, with the per-patch reach mask of every entry
        const uint32_t batch_base = (uint32_t)i * BWD_BATCH;
#pragma unroll
        for (int u = 0; u < BWD_BATCH / LG_TILE_PIX; u++) {
            const unsigned slot = u * LG_TILE_PIX + tid;
            const uint32_t progress = batch_base + slot;
            unsigned mask = 0;
            if (progress < n_eff) {
                const uint32_t pos = range.x + (n_eff - 1u - progress);
                const uint32_t id = point_list[pos];
                mask = entry_masks[pos];
                const float2 m = means2D[id];
                const float4 cq = conic_opacity[id];
                float fv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int c = 0; c < C; c++) fv[c] = colors[(size_t)id * C + c];
                s_ent[slot * 3 + 0] = make_float4(m.x, m.y, __uint_as_float(id), INVD ? 1.0f / depths[id] : 0.0f);
                s_ent[slot * 3 + 1] = cq;
                s_ent[slot * 3 + 2] = make_float4(fv[0], fv[1], fv[2], fv[3]);
            }
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const uint32_t bal = __ballot_sync(0xffffffffu, (mask >> b) & 1u);
                word = lane == (unsigned)b ? bal : word;
            }
            if (lane < 16) s_reach[slot >> 5][lane] = word;
        }
        __syncthreads();
        const int batch = (int)min((uint32_t)BWD_BATCH, n_eff - batch_base);
        // slots whose list position is behind this warp's last contributor are dropped with the unreachable ones:
        // position rel = n_eff-1-(batch_base+j) < warp_max  <=>  j >= n_eff - warp_max - batch_base
        const int first = (int)max((long long)n_eff - (long long)warp_max - (long long)batch_base, 0ll);
        if (first >= batch) continue;
        // compacted lists of the two sub-patches' reachable entries as byte offsets into s_ent (all 32 lanes build both:
        // lane = entry within a chunk of 32), padded with the sentinel to a common even length.  A sub-patch also drops
        // the slots behind its own last contributor.
        // (words of chunks past the end of the list are zero: the staging loop covers all BWD_BATCH slots)
        int cnt = 0;
        {
            const int first0 = (int)max((long long)n_eff - (long long)left_max - (long long)batch_base, 0ll);
            const int first1 = (int)max((long long)n_eff - (long long)right_max - (long long)batch_base, 0ll);
            const uint32_t list_addr0 = (uint32_t)__cvta_generic_to_shared(&s_list[warp][0][0]);
            const uint32_t list_addr1 = (uint32_t)__cvta_generic_to_shared(&s_list[warp][1][0]);
            const unsigned lt = (1u << lane) - 1u;
            const uint32_t my_off = lane * ENT;
            int cnt0 = 0, cnt1 = 0;
#pragma unroll
            for (int c = 0; c < BWD_BATCH / 32; c++) {
                const uint2 reach = *reinterpret_cast<const uint2*>(&s_reach[c][sub0]);
                const int drop0 = min(max(first0 - c * 32, 0), 32), drop1 = min(max(first1 - c * 32, 0), 32);
                const uint32_t bal0 = reach.x & (uint32_t)(0xffffffffull << drop0);
                const uint32_t bal1 = reach.y & (uint32_t)(0xffffffffull << drop1);
                const unsigned short off = (unsigned short)(my_off + c * 32 * ENT);
                if ((bal0 >> lane) & 1u)
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr0 + 2u * (uint32_t)(cnt0 + __popc(bal0 & lt))),
                                 "h"(off) : "memory");
                if ((bal1 >> lane) & 1u)
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr1 + 2u * (uint32_t)(cnt1 + __popc(bal1 & lt))),
                                 "h"(off) : "memory");
                cnt0 += __popc(bal0);
                cnt1 += __popc(bal1);
            }
            cnt = (max(cnt0, cnt1) + 1) & ~1;
            for (int j = cnt0 + (int)lane; j < cnt; j += 32)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr0 + 2u * (uint32_t)j),
                             "h"((unsigned short)(BWD_BATCH * ENT)) : "memory");
            for (int j = cnt1 + (int)lane; j < cnt; j += 32)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr1 + 2u * (uint32_t)j),
                             "h"((unsigned short)(BWD_BATCH * ENT)) : "memory");
            __syncwarp();
            cnt = (int)__reduce_max_sync(0xffffffffu, (unsigned)cnt);  // same in every lane; tells the compiler so
        }
        // this pixel is behind entry j of the batch  <=>  rel = n_eff-1-(batch_base+j) < last_contributor
        //                                           <=>  j*ENT > ENT * (n_eff-1-batch_base-last_contributor)
        const long long thr = (long long)n_eff - 1ll - (long long)batch_base - (long long)last_contributor;
        const int thr_off = ENT * (int)max(min(thr, (long long)(2 * BWD_BATCH)), -1ll);
        // two list entries per trip and half-warp: their power / exp / alpha evaluations are independent of each other and
        // of the pixel state, so they are issued together; the hit arithmetic then runs in list order and the partial
        // sums of both entries go through ONE butterfly over the 16 lanes of the half (19 shuffles for the 2 x 2 x 9
        // values of a trip)
        for (int k0 = 0; k0 < cnt; k0 += 2) {
            const uint32_t offs = *reinterpret_cast<const uint32_t*>(&s_list[warp][half][k0]);
            const char* ents[2] = {ent_base + (offs & 0xffffu), ent_base + (offs >> 16)};
            const int offv[2] = {(int)(offs & 0xffffu), (int)(offs >> 16)};
            float dxs[2], dys[2], Gs[2], alphas[2], ids[2], invds[2];
            bool hits[2];
            unsigned any[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const float4 xy = *reinterpret_cast<const float4*>(ents[u]);
                const float4 co = *reinterpret_cast<const float4*>(ents[u] + 16);
                const float dx = F_SUB(xy.x, px.x), dy = F_SUB(xy.y, px.y);
                const float q = F_FMA(dx, F_MUL(dx, co.x), F_MUL(dy, F_MUL(dy, co.z)));
                const float power = F_FMA(q, -0.5f, -F_MUL(dy, F_MUL(dx, co.y)));  // reference SASS: FFMA(q, -0.5, -m)
                const float G = expf(power);
                const float alpha = fminf(0.99f, F_MUL(co.w, G));
                const bool hit = offv[u] > thr_off && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                dxs[u] = dx; dys[u] = dy; Gs[u] = G; alphas[u] = alpha; ids[u] = xy.z; invds[u] = xy.w;
                hits[u] = hit;
                any[u] = __ballot_sync(0xffffffffu, hit);
            }
            if ((any[0] | any[1]) == 0u) continue;  // neither half hits either entry
            // The first butterfly stage (lane ^ 8) pairs entry 0's sum k with entry 1's sum k; both are the same products
            // of four per-lane scalars, so the keep / send choice is made on the scalars (8 selects) instead of on the
            // 2 x NV products.  Three more stages (lane ^ 4, 2, 1) leave every lane of the half with at most two totals.
            const BwdHit h0 = bwd_hit_scalars<C, INVD>(px, ents[0], hits[0], dxs[0], dys[0], Gs[0], alphas[0], invds[0]);
            const BwdHit h1 = bwd_hit_scalars<C, INVD>(px, ents[1], hits[1], dxs[1], dys[1], Gs[1], alphas[1], invds[1]);
            const bool upper = (lane & 8u) != 0;
            BwdHit keep, send;
            keep.w = upper ? h1.w : h0.w;     send.w = upper ? h0.w : h1.w;
            keep.dx = upper ? h1.dx : h0.dx;  send.dx = upper ? h0.dx : h1.dx;
            keep.dy = upper ? h1.dy : h0.dy;  send.dy = upper ? h0.dy : h1.dy;
            keep.aT = upper ? h1.aT : h0.aT;  send.aT = upper ? h0.aT : h1.aT;
            float v[NV], vs[NV];
            bwd_hit_products<C, INVD, NV>(px, keep, v);
            bwd_hit_products<C, INVD, NV>(px, send, vs);
#pragma unroll
            for (int n = 0; n < NV; n++) v[n] += __shfl_xor_sync(0xffffffffu, vs[n], 8);
            int cnt_v = NV, base = upper ? NV : 0;  // what a stage over 2 * NV values leaves behind
            warp_multi_reduce_stage<NV, NV, 4>(v, lane, base, cnt_v);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (j < cnt_v) {
                    const int idx = base + j;
                    const bool second = idx >= NV;
                    int slot = idx - (second ? NV : 0);
                    if (!INVD && slot >= 6) slot += 1;  // the record keeps its inverse-depth slot
                    // totals of an entry no pixel of this half hit are exact zeros (sentinel padding included): not sent
                    if ((second ? any[1] : any[0]) & half_mask)
                        atomicAdd(grad_rec + (size_t)__float_as_uint(second ? ids[1] : ids[0]) * LG_REC + slot, v[j]);
                }
            }
        }
    }
}

int launch_blend_backward(int P, int C, int W, int H, const GeometryState& g, const BinningState& b,
                          const ImageState& img, const float* features, const float* background,
                          const float* dL_dpix, const float* dL_dinvdepth_pix, float* grad_scratch, bool debug,
                          cudaStream_t stream) {
    LG_CUDA(cudaMemsetAsync(grad_scratch, 0, sizeof(float) * LG_REC * (size_t)P, stream));
    const int gx = num_tiles_x(W), T = gx * num_tiles_y(H);
    const dim3 grid(T, 1, 1), block(LG_TILE_PIX, 1, 1);
#define LG_LAUNCH_BWD(CH, INVD)                                                                                    \
    blend_backward_kernel<CH, INVD><<<grid, block, 0, stream>>>(                                                     \
        img.ranges, b.point_list, W, H, gx, background, g.means2D, g.conic_opacity, features, g.depths,              \
        img.accum_alpha, img.n_contrib, dL_dpix, dL_dinvdepth_pix, grad_scratch, img.tile_order_bwd,                  \
        reinterpret_cast<const uint16_t*>(b.pairs))
    const bool invd = dL_dinvdepth_pix != nullptr;
    switch (C) {
        case 1: if (invd) LG_LAUNCH_BWD(1, true); else LG_LAUNCH_BWD(1, false); break;
        case 2: if (invd) LG_LAUNCH_BWD(2, true); else LG_LAUNCH_BWD(2, false); break;
        case 3: if (invd) LG_LAUNCH_BWD(3, true); else LG_LAUNCH_BWD(3, false); break;
        case 4: if (invd) LG_LAUNCH_BWD(4, true); else LG_LAUNCH_BWD(4, false); break;
        default: set_error("blend backward: unsupported channel count %d", C); return LG_ERR_UNSUPPORTED;
    }
#undef LG_LAUNCH_BWD
    LG_LAUNCH_CHECK(debug, stream);
    return LG_OK;
}

}  // namespace lg
