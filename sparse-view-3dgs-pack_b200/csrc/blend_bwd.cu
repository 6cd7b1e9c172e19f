// blend_bwd.cu — per-tile back-to-front replay of the alpha blend (backward).  Replaces renderCUDA<C> in
// DGR/cuda_rasterizer/backward.cu:452-638.
//
// The reference issues 9-10 global atomicAdds per (pixel, Gaussian) hit.  Here the 32 pixels of a warp (an 8x4
// patch) reduce their 7+C partial gradients with a halving butterfly (12 shuffles for 10 values instead of 50),
// after which 7+C *different lanes* each own one total and issue ONE coalesced red.global.add per warp into a
// packed 12-float per-Gaussian record.  Warps in which no pixel hits the Gaussian skip everything after a ballot,
// and list entries behind the last contributor of every pixel of the tile are never even staged.
#include "common.cuh"

namespace lg {

#ifndef BWD_UNROLL
#define BWD_UNROLL 2
#endif
#ifndef BWD_MIN_BLOCKS
#define BWD_MIN_BLOCKS 4
#endif

#ifndef BWD_BATCH
#define BWD_BATCH 512  // list entries staged per round (a multiple of the 256 threads)
#endif
#define LG_REC 12  // floats per packed gradient record: mean2D.xy, conic.xyw, opacity, invdepth, colour[C], pad

// Sum N per-lane values over the 32 lanes of a warp.  On return lane L holds in `out` the warp total of value
// `idx` (idx < N valid); lanes whose slot is padding get valid=false.  Halving butterfly: at every stage the two
// partner lanes split the n live values between them (the upper lane keeps the upper half), so stage sizes are
// ceil(n/2): 9 -> 5, 3, 2, 1, 1 = 12 shuffles; 18 -> 9, 5, 3, 2, 1 = 20.  Sizes are template parameters so that every
// array index is a compile-time constant (a run-time `n` put the 18-value form into local memory).
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

template <int N>
__device__ __forceinline__ void warp_multi_reduce(float (&v)[N], unsigned lane, float& out, int& idx, bool& valid) {
    int cnt = N;  // true slot count of this lane's group
    int base = 0;
    warp_multi_reduce_stage<N, N, 16>(v, lane, base, cnt);
    out = v[0];
    idx = base;
    valid = cnt >= 1;
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
    const uint32_t* __restrict__ tile_order, const uint8_t* __restrict__ entry_masks) {
    constexpr int NV = 6 + (INVD ? 1 : 0) + C;  // values reduced per (warp, Gaussian)
    constexpr int ENT = 48;                     // bytes per staged entry
    // one staged entry = three float4: (mean.x, mean.y, Gaussian id, 1/depth) (conic a, b, c, opacity) (colours, C <= 4),
    // read as warp-wide broadcasts from a single base address; entry BWD_BATCH is a sentinel that can never hit
    // (opacity 0), which pads odd-length lists
    __shared__ float4 s_ent[(BWD_BATCH + 1) * 3];
    // per 32 staged entries and patch: which of them can touch the patch at all (ballots of the staging warps over the
    // mask bytes the forward pass left per list entry)
    __shared__ uint32_t s_reach[BWD_BATCH / 32][LG_TILE_PIX / 32];
    __shared__ __align__(4) lg_slot_t s_list[LG_TILE_PIX / 32][BWD_BATCH + 2];  // per warp: byte offsets into s_ent
    __shared__ uint32_t s_max;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = tile_order[blockIdx.x];  // deepest tiles first
    const uint32_t tile_x = tile % (uint32_t)grid_x, tile_y = tile / (uint32_t)grid_x;
    const uint32_t pix_x = tile_x * LG_TILE_X + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t pix_y = tile_y * LG_TILE_Y + (warp >> 1) * 4u + (lane >> 3);
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
            for (int b = 0; b < LG_TILE_PIX / 32; b++) {
                const uint32_t bal = __ballot_sync(0xffffffffu, (mask >> b) & 1u);
                word = lane == (unsigned)b ? bal : word;
            }
            if (lane < LG_TILE_PIX / 32) s_reach[slot >> 5][lane] = word;
        }
        __syncthreads();
        const int batch = (int)min((uint32_t)BWD_BATCH, n_eff - batch_base);
        // slots whose list position is behind this warp's last contributor are dropped with the unreachable ones:
        // position rel = n_eff-1-(batch_base+j) < warp_max  <=>  j >= n_eff - warp_max - batch_base
        const int first = (int)max((long long)n_eff - (long long)warp_max - (long long)batch_base, 0ll);
        if (first >= batch) continue;
        // compacted list of this warp's reachable entries as byte offsets into s_ent, padded to an even length
        // (words of chunks past the end of the list are zero: the staging loop covers all BWD_BATCH slots)
        int cnt = 0;
        {
            const uint32_t list_addr = (uint32_t)__cvta_generic_to_shared(&s_list[warp][0]);
            const unsigned lt = (1u << lane) - 1u;
            const uint32_t my_off = lane * ENT;
#pragma unroll
            for (int c = 0; c < BWD_BATCH / 32; c++) {
                const int drop = min(max(first - c * 32, 0), 32);  // leading slots of this chunk behind `first`
                const uint32_t bal = s_reach[c][warp] & (uint32_t)(0xffffffffull << drop);
                if ((bal >> lane) & 1u)
                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr + 2u * (uint32_t)(cnt + __popc(bal & lt))),
                                 "h"((unsigned short)(my_off + c * 32 * ENT)) : "memory");
                cnt += __popc(bal);
            }
            if (lane == 0)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(list_addr + 2u * (uint32_t)cnt),
                             "h"((unsigned short)(BWD_BATCH * ENT)) : "memory");
            __syncwarp();
            cnt = (int)__reduce_max_sync(0xffffffffu, (unsigned)cnt);  // same in every lane; tells the compiler so
        }
        // this pixel is behind entry j of the batch  <=>  rel = n_eff-1-(batch_base+j) < last_contributor
        //                                           <=>  j*ENT > ENT * (n_eff-1-batch_base-last_contributor)
        const long long thr = (long long)n_eff - 1ll - (long long)batch_base - (long long)last_contributor;
        const int thr_off = ENT * (int)max(min(thr, (long long)(2 * BWD_BATCH)), -1ll);
        // two list entries per trip: their power / exp / alpha evaluations are independent of each other and of the
        // pixel state, so they are issued together; the hit arithmetic then runs in list order and the partial sums of
        // both entries go through ONE butterfly (20 shuffles for 2 x 9 values instead of 2 x 12)
        for (int k0 = 0; k0 < cnt; k0 += 2) {
            const uint32_t offs = *reinterpret_cast<const uint32_t*>(&s_list[warp][k0]);
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
            if ((any[0] | any[1]) == 0u) continue;
            float total;
            int slot;
            bool ok;
            float idf;
            if (any[0] != 0u && any[1] != 0u) {
                // both entries hit somewhere in the patch: the first butterfly stage pairs entry 0's sum k with entry
                // 1's sum k (upper half-warp keeps entry 1's).  Both are the same products of four per-lane scalars,
                // so the keep / send choice is made on the scalars (8 selects) instead of on the 2 x NV products
                const BwdHit h0 = bwd_hit_scalars<C, INVD>(px, ents[0], hits[0], dxs[0], dys[0], Gs[0], alphas[0], invds[0]);
                const BwdHit h1 = bwd_hit_scalars<C, INVD>(px, ents[1], hits[1], dxs[1], dys[1], Gs[1], alphas[1], invds[1]);
                const bool upper = (lane & 16u) != 0;
                BwdHit keep, send;
                keep.w = upper ? h1.w : h0.w;     send.w = upper ? h0.w : h1.w;
                keep.dx = upper ? h1.dx : h0.dx;  send.dx = upper ? h0.dx : h1.dx;
                keep.dy = upper ? h1.dy : h0.dy;  send.dy = upper ? h0.dy : h1.dy;
                keep.aT = upper ? h1.aT : h0.aT;  send.aT = upper ? h0.aT : h1.aT;
                float v[NV], vs[NV];
                bwd_hit_products<C, INVD, NV>(px, keep, v);
                bwd_hit_products<C, INVD, NV>(px, send, vs);
#pragma unroll
                for (int n = 0; n < NV; n++) v[n] += __shfl_xor_sync(0xffffffffu, vs[n], 16);
                int cnt = NV, base = upper ? NV : 0;  // what the stage of 2 * NV values leaves behind
                warp_multi_reduce_stage<NV, NV, 8>(v, lane, base, cnt);
                total = v[0];
                slot = base;
                ok = cnt >= 1;
                const bool second = slot >= NV;
                idf = second ? ids[1] : ids[0];
                slot -= second ? NV : 0;
            } else if (any[0] != 0u) {
                float v[NV];
                bwd_hit_values<C, INVD, NV>(px, ents[0], hits[0], dxs[0], dys[0], Gs[0], alphas[0], invds[0], v);
                warp_multi_reduce<NV>(v, lane, total, slot, ok);
                idf = ids[0];
            } else {
                float v[NV];
                bwd_hit_values<C, INVD, NV>(px, ents[1], hits[1], dxs[1], dys[1], Gs[1], alphas[1], invds[1], v);
                warp_multi_reduce<NV>(v, lane, total, slot, ok);
                idf = ids[1];
            }
            if (!INVD && slot >= 6) slot += 1;  // the record keeps its inverse-depth slot
            if (ok) atomicAdd(grad_rec + (size_t)__float_as_uint(idf) * LG_REC + slot, total);
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
        reinterpret_cast<const uint8_t*>(b.pairs))
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
