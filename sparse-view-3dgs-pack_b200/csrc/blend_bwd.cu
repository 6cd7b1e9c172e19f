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
// `idx` (idx < N valid); lanes whose slot is padding get valid=false.
template <int N>
__device__ __forceinline__ void warp_multi_reduce(float (&v)[N], unsigned lane, float& out, int& idx, bool& valid) {
    int n = N;        // padded slot count at this stage (compile-time after unrolling)
    int cnt = N;      // true slot count of this lane's group
    int base = 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
        if (n > 1) {
            const int h = (n + 1) / 2;
#pragma unroll
            for (int k = 0; k < h; k++) {
                const float lo = v[k];
                const float hi = (k + h < n) ? v[k + h] : 0.0f;
                const float send = upper ? lo : hi;
                const float keep = upper ? hi : lo;
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
            if (upper) { base += h; cnt -= h; } else { cnt = min(cnt, h); }
            n = h;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
            if (upper) cnt = 0;  // both partners now hold the same total: the lower lane owns it
        }
    }
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
template <int C, bool INVD>
__global__ void __launch_bounds__(LG_TILE_PIX, BWD_MIN_BLOCKS) blend_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float* __restrict__ bg_color, const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
    const float* __restrict__ colors, const float* __restrict__ depths, const float* __restrict__ final_Ts,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dinvdepth_pix, float* __restrict__ grad_rec,
    const uint32_t* __restrict__ tile_order) {
    constexpr int NV = 6 + (INVD ? 1 : 0) + C;  // values reduced per (warp, Gaussian)
    constexpr int VC = 6 + (INVD ? 1 : 0);      // first colour slot among the reduced values
    // one staged entry = three float4: (mean.x, mean.y, Gaussian id, 1/depth) (conic a, b, c, opacity) (colours, C <= 4),
    // read as warp-wide broadcasts from a single base address
    __shared__ float4 s_ent[BWD_BATCH * 3];
    __shared__ uint8_t s_mask[BWD_BATCH];                    // per staged entry: which of the 8 patches it can touch
    __shared__ lg_slot_t s_list[LG_TILE_PIX / 32][BWD_BATCH];  // per warp: compacted slots it must evaluate
    __shared__ uint32_t s_max;

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile = tile_order[blockIdx.x];  // deepest tiles first
    const uint32_t tile_x = tile % (uint32_t)grid_x, tile_y = tile / (uint32_t)grid_x;
    const uint32_t pix_x = tile_x * LG_TILE_X + (warp & 1u) * 8u + (lane & 7u);
    const uint32_t pix_y = tile_y * LG_TILE_Y + (warp >> 1) * 4u + (lane >> 3);
    const bool inside = pix_x < (uint32_t)W && pix_y < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * pix_y + pix_x;
    const float pixf_x = (float)pix_x, pixf_y = (float)pix_y;
    const float tile_x0 = (float)(tile_x * LG_TILE_X), tile_y0 = (float)(tile_y * LG_TILE_Y);
    const uint2 range = ranges[tile];

    const float T_final = inside ? final_Ts[pix_id] : 0.0f;
    float T = T_final;
    const uint32_t last_contributor = inside ? n_contrib[pix_id] : 0u;

    if (tid == 0) s_max = 0;
    __syncthreads();
    const uint32_t warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0 && warp_max) atomicMax(&s_max, warp_max);
    __syncthreads();
    const uint32_t n_eff = s_max;  // entries [0, n_eff) of this tile's list can contribute to some pixel
    if (n_eff == 0) return;

    float dL_dpixel[C];
    float accum_rec[C], last_color[C];
    float bg_dot_dpixel = 0.0f;
#pragma unroll
    for (int c = 0; c < C; c++) {
        dL_dpixel[c] = inside ? dL_dpixels[(size_t)c * H * W + pix_id] : 0.0f;
        accum_rec[c] = 0.0f;
        last_color[c] = 0.0f;
        bg_dot_dpixel += bg_color[c] * dL_dpixel[c];
    }
    const float bg_term = -T_final * bg_dot_dpixel;
    float dL_invd = 0.0f, accum_invd_rec = 0.0f, last_invd = 0.0f;
    if (INVD) dL_invd = inside ? dL_dinvdepth_pix[pix_id] : 0.0f;
    float last_alpha = 0.0f;

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
                const uint32_t id = point_list[range.x + (n_eff - 1u - progress)];
                const float2 m = means2D[id];
                const float4 cq = conic_opacity[id];
                mask = lg_patch_mask(m.x, m.y, cq, tile_x0, tile_y0);
                float fv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int c = 0; c < C; c++) fv[c] = colors[(size_t)id * C + c];
                s_ent[slot * 3 + 0] = make_float4(m.x, m.y, __uint_as_float(id), INVD ? 1.0f / depths[id] : 0.0f);
                s_ent[slot * 3 + 1] = cq;
                s_ent[slot * 3 + 2] = make_float4(fv[0], fv[1], fv[2], fv[3]);
            }
            s_mask[slot] = (uint8_t)mask;
        }
        __syncthreads();
        const int batch = (int)min((uint32_t)BWD_BATCH, n_eff - batch_base);
        // slots whose list position is behind this warp's last contributor are dropped with the unreachable ones:
        // position rel = n_eff-1-(batch_base+j) < warp_max  <=>  j >= n_eff - warp_max - batch_base
        const int first = (int)max((long long)n_eff - (long long)warp_max - (long long)batch_base, 0ll);
        if (first >= batch) continue;
        const int cnt = lg_compact_patch_list(s_mask, s_list[warp], warp, lane, first, batch);
        // BWD_UNROLL list entries per trip: their power / exp / alpha evaluations are independent of each other and of
        // the pixel state, so they are issued together; the hit paths then run in list order.
        for (int k0 = 0; k0 < cnt; k0 += BWD_UNROLL) {
          int js[BWD_UNROLL];
          float dxs[BWD_UNROLL], dys[BWD_UNROLL], Gs[BWD_UNROLL], alphas[BWD_UNROLL], ids[BWD_UNROLL], invds[BWD_UNROLL];
          bool hits[BWD_UNROLL];
          unsigned any[BWD_UNROLL];
#pragma unroll
          for (int u = 0; u < BWD_UNROLL; u++) {
            const bool has = k0 + u < cnt;
            const int j = s_list[warp][has ? k0 + u : k0];
            const uint32_t rel = n_eff - 1u - (batch_base + (uint32_t)j);  // 0-based position in the tile's list
            const float4 xy = s_ent[j * 3 + 0];
            const float4 co = s_ent[j * 3 + 1];
            const float dx = F_SUB(xy.x, pixf_x), dy = F_SUB(xy.y, pixf_y);
            const float q = F_FMA(dx, F_MUL(dx, co.x), F_MUL(dy, F_MUL(dy, co.z)));
            const float power = F_FMA(q, -0.5f, -F_MUL(dy, F_MUL(dx, co.y)));  // reference SASS: FFMA(q, -0.5, -m)
            const float G = expf(power);
            const float alpha = fminf(0.99f, F_MUL(co.w, G));
            const bool hit = has && rel < last_contributor && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            js[u] = j; dxs[u] = dx; dys[u] = dy; Gs[u] = G; alphas[u] = alpha; ids[u] = xy.z; invds[u] = xy.w;
            hits[u] = hit;
            any[u] = __ballot_sync(0xffffffffu, hit);
          }
#pragma unroll
          for (int u = 0; u < BWD_UNROLL; u++) {
            if (any[u] == 0u) continue;
            const int j = js[u];
            const float dx = dxs[u], dy = dys[u], G = Gs[u], alpha = alphas[u];
            const bool hit = hits[u];
            float4 xy;
            xy.z = ids[u];
            xy.w = invds[u];

            float v[NV];
#pragma unroll
            for (int n = 0; n < NV; n++) v[n] = 0.0f;
            if (hit) {
                float rinv;  // 1 - alpha lies in [0.01, 1]: the bare MUFU.RCP needs no range fix-up
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(1.0f - alpha));
                T = T * rinv;
                const float aT = alpha * T;
                float dL_dalpha = 0.0f;
                const float4 f4 = s_ent[j * 3 + 2];
                const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const float col = fv[c];
                    accum_rec[c] = fmaf(last_alpha, last_color[c] - accum_rec[c], accum_rec[c]);
                    last_color[c] = col;
                    dL_dalpha = fmaf(col - accum_rec[c], dL_dpixel[c], dL_dalpha);
                    v[VC + c] = aT * dL_dpixel[c];
                }
                if (INVD) {
                    const float invd = xy.w;
                    accum_invd_rec = fmaf(last_alpha, last_invd - accum_invd_rec, accum_invd_rec);
                    last_invd = invd;
                    dL_dalpha = fmaf(invd - accum_invd_rec, dL_invd, dL_dalpha);
                    v[6] = aT * dL_invd;
                }
                last_alpha = alpha;
                dL_dalpha = fmaf(dL_dalpha, T, bg_term * rinv);
                const float w = G * dL_dalpha;
                const float wx = w * dx, wy = w * dy;
                v[0] = wx;
                v[1] = wy;
                v[2] = wx * dx;
                v[3] = wx * dy;
                v[4] = wy * dy;
                v[5] = w;
            }
            float total;
            int slot;
            bool ok;
            warp_multi_reduce<NV>(v, lane, total, slot, ok);
            if (!INVD && slot >= 6) slot += 1;  // the record keeps its inverse-depth slot
            if (ok) atomicAdd(grad_rec + (size_t)__float_as_uint(xy.z) * LG_REC + slot, total);
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
        img.accum_alpha, img.n_contrib, dL_dpix, dL_dinvdepth_pix, grad_scratch, img.tile_order_bwd)
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
