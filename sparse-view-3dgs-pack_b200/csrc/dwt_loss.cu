// dwt_loss.cu — fused Haar-DWT loss of LGDWT-GS.  One pass over the render/GT pair produces the eight global
// sub-band L1 sums (2 levels), the GT ELF map at half resolution and the per-128px-patch high-band L1 sums; a small
// second pass turns the ELF map into per-patch means and a one-block epilogue does the k-th-value patch selection
// and assembles the scalar losses.  The backward regenerates the signs from pred/gt in one pass.
// Replaces the PyTorch op chain LG/train.py:131-180 = get_dwt_subbands (LG/utils/loss_utils.py:106-153, four
// pytorch_wavelets.DWTForward(J=1,'symmetric','db1') transforms per call), l1_loss (:40-41), compute_elf_map
// (:336-366) and compute_patch_dwt_loss (:368-442).  Transform definition: SURVEY.md App. B.
#include "common.cuh"

namespace lg {

#define DWT_S 0.70710678118654752440f  // pywt db1 filter tap, stored as fp32 by pytorch_wavelets
#define DWT_TB 16                      // 16x16 threads, one level-2 coefficient (4x4 pixels) each

struct DwtWorkspace {
    double* band_sums;    // 8
    double* patch_sums;   // 3 per patch: LH1, HL1, HH1 |pred-gt| sums
    double* patch_elf;    // 1 per patch: sum of the upsampled ELF map
    float* elf_low;       // H2 * W2
    static DwtWorkspace from_chunk(char*& chunk, int C, int H, int W, int ps) {
        (void)C;
        DwtWorkspace w;
        const size_t H2 = (H + 1) / 2, W2 = (W + 1) / 2;
        const size_t L = ps > 0 ? (size_t)(H / ps) * (W / ps) : 0;
        carve(chunk, w.band_sums, 8);
        carve(chunk, w.patch_sums, 3 * L + 1);
        carve(chunk, w.patch_elf, L + 1);
        carve(chunk, w.elf_low, H2 * W2);
        return w;
    }
};

struct HaarBands { float ll, lh, hl, hh; };

// one 2x2 analysis step in the operation order of pytorch_wavelets' AFB2D: filter along W, then along H
// Explicitly rounded (never FMA-contracted) so that s*a - s*b is exactly 0 when a == b — at the symmetric-extension
// boundary and on flat regions the high bands must vanish exactly, otherwise sign() of a rounding residual leaks a
// spurious +-1 into the gradient — and so that every band equals the fp32 strided-convolution result bit for bit.
__device__ __forceinline__ HaarBands haar2x2(float x00, float x01, float x10, float x11) {
    const float a0 = F_MUL(DWT_S, x00), a1 = F_MUL(DWT_S, x01), b0 = F_MUL(DWT_S, x10), b1 = F_MUL(DWT_S, x11);
    const float lo_t = F_MUL(DWT_S, F_ADD(a0, a1)), hi_t = F_MUL(DWT_S, F_SUB(a0, a1));
    const float lo_b = F_MUL(DWT_S, F_ADD(b0, b1)), hi_b = F_MUL(DWT_S, F_SUB(b0, b1));
    HaarBands b;
    b.ll = F_ADD(lo_t, lo_b);
    b.lh = F_SUB(lo_t, lo_b);
    b.hl = F_ADD(hi_t, hi_b);
    b.hh = F_SUB(hi_t, hi_b);
    return b;
}

// adjoint of haar2x2
__device__ __forceinline__ void haar2x2_adjoint(float g_ll, float g_lh, float g_hl, float g_hh, float& d00, float& d01,
                                                float& d10, float& d11) {
    const float d_lo_t = DWT_S * (g_ll + g_lh), d_lo_b = DWT_S * (g_ll - g_lh);
    const float d_hi_t = DWT_S * (g_hl + g_hh), d_hi_b = DWT_S * (g_hl - g_hh);
    d00 = DWT_S * (d_lo_t + d_hi_t);
    d01 = DWT_S * (d_lo_t - d_hi_t);
    d10 = DWT_S * (d_lo_b + d_hi_b);
    d11 = DWT_S * (d_lo_b - d_hi_b);
}

// 4x4 pixel block behind level-2 coefficient (i2, j2), with symmetric (duplicate-last) extension at both levels.
__device__ __forceinline__ void load_block4(const float* __restrict__ img, int H, int W, int H2, int W2, int i2, int j2,
                                            float (&x)[4][4]) {
    int rows[4], cols[4];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const int i1 = min(2 * i2 + u, H2 - 1), j1 = min(2 * j2 + u, W2 - 1);
        rows[2 * u] = 2 * i1; rows[2 * u + 1] = min(2 * i1 + 1, H - 1);
        cols[2 * u] = 2 * j1; cols[2 * u + 1] = min(2 * j1 + 1, W - 1);
    }
    const bool fast = ((W & 3) == 0) && (4 * j2 + 3 < W);
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const float* row = img + (size_t)rows[a] * W;
        if (fast) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + 4 * j2));
            x[a][0] = v.x; x[a][1] = v.y; x[a][2] = v.z; x[a][3] = v.w;
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++) x[a][b] = __ldg(row + cols[b]);
        }
    }
}

__device__ __forceinline__ float block_sum(float v, float* s_tmp) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_tmp[warp] = v;
    __syncthreads();
    v = lane < (blockDim.x >> 5) ? s_tmp[lane] : 0.f;
    return warp_sum(v);
}

// ---------------------------------------------------------------- forward main pass
__global__ void __launch_bounds__(DWT_TB * DWT_TB) dwt_forward_kernel(const float* __restrict__ pred,
                                                                      const float* __restrict__ gt, int C, int H, int W,
                                                                      int ps, int nPW, int nPH, double* band_sums,
                                                                      double* patch_sums, float* __restrict__ elf_low) {
    __shared__ float s_tmp[32];
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2, H4 = (H2 + 1) / 2, W4 = (W2 + 1) / 2;
    const int tx = threadIdx.x % DWT_TB, ty = threadIdx.x / DWT_TB;
    const int j2 = blockIdx.x * DWT_TB + tx, i2 = blockIdx.y * DWT_TB + ty;
    const bool in = i2 < H4 && j2 < W4;

    float sums[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float elf_ll[2][2] = {{0, 0}, {0, 0}}, elf_hf[2][2] = {{0, 0}, {0, 0}};
    float hb[2][2][3] = {};  // per level-1 position: |pred-gt| of LH, HL, HH summed over channels (patch term)
    if (in) {
        for (int c = 0; c < C; c++) {
            float p[4][4], g[4][4];
            load_block4(pred + (size_t)c * H * W, H, W, H2, W2, i2, j2, p);
            load_block4(gt + (size_t)c * H * W, H, W, H2, W2, i2, j2, g);
            float pll[2][2], gll[2][2];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int v = 0; v < 2; v++) {
                    const HaarBands bp = haar2x2(p[2 * u][2 * v], p[2 * u][2 * v + 1], p[2 * u + 1][2 * v], p[2 * u + 1][2 * v + 1]);
                    const HaarBands bg = haar2x2(g[2 * u][2 * v], g[2 * u][2 * v + 1], g[2 * u + 1][2 * v], g[2 * u + 1][2 * v + 1]);
                    pll[u][v] = bp.ll; gll[u][v] = bg.ll;
                    if (2 * i2 + u < H2 && 2 * j2 + v < W2) {
                        const float dlh = fabsf(bp.lh - bg.lh), dhl = fabsf(bp.hl - bg.hl), dhh = fabsf(bp.hh - bg.hh);
                        sums[0] += fabsf(bp.ll - bg.ll);
                        sums[1] += dlh; sums[2] += dhl; sums[3] += dhh;
                        hb[u][v][0] += dlh; hb[u][v][1] += dhl; hb[u][v][2] += dhh;
                        elf_ll[u][v] += fabsf(bg.ll);
                        elf_hf[u][v] += fabsf(bg.lh) + fabsf(bg.hl) + fabsf(bg.hh);
                    }
                }
            const HaarBands p2 = haar2x2(pll[0][0], pll[0][1], pll[1][0], pll[1][1]);
            const HaarBands g2 = haar2x2(gll[0][0], gll[0][1], gll[1][0], gll[1][1]);
            sums[4] += fabsf(p2.ll - g2.ll); sums[5] += fabsf(p2.lh - g2.lh);
            sums[6] += fabsf(p2.hl - g2.hl); sums[7] += fabsf(p2.hh - g2.hh);
        }
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const int i1 = 2 * i2 + u, j1 = 2 * j2 + v;
                if (i1 < H2 && j1 < W2)
                    elf_low[(size_t)i1 * W2 + j1] = elf_ll[u][v] / (elf_ll[u][v] + elf_hf[u][v] + 1e-8f);
            }
    }
    // patch high-band sums: level-1 coefficient (i1, j1) lies in patch (i1 / (ps/2), j1 / (ps/2))
    if (ps > 0 && nPW > 0 && nPH > 0) {
        const int half = ps / 2;
        const bool uniform = (half % (2 * DWT_TB)) == 0;  // the block's 32x32 level-1 coefficients share one patch
        if (uniform) {
            const int py = (blockIdx.y * 2 * DWT_TB) / half, px = (blockIdx.x * 2 * DWT_TB) / half;
            float t[3] = {0, 0, 0};
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int v = 0; v < 2; v++)
#pragma unroll
                    for (int k = 0; k < 3; k++) t[k] += hb[u][v][k];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float s = block_sum(t[k], s_tmp);
                if (threadIdx.x == 0 && py < nPH && px < nPW) atomicAdd(&patch_sums[3 * (py * nPW + px) + k], (double)s);
            }
        } else if (in) {
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int v = 0; v < 2; v++) {
                    const int i1 = 2 * i2 + u, j1 = 2 * j2 + v;
                    const int py = i1 / half, px = j1 / half;
                    if (i1 < H2 && j1 < W2 && py < nPH && px < nPW)
#pragma unroll
                        for (int k = 0; k < 3; k++) atomicAdd(&patch_sums[3 * (py * nPW + px) + k], (double)hb[u][v][k]);
                }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float s = block_sum(sums[k], s_tmp);
        if (threadIdx.x == 0) atomicAdd(&band_sums[k], (double)s);
    }
}

// ---------------------------------------------------------------- ELF bilinear upsample + per-patch sum
// F.interpolate(mode='bilinear', align_corners=False) to (H, W) (LG/utils/loss_utils.py:364), then the per-patch
// mean of LG/utils/loss_utils.py:396-402.  grid = (patch, slice); each block sums ps*ps/gridDim.y pixels.
__global__ void __launch_bounds__(256) dwt_elf_patch_kernel(const float* __restrict__ elf_low, int H, int W, int ps,
                                                            int nPW, double* patch_elf) {
    __shared__ float s_tmp[32];
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const float scale_y = (float)H2 / (float)H, scale_x = (float)W2 / (float)W;
    const int patch = blockIdx.x, py = patch / nPW, px = patch % nPW;
    const int total = ps * ps;
    float acc = 0.f;
    for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < total; t += gridDim.y * blockDim.x) {
        const int y = py * ps + t / ps, x = px * ps + t % ps;
        float sy = scale_y * ((float)y + 0.5f) - 0.5f, sx = scale_x * ((float)x + 0.5f) - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int y0 = min((int)sy, H2 - 1), x0 = min((int)sx, W2 - 1);
        const int y1 = y0 + (y0 < H2 - 1 ? 1 : 0), x1 = x0 + (x0 < W2 - 1 ? 1 : 0);
        const float ly1 = sy - (float)y0, lx1 = sx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
        const float v00 = elf_low[(size_t)y0 * W2 + x0], v01 = elf_low[(size_t)y0 * W2 + x1];
        const float v10 = elf_low[(size_t)y1 * W2 + x0], v11 = elf_low[(size_t)y1 * W2 + x1];
        acc += ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
    }
    const float s = block_sum(acc, s_tmp);
    if (threadIdx.x == 0) atomicAdd(&patch_elf[patch], (double)s);
}

// ---------------------------------------------------------------- epilogue: selection + scalar losses
struct DwtFinalizeArgs {
    float w[8];
    float w_lh, w_hl;
    int C, H, W, ps, L, k;
};

__global__ void __launch_bounds__(1024) dwt_finalize_kernel(DwtFinalizeArgs a, const double* __restrict__ band_sums,
                                                            const double* __restrict__ patch_sums,
                                                            const double* __restrict__ patch_elf,
                                                            float* __restrict__ out, uint8_t* __restrict__ mask) {
    extern __shared__ float s_mean[];  // L patch means
    __shared__ float s_thr;
    __shared__ double s_sel[4];
    const int H2 = (a.H + 1) / 2, W2 = (a.W + 1) / 2, H4 = (H2 + 1) / 2, W4 = (W2 + 1) / 2;
    const int tid = threadIdx.x;
    if (tid == 0) {
        const double n1 = (double)a.C * H2 * W2, n2 = (double)a.C * H4 * W4;
        float total = 0.f;
        for (int b = 0; b < 8; b++) {
            const float m = (float)(band_sums[b] / (b < 4 ? n1 : n2));
            out[2 + b] = m;
            if (a.w[b] != 0.0f) total += a.w[b] * m;
        }
        out[0] = total;
        s_sel[0] = s_sel[1] = s_sel[2] = s_sel[3] = 0.0;
        s_thr = 0.f;
    }
    for (int i = tid; i < a.L; i += blockDim.x) s_mean[i] = (float)(patch_elf[i] / ((double)a.ps * a.ps));
    __syncthreads();
    if (a.L > 0) {
        // k-th smallest (1-indexed, torch.kthvalue): the element whose stable rank is k-1
        for (int i = tid; i < a.L; i += blockDim.x) {
            const float v = s_mean[i];
            int rank = 0;
            for (int j = 0; j < a.L; j++) {
                const float o = s_mean[j];
                rank += (o < v || (o == v && j < i)) ? 1 : 0;
            }
            if (rank == a.k - 1) s_thr = v;
        }
        __syncthreads();
        const float thr = s_thr;
        for (int i = tid; i < a.L; i += blockDim.x) {
            const bool sel = s_mean[i] >= thr;
            mask[i] = sel ? 1 : 0;
            if (sel) {
                atomicAdd(&s_sel[0], patch_sums[3 * i + 0]);
                atomicAdd(&s_sel[1], patch_sums[3 * i + 1]);
                atomicAdd(&s_sel[2], patch_sums[3 * i + 2]);
                atomicAdd(&s_sel[3], 1.0);
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        float patch_loss = 0.f;
        const double nsel = a.L > 0 ? s_sel[3] : 0.0;
        if (nsel > 0.0) {
            const double half = a.ps / 2;
            const double n = nsel * a.C * half * half;
            const float l_lh = (float)(s_sel[0] / n), l_hl = (float)(s_sel[1] / n), l_hh = (float)(s_sel[2] / n);
            patch_loss = (a.w_lh * l_lh) + (a.w_hl * l_hl) + (0.5f * (a.w_lh + a.w_hl) * l_hh);
        }
        out[1] = patch_loss;
        out[10] = (float)nsel;
        out[11] = s_thr;
    }
}

// ---------------------------------------------------------------- backward
struct DwtBackwardArgs {
    float w[8];
    float w_lh, w_hl;
    int C, H, W, ps, nPW, nPH;
};

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(DWT_TB * DWT_TB) dwt_backward_kernel(const float* __restrict__ pred,
                                                                       const float* __restrict__ gt, DwtBackwardArgs a,
                                                                       const float* __restrict__ g_dwt_p,
                                                                       const float* __restrict__ g_patch_p,
                                                                       const uint8_t* __restrict__ mask,
                                                                       const float* __restrict__ out_losses,
                                                                       const float* __restrict__ g_up, int accumulate,
                                                                       float* __restrict__ dpred) {
    const int H = a.H, W = a.W, C = a.C;
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2, H4 = (H2 + 1) / 2, W4 = (W2 + 1) / 2;
    const int tx = threadIdx.x % DWT_TB, ty = threadIdx.x / DWT_TB;
    const int j2 = blockIdx.x * DWT_TB + tx, i2 = blockIdx.y * DWT_TB + ty;
    if (i2 >= H4 || j2 >= W4) return;
    const float up = g_up ? *g_up : 1.f;  // upstream gradient of the combined loss (device scalar)
    const float g_dwt = (g_dwt_p ? *g_dwt_p : 1.f) * up, g_patch = (g_patch_p ? *g_patch_p : 0.f) * up;
    const float n1 = (float)C * H2 * W2, n2 = (float)C * H4 * W4;
    float k1[4], k2[4];
#pragma unroll
    for (int b = 0; b < 4; b++) { k1[b] = g_dwt * a.w[b] / n1; k2[b] = g_dwt * a.w[4 + b] / n2; }
    const float nsel = a.ps > 0 ? out_losses[10] : 0.f;
    const int half = a.ps > 0 ? a.ps / 2 : 1;
    float kp[3] = {0, 0, 0};
    if (nsel > 0.f) {
        const float n = nsel * C * (float)half * (float)half;
        kp[0] = g_patch * a.w_lh / n; kp[1] = g_patch * a.w_hl / n; kp[2] = g_patch * 0.5f * (a.w_lh + a.w_hl) / n;
    }
    bool selected[2][2];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const int py = (2 * i2 + u) / half, px = (2 * j2 + v) / half;
            selected[u][v] = nsel > 0.f && py < a.nPH && px < a.nPW && mask[py * a.nPW + px] != 0;
        }
    const bool fast = ((W & 3) == 0) && (4 * j2 + 3 < W);
    for (int c = 0; c < C; c++) {
        float p[4][4], g[4][4];
        load_block4(pred + (size_t)c * H * W, H, W, H2, W2, i2, j2, p);
        load_block4(gt + (size_t)c * H * W, H, W, H2, W2, i2, j2, g);
        HaarBands d1[2][2];
        float pll[2][2], gll[2][2];
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const HaarBands bp = haar2x2(p[2 * u][2 * v], p[2 * u][2 * v + 1], p[2 * u + 1][2 * v], p[2 * u + 1][2 * v + 1]);
                const HaarBands bg = haar2x2(g[2 * u][2 * v], g[2 * u][2 * v + 1], g[2 * u + 1][2 * v], g[2 * u + 1][2 * v + 1]);
                pll[u][v] = bp.ll; gll[u][v] = bg.ll;
                d1[u][v].ll = bp.ll - bg.ll; d1[u][v].lh = bp.lh - bg.lh; d1[u][v].hl = bp.hl - bg.hl; d1[u][v].hh = bp.hh - bg.hh;
            }
        const HaarBands p2 = haar2x2(pll[0][0], pll[0][1], pll[1][0], pll[1][1]);
        const HaarBands g2 = haar2x2(gll[0][0], gll[0][1], gll[1][0], gll[1][1]);
        float gl[2][2];  // level-2 adjoint onto the 2x2 LL1 block
        haar2x2_adjoint(k2[0] * sgn(p2.ll - g2.ll), k2[1] * sgn(p2.lh - g2.lh), k2[2] * sgn(p2.hl - g2.hl),
                        k2[3] * sgn(p2.hh - g2.hh), gl[0][0], gl[0][1], gl[1][0], gl[1][1]);
        float dx[4][4];
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const bool valid = 2 * i2 + u < H2 && 2 * j2 + v < W2;
                float G_ll = 0.f, G_lh = 0.f, G_hl = 0.f, G_hh = 0.f;
                if (valid) {
                    G_ll = k1[0] * sgn(d1[u][v].ll) + gl[u][v];
                    G_lh = k1[1] * sgn(d1[u][v].lh);
                    G_hl = k1[2] * sgn(d1[u][v].hl);
                    G_hh = k1[3] * sgn(d1[u][v].hh);
                    if (selected[u][v]) {
                        G_lh += kp[0] * sgn(d1[u][v].lh);
                        G_hl += kp[1] * sgn(d1[u][v].hl);
                        G_hh += kp[2] * sgn(d1[u][v].hh);
                    }
                }
                haar2x2_adjoint(G_ll, G_lh, G_hl, G_hh, dx[2 * u][2 * v], dx[2 * u][2 * v + 1], dx[2 * u + 1][2 * v],
                                dx[2 * u + 1][2 * v + 1]);
            }
        // write the un-padded part of the 4x4 block (gradients of padding samples are dropped = the package's crop)
        float* out = dpred + (size_t)c * H * W;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int y = 4 * i2 + r;
            if (y >= H) continue;
            // accumulate: the image already holds the photometric gradient (lg_image_loss_backward); sum = held + ours
            if (fast) {
                float4* o4 = reinterpret_cast<float4*>(out + (size_t)y * W + 4 * j2);
                float4 v = make_float4(dx[r][0], dx[r][1], dx[r][2], dx[r][3]);
                if (accumulate) {
                    const float4 h = *o4;
                    v = make_float4(h.x + v.x, h.y + v.y, h.z + v.z, h.w + v.w);
                }
                *o4 = v;
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int x = 4 * j2 + q;
                    if (x < W) out[(size_t)y * W + x] = accumulate ? out[(size_t)y * W + x] + dx[r][q] : dx[r][q];
                }
            }
        }
    }
}

// ---------------------------------------------------------------- plain single-level transform (compat module)
__global__ void __launch_bounds__(256) haar_dwt2_forward_kernel(const float* __restrict__ x, int planes, int H, int W,
                                                                float* __restrict__ ll, float* __restrict__ yh) {
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const size_t n = (size_t)planes * H2 * W2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int j = (int)(t % W2), i = (int)((t / W2) % H2);
    const size_t pl = t / ((size_t)W2 * H2);
    const float* img = x + pl * H * W;
    const int r0 = 2 * i, r1 = min(2 * i + 1, H - 1), c0 = 2 * j, c1 = min(2 * j + 1, W - 1);
    const HaarBands b = haar2x2(img[(size_t)r0 * W + c0], img[(size_t)r0 * W + c1], img[(size_t)r1 * W + c0],
                                img[(size_t)r1 * W + c1]);
    const size_t o = (size_t)i * W2 + j, band = (size_t)H2 * W2;
    ll[pl * band + o] = b.ll;
    yh[(pl * 3 + 0) * band + o] = b.lh;
    yh[(pl * 3 + 1) * band + o] = b.hl;
    yh[(pl * 3 + 2) * band + o] = b.hh;
}

__global__ void __launch_bounds__(256) haar_dwt2_backward_kernel(const float* __restrict__ g_ll,
                                                                 const float* __restrict__ g_yh, int planes, int H,
                                                                 int W, float* __restrict__ g_x) {
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const size_t n = (size_t)planes * H2 * W2;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int j = (int)(t % W2), i = (int)((t / W2) % H2);
    const size_t pl = t / ((size_t)W2 * H2);
    const size_t o = (size_t)i * W2 + j, band = (size_t)H2 * W2;
    float d00, d01, d10, d11;
    haar2x2_adjoint(g_ll ? g_ll[pl * band + o] : 0.f, g_yh ? g_yh[(pl * 3 + 0) * band + o] : 0.f,
                    g_yh ? g_yh[(pl * 3 + 1) * band + o] : 0.f, g_yh ? g_yh[(pl * 3 + 2) * band + o] : 0.f, d00, d01,
                    d10, d11);
    float* img = g_x + pl * H * W;
    const int r0 = 2 * i, r1 = 2 * i + 1, c0 = 2 * j, c1 = 2 * j + 1;
    img[(size_t)r0 * W + c0] = d00;
    if (c1 < W) img[(size_t)r0 * W + c1] = d01;
    if (r1 < H) {
        img[(size_t)r1 * W + c0] = d10;
        if (c1 < W) img[(size_t)r1 * W + c1] = d11;
    }
}

}  // namespace lg

using namespace lg;

extern "C" size_t lg_dwt_workspace_bytes(int C, int H, int W, int patch_size) {
    char* p = nullptr;
    DwtWorkspace::from_chunk(p, C, H, W, patch_size);
    return (size_t)p + 128;
}

static int dwt_check(const char* fn, const float* pred, const float* gt, int C, int H, int W, int ps) {
    if (!pred || !gt || C <= 0 || H <= 0 || W <= 0) {
        set_error("%s: invalid image arguments", fn);
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (ps > 0 && (ps & 1)) {
        set_error("%s: patch_size must be even (patch coefficients are read from the global level-1 transform)", fn);
        return LG_ERR_UNSUPPORTED;
    }
    return LG_OK;
}

extern "C" int lg_dwt_loss_forward(const float* pred, const float* gt, int C, int H, int W,
                                   const float* band_weights_host, int patch_size, double percentile, float patch_w_lh,
                                   float patch_w_hl, float* out_losses, uint8_t* patch_mask, char* workspace,
                                   size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int rc = dwt_check("lg_dwt_loss_forward", pred, gt, C, H, W, patch_size);
    if (rc != LG_OK) return rc;
    if (!band_weights_host || !out_losses || !workspace || workspace_bytes < lg_dwt_workspace_bytes(C, H, W, patch_size)) {
        set_error("lg_dwt_loss_forward: missing weights / outputs or workspace too small");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const int ps = patch_size > 0 ? patch_size : 0;
    const int nPH = ps > 0 ? H / ps : 0, nPW = ps > 0 ? W / ps : 0, L = nPH * nPW;
    if (L > 0 && !patch_mask) {
        set_error("lg_dwt_loss_forward: patch_mask is required when the patch term is enabled");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (L > 8192) {
        set_error("lg_dwt_loss_forward: %d patches exceed the supported 8192", L);
        return LG_ERR_UNSUPPORTED;
    }
    char* p = workspace;
    DwtWorkspace w = DwtWorkspace::from_chunk(p, C, H, W, ps);
    const size_t zero_bytes = (size_t)((char*)w.elf_low - (char*)w.band_sums);
    LG_CUDA(cudaMemsetAsync(w.band_sums, 0, zero_bytes, stream));
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2, H4 = (H2 + 1) / 2, W4 = (W2 + 1) / 2;
    const dim3 grid((W4 + DWT_TB - 1) / DWT_TB, (H4 + DWT_TB - 1) / DWT_TB);
    dwt_forward_kernel<<<grid, DWT_TB * DWT_TB, 0, stream>>>(pred, gt, C, H, W, ps, nPW, nPH, w.band_sums, w.patch_sums,
                                                             w.elf_low);
    LG_LAUNCH_CHECK(false, stream);
    if (L > 0) {
        dwt_elf_patch_kernel<<<dim3(L, 8), 256, 0, stream>>>(w.elf_low, H, W, ps, nPW, w.patch_elf);
        LG_LAUNCH_CHECK(false, stream);
    }
    DwtFinalizeArgs a;
    for (int b = 0; b < 8; b++) a.w[b] = band_weights_host[b];
    a.w_lh = patch_w_lh; a.w_hl = patch_w_hl; a.C = C; a.H = H; a.W = W; a.ps = ps; a.L = L;
    // k = clamp(int(L * (1 - percentile)), 1, L) evaluated in double like the Python float arithmetic
    int k = (int)((double)L * (1.0 - percentile));
    k = k < 1 ? 1 : k;
    k = k > L ? L : k;
    a.k = k;
    dwt_finalize_kernel<<<1, 1024, sizeof(float) * (size_t)(L > 0 ? L : 1), stream>>>(a, w.band_sums, w.patch_sums,
                                                                                     w.patch_elf, out_losses, patch_mask);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_dwt_loss_backward(const float* pred, const float* gt, int C, int H, int W,
                                    const float* band_weights_host, int patch_size, float patch_w_lh, float patch_w_hl,
                                    const float* g_dwt_dev, const float* g_patch_dev, const uint8_t* patch_mask,
                                    const float* out_losses, float* dL_dpred, void* stream_v) {
    return lg_dwt_loss_backward_scaled(pred, gt, C, H, W, band_weights_host, patch_size, patch_w_lh, patch_w_hl, g_dwt_dev,
                                       g_patch_dev, nullptr, patch_mask, out_losses, dL_dpred, 0, stream_v);
}

extern "C" int lg_dwt_loss_backward_scaled(const float* pred, const float* gt, int C, int H, int W,
                                           const float* band_weights_host, int patch_size, float patch_w_lh,
                                           float patch_w_hl, const float* g_dwt_dev, const float* g_patch_dev,
                                           const float* g_up_dev, const uint8_t* patch_mask, const float* out_losses,
                                           float* dL_dpred, int accumulate, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int rc = dwt_check("lg_dwt_loss_backward", pred, gt, C, H, W, patch_size);
    if (rc != LG_OK) return rc;
    if (!band_weights_host || !out_losses || !dL_dpred) {
        set_error("lg_dwt_loss_backward: missing weights / forward results / output");
        return LG_ERR_INVALID_ARGUMENT;
    }
    DwtBackwardArgs a;
    for (int b = 0; b < 8; b++) a.w[b] = band_weights_host[b];
    a.w_lh = patch_w_lh; a.w_hl = patch_w_hl; a.C = C; a.H = H; a.W = W;
    a.ps = patch_size > 0 ? patch_size : 0;
    a.nPH = a.ps > 0 ? H / a.ps : 0;
    a.nPW = a.ps > 0 ? W / a.ps : 0;
    if (a.nPH * a.nPW == 0) a.ps = 0;
    if (a.ps > 0 && !patch_mask) {
        set_error("lg_dwt_loss_backward: patch_mask is required when the patch term is enabled");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2, H4 = (H2 + 1) / 2, W4 = (W2 + 1) / 2;
    const dim3 grid((W4 + DWT_TB - 1) / DWT_TB, (H4 + DWT_TB - 1) / DWT_TB);
    dwt_backward_kernel<<<grid, DWT_TB * DWT_TB, 0, stream>>>(pred, gt, a, g_dwt_dev, g_patch_dev, patch_mask, out_losses,
                                                              g_up_dev, accumulate, dL_dpred);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_haar_dwt2_forward(const float* x, int planes, int H, int W, float* ll, float* yh, void* stream_v) {
    if (!x || !ll || !yh || planes <= 0 || H <= 0 || W <= 0) {
        set_error("lg_haar_dwt2_forward: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const size_t n = (size_t)planes * ((H + 1) / 2) * ((W + 1) / 2);
    haar_dwt2_forward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_v>>>(x, planes, H, W, ll, yh);
    LG_LAUNCH_CHECK(false, (cudaStream_t)stream_v);
    return LG_OK;
}

extern "C" int lg_haar_dwt2_backward(const float* g_ll, const float* g_yh, int planes, int H, int W, float* g_x,
                                     void* stream_v) {
    if (!g_x || planes <= 0 || H <= 0 || W <= 0) {
        set_error("lg_haar_dwt2_backward: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const size_t n = (size_t)planes * ((H + 1) / 2) * ((W + 1) / 2);
    haar_dwt2_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_v>>>(g_ll, g_yh, planes, H, W, g_x);
    LG_LAUNCH_CHECK(false, (cudaStream_t)stream_v);
    return LG_OK;
}
