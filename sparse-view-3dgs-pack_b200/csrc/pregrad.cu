// pregrad.cu — per-Gaussian backward: 2-D gradients (packed record written by blend_bwd) -> gradients of means,
// opacity, SH / colours, 3-D covariance, scales and rotations.  One kernel replaces computeCov2DCUDA +
// preprocessCUDA<bwd> (DGR/cuda_rasterizer/backward.cu:147-326, 398-449 with the helpers at :23-142, :330-393)
// and the nine torch::zeros fills of DGR/rasterize_points.cu:163-172: every output element is written here, zeros
// included, so callers hand in uninitialised memory.
#include "common.cuh"

namespace lg {

#define LG_REC 12

struct PreGradArgs {
    int P, D, M, C, W, H;
    const float* __restrict__ means3D;
    const float4* __restrict__ conic_opacity;
    const int* __restrict__ radii;
    const float* __restrict__ shs;
    const uint8_t* __restrict__ clamped;
    const float* __restrict__ opacities;
    const float* __restrict__ scales;
    const float* __restrict__ rotations;
    float scale_modifier;
    const float* __restrict__ cov3Ds;  // precomputed or the forward's
    const float* __restrict__ view;
    const float* __restrict__ proj;
    const float* __restrict__ campos;
    float focal_x, focal_y, tan_fovx, tan_fovy;
    int antialiasing, has_invdepth, sh_path, scale_path, accumulate;
    const float* __restrict__ rec;
    const float* __restrict__ rot_norm;  // non-null: emit gradients w.r.t. the RAW opacity / scaling / rotation
    float* __restrict__ dL_dmean2D;
    float* __restrict__ dL_dconic;
    float* __restrict__ dL_dopacity;
    float* __restrict__ dL_dcolor;
    float* __restrict__ dL_dinvdepth;
    float* __restrict__ dL_dmean3D;
    float* __restrict__ dL_dcov3D;
    float* __restrict__ dL_dsh;
    float* __restrict__ dL_dscale;
    float* __restrict__ dL_drot;
};

__constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                 0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                 -0.5900435899266435f};

// SH coefficients and their gradients are (P, M, 3) arrays: one Gaussian's 3M floats are contiguous, so a thread
// walking its own Gaussian touches one 4-byte word per 32-byte sector per instruction.  Instead every warp moves the
// contiguous 32 x 3M block of its 32 Gaussians with fully coalesced 128-byte accesses through a shared-memory tile
// whose row stride (3M | 1) is odd, so the per-thread walks along a row are bank-conflict free.  The tile first holds
// the coefficients (read by the owner thread), then the gradients (written by it), then the warp streams them out.
// STAGED = false (3M > PG_MAX_ROW, or no SH path) keeps direct global accesses.
#define PG_MAX_ROW 48
#ifndef PG_SH_BULK
#define PG_SH_BULK 1  // 0: keep the register-staged SH tile even where bulk copies apply (A/B measurements)
#endif
#ifndef PG_BULK_PLAIN_STORE
#define PG_BULK_PLAIN_STORE 0
#endif
#ifndef PG_MIN_BLOCKS
#define PG_MIN_BLOCKS 4  // 64 registers: 4 resident blocks per SM measured 4 % faster than 3 at 79 registers
#endif

// SH staging modes as in preprocess.cu: 0 = global memory, 1 = register-staged odd-stride tile per warp, 2 = bulk
// asynchronous copies: the 192-byte coefficient rows of the visible Gaussians come in by TMA (one 384-byte copy per lane
// pair, awaited on a per-warp mbarrier), the gradient rows go out the same way — a plain bulk store, or, when the caller
// accumulates several views, cp.reduce.async.bulk .add.f32, which adds the row into global memory inside the L2
// instead of a load-add-store through registers.
template <int STAGED, int M3C>
__global__ void __launch_bounds__(256, PG_MIN_BLOCKS) preprocess_backward_kernel(PreGradArgs a) {
    extern __shared__ __align__(16) float s_tile_dyn[];
    __shared__ float s_cam[35];
    __shared__ __align__(8) uint64_t s_bar[8];
    if (threadIdx.x < 16) s_cam[threadIdx.x] = __ldg(a.view + threadIdx.x);
    else if (threadIdx.x < 32) s_cam[threadIdx.x] = __ldg(a.proj + threadIdx.x - 16);
    else if (threadIdx.x < 35) s_cam[threadIdx.x] = __ldg(a.campos + threadIdx.x - 32);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int M3 = a.M * 3;
    const int row = M3 | 1;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const size_t warp_first = (size_t)blockIdx.x * blockDim.x + (size_t)warp * 32u;
    float* s_wtile = s_tile_dyn + (size_t)warp * 32u * row;
    int warp_floats = 0;
    bool pair_visible = false;
    if (STAGED == 2 && lane == 0) {
        lg_mbar_init(&s_bar[warp], 32);  // every lane of the warp arrives once
        lg_mbar_init_fence();
    }
    if (STAGED == 1) {
        const long long left = (long long)a.P - (long long)warp_first;
        warp_floats = left <= 0 ? 0 : (int)(left < 32 ? left : 32) * M3;
        lg_warp_rows_to_tile<M3C>(a.shs + warp_first * M3, s_wtile, M3, warp_floats, lane);
    }
    __syncthreads();
    if (STAGED == 2) {
        const bool vis = idx < a.P && a.radii[idx] > 0;
        const bool other_vis = __shfl_xor_sync(0xffffffffu, vis, 1);  // (not inside `||`: every lane must take part)
        pair_visible = vis || other_vis;
        if (pair_visible && (lane & 1u) == 0) {
            const unsigned bytes = (idx + 1 < a.P ? 2u : 1u) * LG_SH_ROW_FLOATS * 4u;
            lg_mbar_arrive_expect_tx(&s_bar[warp], bytes);
            lg_bulk_load(lg_sh_row(s_tile_dyn, threadIdx.x), a.shs + (size_t)idx * LG_SH_ROW_FLOATS, bytes, &s_bar[warp]);
        } else {
            lg_mbar_arrive(&s_bar[warp]);
        }
    }
    // bulk variant: the coefficients of this thread's row, and what its gradient row is made of —
    // dL/dsh[3k + c] = basis_k(direction) * g[c], zero above the active degree — kept as 16 + 3 values until the store
    float sh_regs[STAGED == 2 ? LG_SH_ROW_FLOATS : 1];
    float basis[16], sh_g[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 16; k++) basis[k] = 0.f;
    if (idx < a.P) {
    const float* V = s_cam;
    const float* PR = s_cam + 16;
    const size_t i = (size_t)idx;
    const int C = a.C;
    // SH rows: the staged tile row is read (coefficients) and then overwritten (gradients) by this thread alone
    const float* sh = STAGED == 2 ? sh_regs : (STAGED == 1 ? s_wtile + lane * row : a.shs + i * M3);
    float* out_sh = STAGED == 1 ? s_wtile + lane * row : a.dL_dsh + i * M3;  // unused by the bulk variant

    const bool visible = a.radii[idx] > 0;
    float r[LG_REC];
    {
        const float4* rp = reinterpret_cast<const float4*>(a.rec + i * LG_REC);
        const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
        r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
        r[8] = r2.x; r[9] = r2.y; r[10] = r2.z; r[11] = r2.w;
    }
    // The blend backward accumulates the moments S = sum over hits of w*(dx, dy, dx^2, dx*dy, dy^2, 1) with
    // w = G*dL/dalpha and d = mean2D - pixel; the reference's per-hit expressions (backward.cu:598-632) are linear
    // in them: dL/dmean2D = -o*(a*Sx + b*Sy, c*Sy + b*Sx) * (W/2, H/2), dL/dconic = -o/2 * (Sxx, Sxy, Syy),
    // dL/dopacity = S1.  Invisible Gaussians carry an all-zero record.
    if (visible) {
        const float4 co = a.conic_opacity[i];
        const float sx = r[0], sy = r[1];
        r[0] = -co.w * (co.x * sx + co.y * sy) * (0.5f * (float)a.W);
        r[1] = -co.w * (co.z * sy + co.y * sx) * (0.5f * (float)a.H);
        const float k = -0.5f * co.w;
        r[2] *= k;
        r[3] *= k;
        r[4] *= k;
    }
    // pass-through outputs (blend-stage gradients)
    a.dL_dmean2D[3 * i + 0] = r[0];
    a.dL_dmean2D[3 * i + 1] = r[1];
    a.dL_dmean2D[3 * i + 2] = 0.0f;
    if (a.dL_dconic) reinterpret_cast<float4*>(a.dL_dconic)[i] = make_float4(r[2], r[3], 0.0f, r[4]);
    if (a.dL_dcolor)  // only a colors_precomp caller needs the per-Gaussian colour gradient in memory
        for (int c = 0; c < C; c++) a.dL_dcolor[i * C + c] = r[7 + c];
    if (a.dL_dinvdepth) a.dL_dinvdepth[i] = r[6];

    float dmean[3] = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float dopacity = r[5];

    if (visible) {
        const float mx = a.means3D[3 * i + 0], my = a.means3D[3 * i + 1], mz = a.means3D[3 * i + 2];
        float c3[6];
#pragma unroll
        for (int k = 0; k < 6; k++) c3[k] = a.cov3Ds[6 * i + k];

        // ---------------- computeCov2DCUDA (backward.cu:147-326)
        const float dcon_x = r[2], dcon_y = r[3], dcon_z = r[4];
        float tx = V[0] * mx + V[4] * my + V[8] * mz + V[12];
        float ty = V[1] * mx + V[5] * my + V[9] * mz + V[13];
        const float tz = V[2] * mx + V[6] * my + V[10] * mz + V[14];
        const float limx = 1.3f * a.tan_fovx, limy = 1.3f * a.tan_fovy;
        const float txtz = tx / tz, tytz = ty / tz;
        tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
        ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        const float hx = a.focal_x, hy = a.focal_y;
        const float Ja = hx / tz, Jb = -(hx * tx) / (tz * tz), Jc = hy / tz, Jd = -(hy * ty) / (tz * tz);
        // glm-style [col][row]
        const float Wm[3][3] = {{V[0], V[4], V[8]}, {V[1], V[5], V[9]}, {V[2], V[6], V[10]}};
        float Tm[2][3];
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
            Tm[0][rr] = Wm[0][rr] * Ja + Wm[2][rr] * Jb;
            Tm[1][rr] = Wm[1][rr] * Jc + Wm[2][rr] * Jd;
        }
        const float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        // TV[a][m] = sum_k T[a][k] * Vrk[k][m]
        float TV[2][3];
#pragma unroll
        for (int aa = 0; aa < 2; aa++)
#pragma unroll
            for (int m = 0; m < 3; m++)
                TV[aa][m] = Tm[aa][0] * Vrk[0][m] + Tm[aa][1] * Vrk[1][m] + Tm[aa][2] * Vrk[2][m];
        float c_xx = TV[0][0] * Tm[0][0] + TV[0][1] * Tm[0][1] + TV[0][2] * Tm[0][2];
        float c_xy = TV[1][0] * Tm[0][0] + TV[1][1] * Tm[0][1] + TV[1][2] * Tm[0][2];
        float c_yy = TV[1][0] * Tm[1][0] + TV[1][1] * Tm[1][1] + TV[1][2] * Tm[1][2];

        const float h_var = 0.3f;
        float d_inside_root = 0.f;
        if (a.antialiasing) {
            const float det_cov = c_xx * c_yy - c_xy * c_xy;
            c_xx += h_var;
            c_yy += h_var;
            const float det_cov_plus_h = c_xx * c_yy - c_xy * c_xy;
            const float h_scaling = sqrtf(fmaxf(0.000025f, det_cov / det_cov_plus_h));
            const float d_h_scaling = dopacity * a.opacities[idx];
            d_inside_root = (det_cov / det_cov_plus_h) <= 0.000025f ? 0.f : d_h_scaling / (2.f * h_scaling);
            dopacity = dopacity * h_scaling;
        } else {
            c_xx += h_var;
            c_yy += h_var;
        }
        float dL_dc_xx = 0.f, dL_dc_xy = 0.f, dL_dc_yy = 0.f;
        if (a.antialiasing) {
            const float x = c_xx, y = c_yy, z = c_xy, w = h_var;
            const float sq = w * w + w * (x + y) + x * y - z * z;
            const float denom_f = d_inside_root / (sq * sq);
            dL_dc_xx = w * (w * y + y * y + z * z) * denom_f;
            dL_dc_yy = w * (w * x + x * x + z * z) * denom_f;
            dL_dc_xy = -2.f * w * z * (w + x + y) * denom_f;
        }
        const float denom = c_xx * c_yy - c_xy * c_xy;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        if (denom2inv != 0.f) {
            dL_dc_xx += denom2inv * (-c_yy * c_yy * dcon_x + 2.f * c_xy * c_yy * dcon_y + (denom - c_xx * c_yy) * dcon_z);
            dL_dc_yy += denom2inv * (-c_xx * c_xx * dcon_z + 2.f * c_xx * c_xy * dcon_y + (denom - c_xx * c_yy) * dcon_x);
            dL_dc_xy += denom2inv * 2.f * (c_xy * c_yy * dcon_x - (denom + 2.f * c_xy * c_xy) * dcon_y + c_xx * c_xy * dcon_z);
            dcov[0] = Tm[0][0] * Tm[0][0] * dL_dc_xx + Tm[0][0] * Tm[1][0] * dL_dc_xy + Tm[1][0] * Tm[1][0] * dL_dc_yy;
            dcov[3] = Tm[0][1] * Tm[0][1] * dL_dc_xx + Tm[0][1] * Tm[1][1] * dL_dc_xy + Tm[1][1] * Tm[1][1] * dL_dc_yy;
            dcov[5] = Tm[0][2] * Tm[0][2] * dL_dc_xx + Tm[0][2] * Tm[1][2] * dL_dc_xy + Tm[1][2] * Tm[1][2] * dL_dc_yy;
            dcov[1] = 2.f * Tm[0][0] * Tm[0][1] * dL_dc_xx + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_dc_xy +
                      2.f * Tm[1][0] * Tm[1][1] * dL_dc_yy;
            dcov[2] = 2.f * Tm[0][0] * Tm[0][2] * dL_dc_xx + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_dc_xy +
                      2.f * Tm[1][0] * Tm[1][2] * dL_dc_yy;
            dcov[4] = 2.f * Tm[0][2] * Tm[0][1] * dL_dc_xx + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_dc_xy +
                      2.f * Tm[1][1] * Tm[1][2] * dL_dc_yy;
        }
        // dL/dT (upper 2x3): TV[a][m] is exactly the (T[a][.] . Vrk[m][.]) dot product the reference spells out
        const float dL_dT00 = 2.f * TV[0][0] * dL_dc_xx + TV[1][0] * dL_dc_xy;
        const float dL_dT01 = 2.f * TV[0][1] * dL_dc_xx + TV[1][1] * dL_dc_xy;
        const float dL_dT02 = 2.f * TV[0][2] * dL_dc_xx + TV[1][2] * dL_dc_xy;
        const float dL_dT10 = 2.f * TV[1][0] * dL_dc_yy + TV[0][0] * dL_dc_xy;
        const float dL_dT11 = 2.f * TV[1][1] * dL_dc_yy + TV[0][1] * dL_dc_xy;
        const float dL_dT12 = 2.f * TV[1][2] * dL_dc_yy + TV[0][2] * dL_dc_xy;
        const float dL_dJ00 = Wm[0][0] * dL_dT00 + Wm[0][1] * dL_dT01 + Wm[0][2] * dL_dT02;
        const float dL_dJ02 = Wm[2][0] * dL_dT00 + Wm[2][1] * dL_dT01 + Wm[2][2] * dL_dT02;
        const float dL_dJ11 = Wm[1][0] * dL_dT10 + Wm[1][1] * dL_dT11 + Wm[1][2] * dL_dT12;
        const float dL_dJ12 = Wm[2][0] * dL_dT10 + Wm[2][1] * dL_dT11 + Wm[2][2] * dL_dT12;
        const float itz = 1.f / tz, itz2 = itz * itz, itz3 = itz2 * itz;
        const float dL_dtx = x_grad_mul * -hx * itz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -hy * itz2 * dL_dJ12;
        float dL_dtz = -hx * itz2 * dL_dJ00 - hy * itz2 * dL_dJ11 + (2.f * hx * tx) * itz3 * dL_dJ02 +
                       (2.f * hy * ty) * itz3 * dL_dJ12;
        if (a.has_invdepth) dL_dtz -= r[6] / (tz * tz);
        // transformVec4x3Transpose
        dmean[0] = V[0] * dL_dtx + V[1] * dL_dty + V[2] * dL_dtz;
        dmean[1] = V[4] * dL_dtx + V[5] * dL_dty + V[6] * dL_dtz;
        dmean[2] = V[8] * dL_dtx + V[9] * dL_dty + V[10] * dL_dtz;

        // ---------------- preprocessCUDA backward (backward.cu:398-449): screen-space mean -> 3-D mean
        {
            const float m_hw = PR[3] * mx + PR[7] * my + PR[11] * mz + PR[15];
            const float m_w = 1.0f / (m_hw + 0.0000001f);
            const float mul1 = (PR[0] * mx + PR[4] * my + PR[8] * mz + PR[12]) * m_w * m_w;
            const float mul2 = (PR[1] * mx + PR[5] * my + PR[9] * mz + PR[13]) * m_w * m_w;
            const float gx = r[0], gy = r[1];
            dmean[0] += (PR[0] * m_w - PR[3] * mul1) * gx + (PR[1] * m_w - PR[3] * mul2) * gy;
            dmean[1] += (PR[4] * m_w - PR[7] * mul1) * gx + (PR[5] * m_w - PR[7] * mul2) * gy;
            dmean[2] += (PR[8] * m_w - PR[11] * mul1) * gx + (PR[9] * m_w - PR[11] * mul2) * gy;
        }
    }

    // ---------------- SH backward (backward.cu:23-142).  Every coefficient is read before its slot is written:
    // in the staged variant `sh` and `out_sh` are the same shared-memory row.
    if (a.sh_path) {
        if (!visible) {
            if (STAGED != 2)
                for (int k = 0; k < M3; k++) out_sh[k] = 0.f;
        } else {
            if (STAGED == 2) {  // the row has arrived (or arrives now): 12 conflict-free float4 reads
                lg_mbar_wait(&s_bar[warp], 0);
                const float4* row4 = reinterpret_cast<const float4*>(lg_sh_row(s_tile_dyn, threadIdx.x));
#pragma unroll
                for (int j = 0; j < LG_SH_ROW_FLOATS / 4; j++) {
                    const float4 v = row4[j];
                    sh_regs[4 * j + 0] = v.x; sh_regs[4 * j + 1] = v.y; sh_regs[4 * j + 2] = v.z; sh_regs[4 * j + 3] = v.w;
                }
            }
            const float mx = a.means3D[3 * i + 0], my = a.means3D[3 * i + 1], mz = a.means3D[3 * i + 2];
            const float ox = mx - V[32], oy = my - V[33], oz = mz - V[34];
            const float ilen = 1.0f / sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox * ilen, y = oy * ilen, z = oz * ilen;
            float g[3];
#pragma unroll
            for (int c = 0; c < 3; c++) g[c] = a.clamped[3 * i + c] ? 0.f : r[7 + c];
            float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0};
            const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
            const int deg = a.D;
            int written = 1;
            basis[0] = C0;
#pragma unroll
            for (int c = 0; c < 3; c++) if (STAGED != 2) out_sh[c] = C0 * g[c];
            if (deg > 0) {
                written = 4;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float s1 = sh[3 + c], s2 = sh[6 + c], s3 = sh[9 + c];
                    dx[c] = -C1 * s3;
                    dy[c] = -C1 * s1;
                    dz[c] = C1 * s2;
                    if (STAGED != 2) {
                        out_sh[3 + c] = -C1 * y * g[c];
                        out_sh[6 + c] = C1 * z * g[c];
                        out_sh[9 + c] = -C1 * x * g[c];
                    }
                }
                basis[1] = -C1 * y; basis[2] = C1 * z; basis[3] = -C1 * x;
                if (deg > 1) {
                    written = 9;
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    const float b4 = c_SH_C2[0] * xy, b5 = c_SH_C2[1] * yz, b6 = c_SH_C2[2] * (2.f * zz - xx - yy),
                                b7 = c_SH_C2[3] * xz, b8 = c_SH_C2[4] * (xx - yy);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float s4 = sh[12 + c], s5 = sh[15 + c], s6 = sh[18 + c], s7 = sh[21 + c], s8 = sh[24 + c];
                        dx[c] += c_SH_C2[0] * y * s4 + c_SH_C2[2] * 2.f * -x * s6 + c_SH_C2[3] * z * s7 + c_SH_C2[4] * 2.f * x * s8;
                        dy[c] += c_SH_C2[0] * x * s4 + c_SH_C2[1] * z * s5 + c_SH_C2[2] * 2.f * -y * s6 + c_SH_C2[4] * 2.f * -y * s8;
                        dz[c] += c_SH_C2[1] * y * s5 + c_SH_C2[2] * 2.f * 2.f * z * s6 + c_SH_C2[3] * x * s7;
                        if (STAGED != 2) {
                            out_sh[12 + c] = b4 * g[c];
                            out_sh[15 + c] = b5 * g[c];
                            out_sh[18 + c] = b6 * g[c];
                            out_sh[21 + c] = b7 * g[c];
                            out_sh[24 + c] = b8 * g[c];
                        }
                    }
                    basis[4] = b4; basis[5] = b5; basis[6] = b6; basis[7] = b7; basis[8] = b8;
                    if (deg > 2) {
                        written = 16;
                        const float b9 = c_SH_C3[0] * y * (3.f * xx - yy), b10 = c_SH_C3[1] * xy * z,
                                    b11 = c_SH_C3[2] * y * (4.f * zz - xx - yy),
                                    b12 = c_SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy),
                                    b13 = c_SH_C3[4] * x * (4.f * zz - xx - yy), b14 = c_SH_C3[5] * z * (xx - yy),
                                    b15 = c_SH_C3[6] * x * (xx - 3.f * yy);
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const float s9 = sh[27 + c], s10 = sh[30 + c], s11 = sh[33 + c], s12 = sh[36 + c],
                                        s13 = sh[39 + c], s14 = sh[42 + c], s15 = sh[45 + c];
                            dx[c] += c_SH_C3[0] * s9 * 3.f * 2.f * xy + c_SH_C3[1] * s10 * yz + c_SH_C3[2] * s11 * -2.f * xy +
                                     c_SH_C3[3] * s12 * -3.f * 2.f * xz + c_SH_C3[4] * s13 * (-3.f * xx + 4.f * zz - yy) +
                                     c_SH_C3[5] * s14 * 2.f * xz + c_SH_C3[6] * s15 * 3.f * (xx - yy);
                            dy[c] += c_SH_C3[0] * s9 * 3.f * (xx - yy) + c_SH_C3[1] * s10 * xz +
                                     c_SH_C3[2] * s11 * (-3.f * yy + 4.f * zz - xx) + c_SH_C3[3] * s12 * -3.f * 2.f * yz +
                                     c_SH_C3[4] * s13 * -2.f * xy + c_SH_C3[5] * s14 * -2.f * yz +
                                     c_SH_C3[6] * s15 * -3.f * 2.f * xy;
                            dz[c] += c_SH_C3[1] * s10 * xy + c_SH_C3[2] * s11 * 4.f * 2.f * yz +
                                     c_SH_C3[3] * s12 * 3.f * (2.f * zz - xx - yy) + c_SH_C3[4] * s13 * 4.f * 2.f * xz +
                                     c_SH_C3[5] * s14 * (xx - yy);
                            if (STAGED != 2) {
                                out_sh[27 + c] = b9 * g[c];
                                out_sh[30 + c] = b10 * g[c];
                                out_sh[33 + c] = b11 * g[c];
                                out_sh[36 + c] = b12 * g[c];
                                out_sh[39 + c] = b13 * g[c];
                                out_sh[42 + c] = b14 * g[c];
                                out_sh[45 + c] = b15 * g[c];
                            }
                        }
                        basis[9] = b9; basis[10] = b10; basis[11] = b11; basis[12] = b12; basis[13] = b13;
                        basis[14] = b14; basis[15] = b15;
                    }
                }
            }
            if (STAGED != 2)
                for (int k = written * 3; k < M3; k++) out_sh[k] = 0.f;  // coefficients above the active degree
            sh_g[0] = g[0]; sh_g[1] = g[1]; sh_g[2] = g[2];
            const float ddir_x = dx[0] * g[0] + dx[1] * g[1] + dx[2] * g[2];
            const float ddir_y = dy[0] * g[0] + dy[1] * g[1] + dy[2] * g[2];
            const float ddir_z = dz[0] * g[0] + dz[1] * g[1] + dz[2] * g[2];
            // dnormvdv (auxiliary.h:119-129)
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((+sum2 - ox * ox) * ddir_x - oy * ox * ddir_y - oz * ox * ddir_z) * invsum32;
            dmean[1] += (-ox * oy * ddir_x + (sum2 - oy * oy) * ddir_y - oz * oy * ddir_z) * invsum32;
            dmean[2] += (-ox * oz * ddir_x - oy * oz * ddir_y + (sum2 - oz * oz) * ddir_z) * invsum32;
        }
    }

    // ---------------- scale / rotation backward (backward.cu:330-393)
    if (a.scale_path) {
        float ds[3] = {0, 0, 0};
        float dq[4] = {0, 0, 0, 0};
        if (visible) {
            const float4 q = reinterpret_cast<const float4*>(a.rotations)[i];
            const float qr = q.x, qx = q.y, qy = q.z, qz = q.w;
            const float R[3][3] = {
                {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qr * qz), 2.f * (qx * qz + qr * qy)},
                {2.f * (qx * qy + qr * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qr * qx)},
                {2.f * (qx * qz - qr * qy), 2.f * (qy * qz + qr * qx), 1.f - 2.f * (qx * qx + qy * qy)}};
            const float s[3] = {a.scale_modifier * a.scales[3 * i + 0], a.scale_modifier * a.scales[3 * i + 1],
                                a.scale_modifier * a.scales[3 * i + 2]};
            float Mm[3][3];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int rr = 0; rr < 3; rr++) Mm[c][rr] = s[rr] * R[c][rr];
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            // dL_dM = 2 * M * dL_dSigma ; dL_dMt = transpose(dL_dM): dMt[c][r] = dM[r][c]
            float dMt[3][3];
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int rr = 0; rr < 3; rr++)
                    dMt[rr][c] = 2.0f * (Mm[0][rr] * dS[c][0] + Mm[1][rr] * dS[c][1] + Mm[2][rr] * dS[c][2]);
            // Rt[c][r] = R[r][c]; dL_dscale[k] = dot(Rt[k], dMt[k])
#pragma unroll
            for (int k = 0; k < 3; k++) ds[k] = R[0][k] * dMt[k][0] + R[1][k] * dMt[k][1] + R[2][k] * dMt[k][2];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                dMt[k][0] *= s[k];
                dMt[k][1] *= s[k];
                dMt[k][2] *= s[k];
            }
            dq[0] = 2 * qz * (dMt[0][1] - dMt[1][0]) + 2 * qy * (dMt[2][0] - dMt[0][2]) + 2 * qx * (dMt[1][2] - dMt[2][1]);
            dq[1] = 2 * qy * (dMt[1][0] + dMt[0][1]) + 2 * qz * (dMt[2][0] + dMt[0][2]) + 2 * qr * (dMt[1][2] - dMt[2][1]) -
                    4 * qx * (dMt[2][2] + dMt[1][1]);
            dq[2] = 2 * qx * (dMt[1][0] + dMt[0][1]) + 2 * qr * (dMt[2][0] - dMt[0][2]) + 2 * qz * (dMt[1][2] + dMt[2][1]) -
                    4 * qy * (dMt[2][2] + dMt[0][0]);
            dq[3] = 2 * qr * (dMt[0][1] - dMt[1][0]) + 2 * qx * (dMt[2][0] + dMt[0][2]) + 2 * qy * (dMt[1][2] + dMt[2][1]) -
                    4 * qz * (dMt[1][1] + dMt[0][0]);
        }
        if (a.rot_norm != nullptr && visible) {
            // chain rule of the activations of LG/scene/gaussian_model.py:36-50 (the reference leaves it to autograd):
            // scales = exp(raw) -> d/draw = d/dscale * scale; rotations = raw / max(|raw|, 1e-12) ->
            // d/draw = (g - q (q . g)) / max(|raw|, 1e-12) with q the unit quaternion the forward was given
            const float4 q = reinterpret_cast<const float4*>(a.rotations)[i];
#pragma unroll
            for (int k = 0; k < 3; k++) ds[k] *= a.scales[3 * i + k];
            const float dot = q.x * dq[0] + q.y * dq[1] + q.z * dq[2] + q.w * dq[3];
            const float inv_n = 1.0f / fmaxf(a.rot_norm[i], 1e-12f);
            dq[0] = (dq[0] - q.x * dot) * inv_n;
            dq[1] = (dq[1] - q.y * dot) * inv_n;
            dq[2] = (dq[2] - q.z * dot) * inv_n;
            dq[3] = (dq[3] - q.w * dot) * inv_n;
        }
        if (a.accumulate) {
            const float4 old = reinterpret_cast<const float4*>(a.dL_drot)[i];
            dq[0] += old.x; dq[1] += old.y; dq[2] += old.z; dq[3] += old.w;
#pragma unroll
            for (int k = 0; k < 3; k++) ds[k] += a.dL_dscale[3 * i + k];
        }
        a.dL_dscale[3 * i + 0] = ds[0];
        a.dL_dscale[3 * i + 1] = ds[1];
        a.dL_dscale[3 * i + 2] = ds[2];
        reinterpret_cast<float4*>(a.dL_drot)[i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
    }

    if (a.rot_norm != nullptr) {  // opacities = sigmoid(raw) -> d/draw = d/dopacity * o (1 - o)
        const float o = a.opacities[i];
        dopacity *= o * (1.0f - o);
    }
    if (a.accumulate) {
        dopacity += a.dL_dopacity[i];
#pragma unroll
        for (int k = 0; k < 3; k++) dmean[k] += a.dL_dmean3D[3 * i + k];
    }
    a.dL_dopacity[i] = dopacity;
    a.dL_dmean3D[3 * i + 0] = dmean[0];
    a.dL_dmean3D[3 * i + 1] = dmean[1];
    a.dL_dmean3D[3 * i + 2] = dmean[2];
    if (a.dL_dcov3D) {  // only a cov3D_precomp caller needs it in memory
#pragma unroll
        for (int k = 0; k < 6; k++) a.dL_dcov3D[6 * i + k] = dcov[k];
    }
    }  // idx < P

    if (STAGED == 1) {  // stream the warp's 32 x 3M gradient block out with coalesced stores
        __syncwarp();
        lg_warp_tile_to_rows<M3C>(a.dL_dsh + warp_first * M3, s_wtile, M3, warp_floats, lane, a.accumulate != 0);
    }
    if (STAGED == 2) {
        // gradient rows: registers -> the thread's shared-memory row (over the coefficients, which every lane of the
        // warp has read by now: the mbarrier wait below also covers warps whose lanes never needed them) -> global memory
        // by one bulk copy per lane pair.  Accumulation adds inside the L2 (cp.reduce.async.bulk); pairs without a
        // visible Gaussian contribute nothing then and are skipped.
        lg_mbar_wait(&s_bar[warp], 0);
        const bool write_pair = a.accumulate ? pair_visible : true;
        if (idx < a.P && write_pair) {
            float4* row4 = reinterpret_cast<float4*>(lg_sh_row(s_tile_dyn, threadIdx.x));
#pragma unroll
            for (int j = 0; j < LG_SH_ROW_FLOATS / 4; j++)  // element e = 4 j + q is coefficient e / 3, channel e % 3
                row4[j] = make_float4(basis[(4 * j) / 3] * sh_g[(4 * j) % 3], basis[(4 * j + 1) / 3] * sh_g[(4 * j + 1) % 3],
                                      basis[(4 * j + 2) / 3] * sh_g[(4 * j + 2) % 3], basis[(4 * j + 3) / 3] * sh_g[(4 * j + 3) % 3]);
        }
        lg_fence_proxy_async();
        __syncwarp();
#if PG_BULK_PLAIN_STORE
        if (idx < a.P && write_pair) {  // diagnostic variant: rows written back with ordinary stores
            const float4* src4 = reinterpret_cast<const float4*>(lg_sh_row(s_tile_dyn, threadIdx.x));
            float4* dst4 = reinterpret_cast<float4*>(a.dL_dsh + (size_t)idx * LG_SH_ROW_FLOATS);
            for (int j = 0; j < LG_SH_ROW_FLOATS / 4; j++) {
                float4 v = src4[j];
                if (a.accumulate) { const float4 o = dst4[j]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                dst4[j] = v;
            }
        }
        if (false) {
            const unsigned bytes = 0; float* dst = nullptr;
#else
        if (idx < a.P && write_pair && (lane & 1u) == 0) {
            const unsigned bytes = (idx + 1 < a.P ? 2u : 1u) * LG_SH_ROW_FLOATS * 4u;
            float* dst = a.dL_dsh + (size_t)idx * LG_SH_ROW_FLOATS;
#endif
            if (a.accumulate) lg_bulk_reduce_add_f32(dst, lg_sh_row(s_tile_dyn, threadIdx.x), bytes);
            else lg_bulk_store(dst, lg_sh_row(s_tile_dyn, threadIdx.x), bytes);
            lg_bulk_commit();
            lg_bulk_wait_read();  // the rows must stay in place until the copy engine has read them
        }
    }
}

int launch_preprocess_backward(const BackwardArgs& b, const GeometryState& g, const int* radii, bool debug,
                               cudaStream_t stream) {
    PreGradArgs a;
    a.P = b.P; a.D = b.D; a.M = b.M; a.C = b.C;
    a.W = b.W; a.H = b.H; a.means3D = b.means3D; a.conic_opacity = g.conic_opacity; a.radii = radii; a.shs = b.shs; a.clamped = g.clamped; a.opacities = b.opacities;
    a.scales = b.scales; a.rotations = b.rotations; a.scale_modifier = b.scale_modifier;
    a.cov3Ds = b.cov3D_precomp ? b.cov3D_precomp : g.cov3D;
    a.view = b.viewmatrix; a.proj = b.projmatrix; a.campos = b.campos;
    a.focal_x = b.focal_x; a.focal_y = b.focal_y; a.tan_fovx = b.tan_fovx; a.tan_fovy = b.tan_fovy;
    a.antialiasing = b.antialiasing; a.has_invdepth = b.has_invdepth; a.accumulate = b.accumulate ? 1 : 0;
    a.sh_path = (b.shs != nullptr && b.dL_dsh != nullptr && b.M > 0) ? 1 : 0;
    a.scale_path = (b.scales != nullptr && b.rotations != nullptr) ? 1 : 0;
    a.rec = g.grad_scratch;
    a.rot_norm = (a.scale_path ? b.raw_rot_norm : nullptr);
    a.dL_dmean2D = b.dL_dmean2D; a.dL_dconic = b.dL_dconic; a.dL_dopacity = b.dL_dopacity; a.dL_dcolor = b.dL_dcolor;
    a.dL_dinvdepth = b.dL_dinvdepth; a.dL_dmean3D = b.dL_dmean3D; a.dL_dcov3D = b.dL_dcov3D; a.dL_dsh = b.dL_dsh;
    a.dL_dscale = b.dL_dscale; a.dL_drot = b.dL_drot;
    const int M3 = 3 * b.M;
    if (a.sh_path && M3 <= PG_MAX_ROW) {
        const size_t smem = sizeof(float) * 256 * (size_t)(M3 | 1);
        if (M3 == LG_SH_ROW_FLOATS && PG_SH_BULK && ((reinterpret_cast<uintptr_t>(b.shs) | reinterpret_cast<uintptr_t>(b.dL_dsh)) & 15u) == 0) {
            const size_t bulk_smem = sizeof(float) * 128 * (size_t)LG_SH_PAIR_FLOATS;
            LG_CUDA(cudaFuncSetAttribute(preprocess_backward_kernel<2, 48>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_smem));
            preprocess_backward_kernel<2, 48><<<(b.P + 255) / 256, 256, bulk_smem, stream>>>(a);
        } else if (M3 == 48) {
            LG_CUDA(cudaFuncSetAttribute(preprocess_backward_kernel<1, 48>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            preprocess_backward_kernel<1, 48><<<(b.P + 255) / 256, 256, smem, stream>>>(a);
        } else {
            LG_CUDA(cudaFuncSetAttribute(preprocess_backward_kernel<1, 0>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(float) * 256 * (PG_MAX_ROW | 1))));
            preprocess_backward_kernel<1, 0><<<(b.P + 255) / 256, 256, smem, stream>>>(a);
        }
    } else {
        if (a.sh_path && a.accumulate) {
            set_error("gradient accumulation needs SH rows of at most %d floats (got %d)", PG_MAX_ROW, M3);
            return LG_ERR_UNSUPPORTED;
        }
        preprocess_backward_kernel<0, 0><<<(b.P + 255) / 256, 256, 0, stream>>>(a);
    }
    LG_LAUNCH_CHECK(debug, stream);
    return LG_OK;
}

}  // namespace lg
