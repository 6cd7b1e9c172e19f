// preprocess.cu — per-Gaussian forward stage: frustum cull, EWA projection, 2-D covariance / conic, SH -> RGB,
// radius and tile rectangle (the reference's preprocessCUDA, DGR/cuda_rasterizer/forward.cu:151-269), fused with the
// first step of the binning: every Gaussian adds 1 to the list length of each tile its rectangle covers
// (csrc/binning.cu; takes the place of cub::DeviceScan::InclusiveSum over tiles_touched, rasterizer_impl.cu:280).
//
// Numerical contract: every quantity that decides radii / tile rectangles / depth bits is evaluated with
// explicitly rounded intrinsics in exactly the operation order of the reference's sm_100a SASS (nvcc 12.9 -O3):
// the NVVM front end contracts some a*b+c into fma.rn in the PTX and ptxas then fuses further mul/add pairs, so the
// SASS, not the source or the PTX, is the contract (DESIGN.md "bit-exact contract").  Comments "fNNN" / "Rnn" name
// the PTX / SASS registers of that build.
#include "common.cuh"

namespace lg {

#ifndef PRE_BLOCK
#define PRE_BLOCK 256
#endif
#define PRE_MAX_ROW 48  // widest SH row (floats per Gaussian) staged through shared memory
#define PRE_COOP 8      // tile rectangles above this many tiles are counted by the whole warp
#ifndef PRE_SH_BULK
#define PRE_SH_BULK 1   // 0: keep the register-staged SH tile even where bulk copies apply (A/B measurements)
#endif

struct PreArgs {
    int P, D, M, C;
    const float* __restrict__ means3D;
    const float* __restrict__ scales;
    float scale_modifier;
    const float* __restrict__ rotations;
    const float* __restrict__ opacities;
    const float* __restrict__ shs;
    const float* __restrict__ cov3D_precomp;
    const float* __restrict__ colors_precomp;
    const float* __restrict__ viewmatrix;
    const float* __restrict__ projmatrix;
    const float* __restrict__ cam_pos;
    int W, H;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int grid_x, grid_y;
    int prefiltered, antialiasing;
    // outputs
    int* __restrict__ radii;
    float2* __restrict__ means2D;
    float* __restrict__ depths;
    float* __restrict__ cov3Ds;
    float* __restrict__ rgb;
    float4* __restrict__ conic_opacity;
    uint8_t* __restrict__ clamped;
    uint32_t* __restrict__ tiles_touched;
    uint32_t* __restrict__ rect_packed;
    uint32_t* tile_ctr;    // per-tile counter lines (zeroed by the launcher): word 0 = list length, accumulated here
};

// transformPoint4x3 / 4x4 row (DGR/cuda_rasterizer/auxiliary.h:70-89): m0*x + m4*y + m8*z + m12 compiles to
// mul(y,m4); fma(x,m0,.); fma(z,m8,.); add(m12,.)
__device__ __forceinline__ float xform_row(float x, float y, float z, float m0, float m4, float m8, float m12) {
    return F_ADD(m12, F_FMA(z, m8, F_FMA(x, m0, F_MUL(y, m4))));
}

// a*b + c*d + e*f in the order the reference's glm products contract to: mul(c,d); fma(a,b,.); fma(e,f,.)
__device__ __forceinline__ float dot3_mid_first(float a, float b, float c, float d, float e, float f) {
    return F_FMA(e, f, F_FMA(a, b, F_MUL(c, d)));
}

// ndc2Pix (auxiliary.h:40-43): double-precision literals => evaluated in fp64, rounded once.
__device__ __forceinline__ float ndc2pix(float v, int S) {
    double t = __dadd_rn((double)v, 1.0);
    t = __fma_rn(t, (double)S, -1.0);
    t = __dmul_rn(t, 0.5);
    return __double2float_rn(t);
}

// computeCov3D (forward.cu:114-148): Sigma = (S R)^T (S R) with glm column-major products; the x*0 terms of
// the diagonal S are kept so that non-finite inputs propagate exactly as in the reference.
__device__ __forceinline__ void compute_cov3d(float sx0, float sy0, float sz0, float mod, float qr, float qx, float qy,
                                              float qz, float* cov) {
    const float sx = F_MUL(mod, sx0), sy = F_MUL(mod, sy0), sz = F_MUL(mod, sz0);
    // Two-product sums are single FFMAs in the reference SASS (ptxas fuses the PTX mul+add pairs), squares that are
    // reused stay separate multiplies.  "Rnn" = SASS register of the reference build.
    const float xz = F_MUL(qx, qz);                 // R21
    const float rx = F_MUL(qr, qx);                 // R29
    const float xz_p_ry = F_FMA(qr, qy, xz);        // R10
    const float xz_m_ry = F_FMA(-qr, qy, xz);       // R21'
    const float yz_m_rx = F_FMA(qy, qz, -rx);       // R11
    const float yy = F_MUL(qy, qy);                 // R20
    const float rz = F_MUL(qr, qz);                 // R12
    const float yz_p_rx = F_FMA(qy, qz, rx);        // R29'
    const float zz = F_MUL(qz, qz);                 // R15
    const float xx_yy = F_FMA(qx, qx, yy);          // R30
    const float xy_m_rz = F_FMA(qx, qy, -rz);       // R8
    const float xy_p_rz = F_FMA(qx, qy, rz);        // R14
    const float yy_zz = F_ADD(yy, zz);              // R20'
    const float xx_zz = F_FMA(qx, qx, zz);          // R13
    const float R22 = F_SUB(1.0f, F_ADD(xx_yy, xx_yy));     // f131
    const float R21 = F_ADD(yz_p_rx, yz_p_rx);              // f132
    const float R20 = F_ADD(xz_m_ry, xz_m_ry);              // f133
    const float R12 = F_ADD(yz_m_rx, yz_m_rx);              // f134
    const float R11 = F_SUB(1.0f, F_ADD(xx_zz, xx_zz));     // f136
    const float R10 = F_ADD(xy_p_rz, xy_p_rz);              // f137
    const float R02 = F_ADD(xz_p_ry, xz_p_ry);              // f138
    const float R01 = F_ADD(xy_m_rz, xy_m_rz);              // f139
    const float R00 = F_SUB(1.0f, F_ADD(yy_zz, yy_zz));     // f141
    // M = S * R (glm): M[c][r] = S[0][r]*R[c][0] + S[1][r]*R[c][1] + S[2][r]*R[c][2]
    const float M00 = F_FMA(R02, 0.0f, F_FMA(R01, 0.0f, F_MUL(sx, R00)));            // f144
    const float z00 = F_MUL(R00, 0.0f);                                              // f145
    const float M01 = F_FMA(R02, 0.0f, F_FMA(sy, R01, z00));                         // f147
    const float M02 = F_FMA(sz, R02, F_FMA(R01, 0.0f, z00));                         // f149
    const float z11 = F_MUL(R11, 0.0f);                                              // f150
    const float M10 = F_FMA(R12, 0.0f, F_FMA(sx, R10, z11));                         // f152
    const float M11 = F_FMA(R12, 0.0f, F_FMA(R10, 0.0f, F_MUL(sy, R11)));            // f155
    const float M12 = F_FMA(sz, R12, F_FMA(R10, 0.0f, z11));                         // f157
    const float z21 = F_MUL(R21, 0.0f);                                              // f158
    const float M20 = F_FMA(R22, 0.0f, F_FMA(sx, R20, z21));                         // f160
    const float M21 = F_FMA(R22, 0.0f, F_FMA(R20, 0.0f, F_MUL(sy, R21)));            // f163
    const float M22 = F_FMA(sz, R22, F_FMA(R20, 0.0f, z21));                         // f165
    // Sigma = transpose(M) * M, upper triangle
    cov[0] = dot3_mid_first(M00, M00, M01, M01, M02, M02);   // f450
    cov[1] = dot3_mid_first(M10, M00, M11, M01, M12, M02);   // f449
    cov[2] = dot3_mid_first(M20, M00, M21, M01, M22, M02);   // f448
    cov[3] = dot3_mid_first(M10, M10, M11, M11, M12, M12);   // f447
    cov[4] = dot3_mid_first(M20, M10, M21, M11, M22, M12);   // f446
    cov[5] = dot3_mid_first(M20, M20, M21, M21, M22, M22);   // f445
}

// SH staging modes: 0 = read the coefficients straight from global memory; 1 = per-warp register-staged copy into an
// odd-stride shared tile (any row length); 2 = bulk asynchronous copies (TMA) of the 192-byte rows, one 384-byte copy
// per pair of Gaussians (SH degree 3, 16-byte aligned tensor), issued as soon as a Gaussian of the pair is known to be
// in front of the camera and awaited on a per-warp mbarrier right before the SH evaluation.
template <int STAGED, int M3C>
__global__ void __launch_bounds__(PRE_BLOCK) preprocess_kernel(PreArgs a) {
    extern __shared__ __align__(16) float s_tile_dyn[];
    __shared__ float s_cam[35];  // view 0..15, proj 16..31, campos 32..34
    __shared__ __align__(8) uint64_t s_bar[PRE_BLOCK / 32];
    if (threadIdx.x < 16) s_cam[threadIdx.x] = __ldg(a.viewmatrix + threadIdx.x);
    else if (threadIdx.x < 32) s_cam[threadIdx.x] = __ldg(a.projmatrix + threadIdx.x - 16);
    else if (threadIdx.x < 35) s_cam[threadIdx.x] = __ldg(a.cam_pos + threadIdx.x - 32);
    if (STAGED == 2 && (threadIdx.x & 31u) == 0) {
        lg_mbar_init(&s_bar[threadIdx.x >> 5], 32);  // every lane of the warp arrives once
        lg_mbar_init_fence();
    }
    __syncthreads();
    const uint32_t tile = blockIdx.x;
    const int idx = (int)(tile * PRE_BLOCK + threadIdx.x);
    const float* V = s_cam;
    const float* PM = s_cam + 16;
    if (STAGED == 2) {
        // start the SH row on its way before anything else: it arrives while the projection is computed
        bool fetch = false;
        if (idx < a.P) {
            const float qx = __ldg(a.means3D + 3 * (size_t)idx + 0), qy = __ldg(a.means3D + 3 * (size_t)idx + 1),
                        qz = __ldg(a.means3D + 3 * (size_t)idx + 2);
            fetch = !(xform_row(qx, qy, qz, V[2], V[6], V[10], V[14]) <= 0.2f);  // same predicate as in_frustum below
        }
        // rows travel in pairs (lanes 2p, 2p+1): the even lane fetches both if either Gaussian may need its colours
        const bool other_fetch = __shfl_xor_sync(0xffffffffu, fetch, 1);  // (not inside `||`: every lane must take part)
        const bool pair_fetch = fetch || other_fetch;
        uint64_t* bar = &s_bar[threadIdx.x >> 5];
        if (pair_fetch && (threadIdx.x & 1u) == 0) {
            const unsigned bytes = (idx + 1 < a.P ? 2u : 1u) * LG_SH_ROW_FLOATS * 4u;
            lg_mbar_arrive_expect_tx(bar, bytes);
            lg_bulk_load(lg_sh_row(s_tile_dyn, threadIdx.x), a.shs + (size_t)idx * LG_SH_ROW_FLOATS, bytes, bar);
        } else {
            lg_mbar_arrive(bar);
        }
    }

    uint32_t tiles = 0;
    uint32_t rx0 = 0, ry0 = 0, rx1 = 0, ry1 = 0;  // tile rectangle (all zero = emits nothing)
    int radius_out = 0;
    bool need_sh = false;
    float depth_out = 0.0f;
    uint32_t rect_out = 0;
    float px = 0.0f, py = 0.0f, pz = 0.0f;
    if (idx < a.P) {
        px = __ldg(a.means3D + 3 * (size_t)idx + 0);
        py = __ldg(a.means3D + 3 * (size_t)idx + 1);
        pz = __ldg(a.means3D + 3 * (size_t)idx + 2);
        // in_frustum (auxiliary.h:151-176)
        const float depth = xform_row(px, py, pz, V[2], V[6], V[10], V[14]);  // f1
        if (!(depth > 0.2f)) {
            if (a.prefiltered && depth <= 0.2f) {
                printf("Point is filtered although prefiltered is set. This shouldn't happen!");
                __trap();
            }
        }
        if (!(depth <= 0.2f)) {  // same predicate as the reference's `setp.le` early-out (NaN passes, as there)
            const float hx = xform_row(px, py, pz, PM[0], PM[4], PM[8], PM[12]);
            const float hy = xform_row(px, py, pz, PM[1], PM[5], PM[9], PM[13]);
            const float hw = xform_row(px, py, pz, PM[3], PM[7], PM[11], PM[15]);
            const float p_w = F_RCP(F_ADD(hw, 0.0000001f));
            const float proj_x = F_MUL(hx, p_w);  // f5
            const float proj_y = F_MUL(hy, p_w);  // f6

            float c3[6];
            if (a.cov3D_precomp != nullptr) {
#pragma unroll
                for (int k = 0; k < 6; k++) c3[k] = __ldg(a.cov3D_precomp + 6 * (size_t)idx + k);
            } else {
                const float s0 = __ldg(a.scales + 3 * (size_t)idx + 0);
                const float s1 = __ldg(a.scales + 3 * (size_t)idx + 1);
                const float s2 = __ldg(a.scales + 3 * (size_t)idx + 2);
                const float4 q = __ldg(reinterpret_cast<const float4*>(a.rotations) + idx);
                compute_cov3d(s0, s1, s2, a.scale_modifier, q.x, q.y, q.z, q.w, c3);
#pragma unroll
                for (int k = 0; k < 6; k++) a.cov3Ds[6 * (size_t)idx + k] = c3[k];
            }

            // computeCov2D (forward.cu:74-109)
            const float tx = xform_row(px, py, pz, V[0], V[4], V[8], V[12]);   // f186
            const float ty = xform_row(px, py, pz, V[1], V[5], V[9], V[13]);   // f194
            const float tz = depth;                                            // f202 (same expression as f1)
            const float limx = F_MUL(a.tan_fovx, 1.3f);
            const float limy = F_MUL(a.tan_fovy, 1.3f);
            const float txtz = F_DIV(tx, tz);
            const float tytz = F_DIV(ty, tz);
            const float cx = fminf(limx, fmaxf(-limx, txtz));                  // f209
            const float cy = fminf(limy, fmaxf(-limy, tytz));                  // f212
            const float J00 = F_DIV(a.focal_x, tz);                            // f213
            const float ntz = -tz;
            const float tz2 = F_MUL(tz, tz);                                   // f217
            const float J02 = F_DIV(F_MUL(a.focal_x, F_MUL(cx, ntz)), tz2);    // f218
            const float J11 = F_DIV(a.focal_y, tz);                            // f219
            const float J12 = F_DIV(F_MUL(a.focal_y, F_MUL(cy, ntz)), tz2);    // f222
            // T = W * J, W = view rotation as written in the reference (columns (v0,v4,v8),(v1,v5,v9),(v2,v6,v10))
            const float T00 = F_FMA(V[2], J02, F_FMA(V[0], J00, F_MUL(V[1], 0.0f)));    // f225
            const float T01 = F_FMA(V[6], J02, F_FMA(V[4], J00, F_MUL(V[5], 0.0f)));    // f228
            const float T02 = F_FMA(J02, V[10], F_FMA(V[8], J00, F_MUL(V[9], 0.0f)));   // f231
            const float T10 = F_FMA(V[2], J12, F_FMA(V[0], 0.0f, F_MUL(J11, V[1])));    // f234
            const float T11 = F_FMA(V[6], J12, F_FMA(V[4], 0.0f, F_MUL(J11, V[5])));    // f237
            const float T12 = F_FMA(J12, V[10], F_FMA(V[8], 0.0f, F_MUL(J11, V[9])));   // f240
            // A = transpose(T) * transpose(Vrk)
            const float A00 = dot3_mid_first(T00, c3[0], T01, c3[1], T02, c3[2]);  // f243
            const float A01 = dot3_mid_first(T10, c3[0], T11, c3[1], T12, c3[2]);  // f246
            const float A10 = dot3_mid_first(T00, c3[1], T01, c3[3], T02, c3[4]);  // f249
            const float A11 = dot3_mid_first(T10, c3[1], T11, c3[3], T12, c3[4]);  // f252
            const float A20 = dot3_mid_first(T00, c3[2], T01, c3[4], T02, c3[5]);  // f255
            const float A21 = dot3_mid_first(T10, c3[2], T11, c3[4], T12, c3[5]);  // f258
            // cov = A * T
            const float cov_xx = dot3_mid_first(T00, A00, T01, A10, T02, A20);     // f261
            const float cov_xy = dot3_mid_first(T00, A01, T01, A11, T02, A21);     // f25
            const float cov_yy = dot3_mid_first(T10, A01, T11, A11, T12, A21);     // f266

            // low-pass dilation + optional anti-aliasing compensation (forward.cu:215-231)
            const float xy2 = F_MUL(cov_xy, cov_xy);                               // f268
            const float det_cov = F_FMA(cov_xx, cov_yy, -xy2);                     // SASS: one FFMA
            const float cxx = F_ADD(cov_xx, 0.3f);                                 // f27
            const float cyy = F_ADD(cov_yy, 0.3f);                                 // f28
            const float det = F_FMA(cxx, cyy, -xy2);                               // SASS: one FFMA
            float h_scaling = 1.0f;
            if (a.antialiasing) h_scaling = F_SQRT(fmaxf(F_DIV(det_cov, det), 0.000025f));
            if (!(det == 0.0f)) {
                const float det_inv = F_RCP(det);
                const float con_x = F_MUL(cyy, det_inv);
                const float con_y = F_MUL(det_inv, -cov_xy);
                const float con_z = F_MUL(cxx, det_inv);
                // screen-space extent (forward.cu:237-240)
                const float mid = F_MUL(F_ADD(cxx, cyy), 0.5f);
                const float disc = F_SQRT(fmaxf(F_FMA(mid, mid, -det), 0.1f));  // SASS: one FFMA
                const float lam = fmaxf(F_ADD(mid, disc), F_SUB(mid, disc));
                const float radius_f = ceilf(F_MUL(F_SQRT(lam), 3.0f));
                const float pix_x = ndc2pix(proj_x, a.W);
                const float pix_y = ndc2pix(proj_y, a.H);
                const int radius = __float2int_rz(radius_f);
                uint32_t x0, y0, x1, y1;
                lg_get_rect(pix_x, pix_y, radius, a.grid_x, a.grid_y, x0, y0, x1, y1);
                const uint32_t n = (x1 - x0) * (y1 - y0);
                if (n != 0) {
                    need_sh = a.colors_precomp == nullptr;
                    a.depths[idx] = depth;
                    depth_out = depth;
                    rect_out = x0 | (y0 << 8) | (x1 << 16) | (y1 << 24);
                    a.means2D[idx] = make_float2(pix_x, pix_y);
                    const float opacity = __ldg(a.opacities + idx);
                    a.conic_opacity[idx] = make_float4(con_x, con_y, con_z, F_MUL(h_scaling, opacity));
                    tiles = n;
                    rx0 = x0; ry0 = y0; rx1 = x1; ry1 = y1;
                    radius_out = radius;
                }
            }
        }
        a.radii[idx] = radius_out;
        a.tiles_touched[idx] = tiles;
        a.rect_packed[idx] = rect_out;
    }

    // ---- binning step 1 (binning.cu): list length of every tile.  Fire-and-forget reductions (RED.ADD), one per
    // overlap; rectangles of more than PRE_COOP tiles are walked by the whole warp.
    {
        const unsigned lane = threadIdx.x & 31u;
        if (tiles > 0 && tiles <= PRE_COOP) {
            for (uint32_t y = ry0; y < ry1; y++)
                for (uint32_t x = rx0; x < rx1; x++) atomicAdd(a.tile_ctr + (size_t)(y * (uint32_t)a.grid_x + x) * LG_CTR_STRIDE, 1u);
        }
        unsigned big = __ballot_sync(0xffffffffu, tiles > PRE_COOP);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t bx0 = __shfl_sync(0xffffffffu, rx0, src), by0 = __shfl_sync(0xffffffffu, ry0, src);
            const uint32_t bw = __shfl_sync(0xffffffffu, rx1, src) - bx0;
            const uint32_t bn = __shfl_sync(0xffffffffu, tiles, src);
            for (uint32_t k = lane; k < bn; k += 32)
                atomicAdd(a.tile_ctr + (size_t)((by0 + k / bw) * (uint32_t)a.grid_x + bx0 + k % bw) * LG_CTR_STRIDE, 1u);
        }
    }

    // ---- SH -> RGB (forward.cu:20-71) for the Gaussians that survived culling; needs only 1e-5 image parity, the
    // association order of the reference is kept anyway.  One Gaussian's 3M coefficients are contiguous, so the warp
    // first moves the rows it needs into a shared-memory tile with coalesced 128-byte loads (row stride 3M | 1 is
    // odd: the per-thread walk along a row is bank-conflict free); STAGED = false reads global memory directly.
    {
        const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        const int M3 = a.M * 3;
        const float* sh = a.shs + (size_t)idx * M3;
        const unsigned need_mask = __ballot_sync(0xffffffffu, need_sh);
        float shv[LG_SH_ROW_FLOATS];
        if (STAGED == 2) {
            // every lane waits (a warp must not retire while the copy engine still writes into its rows)
            lg_mbar_wait(&s_bar[warp], 0);
            if (need_sh) {
                const float4* row4 = reinterpret_cast<const float4*>(lg_sh_row(s_tile_dyn, threadIdx.x));
#pragma unroll
                for (int j = 0; j < LG_SH_ROW_FLOATS / 4; j++) {
                    const float4 v = row4[j];
                    shv[4 * j + 0] = v.x; shv[4 * j + 1] = v.y; shv[4 * j + 2] = v.z; shv[4 * j + 3] = v.w;
                }
                sh = shv;
            }
        }
        if (STAGED == 1 && need_mask) {
            const int row = M3 | 1;
            float* s_wtile = s_tile_dyn + (size_t)warp * 32u * row;
            const size_t warp_first = (size_t)tile * PRE_BLOCK + (size_t)warp * 32u;
            const long long left = (long long)a.P - (long long)warp_first;
            const int warp_floats = (int)(left < 32 ? left : 32) * M3;
            lg_warp_rows_to_tile<M3C>(a.shs + warp_first * M3, s_wtile, M3, warp_floats, lane);
            __syncwarp();
            sh = s_wtile + lane * row;
        }
        if (need_sh) {
            const float dx = F_SUB(px, V[32]), dy = F_SUB(py, V[33]), dz = F_SUB(pz, V[34]);
            const float len = F_SQRT(F_FMA(dz, dz, F_FMA(dx, dx, F_MUL(dy, dy))));
            const float x = F_DIV(dx, len), y = F_DIV(dy, len), z = F_DIV(dz, len);
            float r[3];
#pragma unroll
            for (int c = 0; c < 3; c++) r[c] = F_MUL(sh[c], 0.28209479177387814f);
            if (a.D > 0) {
                const float k1y = F_MUL(y, 0.4886025119029199f);
                const float k1z = F_MUL(z, 0.4886025119029199f);
                const float k1x = F_MUL(x, 0.4886025119029199f);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float v = F_FMA(-k1y, sh[3 + c], r[c]);
                    v = F_FMA(k1z, sh[6 + c], v);
                    r[c] = F_FMA(-k1x, sh[9 + c], v);
                }
                if (a.D > 1) {
                    const float xx = F_MUL(x, x), yy = F_MUL(y, y), zz = F_MUL(z, z);
                    const float xy = F_MUL(x, y), yz = F_MUL(y, z), xz = F_MUL(x, z);
                    const float zz2 = F_ADD(zz, zz);
                    const float xx_yy = F_SUB(xx, yy);
                    const float k4 = F_MUL(xy, 1.0925484305920792f);
                    const float k5 = F_MUL(yz, -1.0925484305920792f);
                    const float k6 = F_MUL(F_SUB(F_SUB(zz2, xx), yy), 0.31539156525252005f);
                    const float k7 = F_MUL(xz, -1.0925484305920792f);
                    const float k8 = F_MUL(xx_yy, 0.5462742152960396f);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float v = F_FMA(k4, sh[12 + c], r[c]);
                        v = F_FMA(k5, sh[15 + c], v);
                        v = F_FMA(k6, sh[18 + c], v);
                        v = F_FMA(k7, sh[21 + c], v);
                        r[c] = F_FMA(k8, sh[24 + c], v);
                    }
                    if (a.D > 2) {
                        const float zz4_xx_yy = F_SUB(F_FMA(zz, 4.0f, -xx), yy);
                        const float k9 = F_MUL(F_MUL(y, -0.5900435899266435f), F_FMA(xx, 3.0f, -yy));
                        const float k10 = F_MUL(z, F_MUL(xy, 2.890611442640554f));
                        const float k11 = F_MUL(F_MUL(y, -0.4570457994644658f), zz4_xx_yy);
                        const float k12 = F_MUL(F_MUL(z, 0.3731763325901154f), F_FMA(yy, -3.0f, F_FMA(xx, -3.0f, zz2)));
                        const float k13 = F_MUL(F_MUL(x, -0.4570457994644658f), zz4_xx_yy);
                        const float k14 = F_MUL(F_MUL(z, 1.445305721320277f), xx_yy);
                        const float k15 = F_MUL(F_MUL(x, -0.5900435899266435f), F_FMA(yy, -3.0f, xx));
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            float v = F_FMA(k9, sh[27 + c], r[c]);
                            v = F_FMA(k10, sh[30 + c], v);
                            v = F_FMA(k11, sh[33 + c], v);
                            v = F_FMA(k12, sh[36 + c], v);
                            v = F_FMA(k13, sh[39 + c], v);
                            v = F_FMA(k14, sh[42 + c], v);
                            r[c] = F_FMA(k15, sh[45 + c], v);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float v = F_ADD(r[c], 0.5f);
                const bool neg = v < 0.0f;
                a.clamped[3 * (size_t)idx + c] = neg ? 1 : 0;
                a.rgb[3 * (size_t)idx + c] = neg ? 0.0f : v;
            }
        }
    }

}

// inspection only (lg_state_read "point_offsets"): the reference's inclusive scan of tiles_touched
// (rasterizer_impl.cu:280); one block walks the array with a running prefix
__global__ void __launch_bounds__(1024) point_offsets_kernel(int P, const uint32_t* __restrict__ tiles_touched,
                                                             uint32_t* __restrict__ offsets) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_running;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (int base = 0; base < P; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const uint32_t v = i < P ? tiles_touched[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = s_running;
        for (unsigned w = 0; w < warp; w++) before += s_warp[w];
        if (i < P) offsets[i] = before + incl;
        __syncthreads();
        if (threadIdx.x == 1023) s_running = before + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ viewmatrix,
                                                           uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float px = __ldg(means3D + 3 * (size_t)idx + 0);
    const float py = __ldg(means3D + 3 * (size_t)idx + 1);
    const float pz = __ldg(means3D + 3 * (size_t)idx + 2);
    const float depth = xform_row(px, py, pz, __ldg(viewmatrix + 2), __ldg(viewmatrix + 6), __ldg(viewmatrix + 10),
                                  __ldg(viewmatrix + 14));
    present[idx] = (depth <= 0.2f) ? 0 : 1;  // in_frustum (auxiliary.h:166)
}

int launch_preprocess(const ForwardArgs& f, GeometryState& g, ImageState& img, int* radii, cudaStream_t stream) {
    PreArgs a;
    a.P = f.P; a.D = f.D; a.M = f.M; a.C = f.C;
    a.means3D = f.means3D; a.scales = f.scales; a.scale_modifier = f.scale_modifier; a.rotations = f.rotations;
    a.opacities = f.opacities; a.shs = f.shs; a.cov3D_precomp = f.cov3D_precomp; a.colors_precomp = f.colors_precomp;
    a.viewmatrix = f.viewmatrix; a.projmatrix = f.projmatrix; a.cam_pos = f.cam_pos;
    a.W = f.W; a.H = f.H; a.tan_fovx = f.tan_fovx; a.tan_fovy = f.tan_fovy; a.focal_x = f.focal_x; a.focal_y = f.focal_y;
    a.grid_x = num_tiles_x(f.W); a.grid_y = num_tiles_y(f.H);
    a.prefiltered = f.prefiltered; a.antialiasing = f.antialiasing;
    a.radii = radii; a.means2D = g.means2D; a.depths = g.depths; a.cov3Ds = g.cov3D; a.rgb = g.rgb;
    a.conic_opacity = g.conic_opacity; a.clamped = g.clamped; a.tiles_touched = g.tiles_touched;
    a.rect_packed = g.rect_packed; a.tile_ctr = img.tile_ctr;
    const int blocks = (f.P + PRE_BLOCK - 1) / PRE_BLOCK;
    // counters and the per-tile counter lines are adjacent: one memset
    LG_CUDA(cudaMemsetAsync(img.counters, 0, sizeof(uint32_t) * LG_CTR_STRIDE * (1 + (size_t)a.grid_x * a.grid_y), stream));
    const int M3 = 3 * f.M;
    if (f.colors_precomp == nullptr && M3 <= PRE_MAX_ROW) {
        const size_t smem = sizeof(float) * PRE_BLOCK * (size_t)(M3 | 1);
        if (M3 == LG_SH_ROW_FLOATS && (reinterpret_cast<uintptr_t>(f.shs) & 15u) == 0 && PRE_SH_BULK) {
            // SH degree 3, 16-byte aligned rows: bulk asynchronous copies (one per Gaussian in front of the camera)
            const size_t bulk_smem = sizeof(float) * (PRE_BLOCK / 2) * (size_t)LG_SH_PAIR_FLOATS;
            LG_CUDA(cudaFuncSetAttribute(preprocess_kernel<2, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bulk_smem));
            preprocess_kernel<2, 48><<<blocks, PRE_BLOCK, bulk_smem, stream>>>(a);
        } else if (M3 == 48) {
            LG_CUDA(cudaFuncSetAttribute(preprocess_kernel<1, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
            preprocess_kernel<1, 48><<<blocks, PRE_BLOCK, smem, stream>>>(a);
        } else {
            LG_CUDA(cudaFuncSetAttribute(preprocess_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(float) * PRE_BLOCK * (PRE_MAX_ROW | 1))));
            preprocess_kernel<1, 0><<<blocks, PRE_BLOCK, smem, stream>>>(a);
        }
    } else {
        preprocess_kernel<0, 0><<<blocks, PRE_BLOCK, 0, stream>>>(a);
    }
    LG_LAUNCH_CHECK(f.debug, stream);
    return LG_OK;
}

int launch_point_offsets(int P, const GeometryState& g, uint32_t* offsets_out, cudaStream_t stream) {
    point_offsets_kernel<<<1, 1024, 0, stream>>>(P, g.tiles_touched, offsets_out);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

}  // namespace lg
