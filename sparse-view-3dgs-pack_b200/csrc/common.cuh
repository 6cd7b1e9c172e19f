// common.cuh — shared declarations for the B200-native LGDWT-GS hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include "../../include/lgdwt_b200.h"

#define LG_TILE_X 16
#define LG_TILE_Y 16
#define LG_TILE_PIX (LG_TILE_X * LG_TILE_Y)
#define LG_MAX_CHANNELS 4
#define LG_NUM_SMS 148
#define LG_CTR_STRIDE 32  // words between the binning counters of consecutive tiles (one 128-byte line each)

namespace lg {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
void count_launch();
// optional per-stage CUDA-event timing (lg_stage_timing_enable); no-ops when disabled
enum Stage { ST_PREPROCESS = 0, ST_BINNING, ST_BLEND_FWD, ST_BLEND_BWD, ST_PREGRAD, ST_COUNT };
void stage_begin(int stage, cudaStream_t stream);
void stage_end(int stage, cudaStream_t stream);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define LG_CUDA(call)                                                            \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) return lg::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// After a kernel launch: always catch launch-configuration errors; with debug also sync and check
// (the reference's CHECK_CUDA(debug) behaviour, DGR/cuda_rasterizer/auxiliary.h:178-185).
#define LG_LAUNCH_CHECK(debug, stream)                                           \
    do {                                                                         \
        lg::count_launch();                                                      \
        LG_CUDA(cudaGetLastError());                                             \
        if (debug) LG_CUDA(cudaStreamSynchronize(stream));                       \
    } while (0)

// ---------------------------------------------------------------- opaque state layouts
template <typename T>
static inline void carve(char*& chunk, T*& ptr, size_t count, size_t alignment = 128) {
    size_t offset = (reinterpret_cast<uintptr_t>(chunk) + alignment - 1) & ~(alignment - 1);
    ptr = reinterpret_cast<T*>(offset);
    chunk = reinterpret_cast<char*>(ptr + count);
}

// Per-Gaussian state (P-sized).  Same information as the reference GeometryState
// (DGR/cuda_rasterizer/rasterizer_impl.h:32-46) in a layout of our own.
struct GeometryState {
    float* depths;            // P
    uint8_t* clamped;         // 3P (SH clamp flags)
    int* internal_radii;      // P
    float2* means2D;          // P
    float* cov3D;             // 6P
    float4* conic_opacity;    // P
    float* rgb;               // channels * P
    uint32_t* tiles_touched;  // P
    uint32_t* rect_packed;    // P: tile rectangle x0 | y0 << 8 | x1 << 16 | y1 << 24 (grids up to 255 x 255 tiles)
    float* grad_scratch;      // backward only: 12 floats / Gaussian packed 2-D gradient record
    static GeometryState from_chunk(char*& chunk, size_t P, int channels);
};
size_t geometry_state_bytes(size_t P, int channels);

struct ImageState {
    float* accum_alpha;   // W*H final transmittance
    uint32_t* n_contrib;  // W*H
    uint2* ranges;        // T
    // launch order of the tiles in the sort / blend kernels, heaviest first (longest-processing-time-first scheduling:
    // the hardware hands out blocks in blockIdx order, so the long tiles of the image centre no longer form the tail)
    uint32_t* tile_order;      // T: tile sort and blend forward, by list length
    uint32_t* tile_neff;       // T: entries that reached some pixel of the tile (max n_contrib), written by the forward
    uint32_t* tile_order_bwd;  // T: backward, by tile_neff
    uint32_t* counters;        // [1] num_rendered (total list length, written by the tile scan), [3] blend-forward
                               // completion ticket (the last block derives the backward's launch order),
                               // [4] tile-sort error word, [5] longest tile list (tile scan)
    // Per-tile binning counters, one 128-byte line per tile (atomics on one line serialise in its L2 slice: with the
    // counters packed, 3.6 M increments on 80 lines took 160 us; one line per tile spreads them over all slices).
    // word 0 = list length (counted by the preprocess kernel), word 1 = write cursor of the scatter step.
    // Directly behind `counters`, cleared by the same memset.
    uint32_t* tile_ctr;        // T * LG_CTR_STRIDE
    static ImageState from_chunk(char*& chunk, size_t W, size_t H);
};
size_t image_state_bytes(size_t W, size_t H);

struct BinningState {
    uint2* pairs;          // R (depth bits, Gaussian id), grouped by tile, unsorted inside a tile
    uint2* pairs_alt;      // R ping-pong buffer of lists too long for shared memory
    uint32_t* point_list;  // R Gaussian ids sorted by (tile, depth, id)
    static BinningState from_chunk(char*& chunk, size_t R);
};
size_t binning_state_bytes(size_t R);

static inline int num_tiles_x(int W) { return (W + LG_TILE_X - 1) / LG_TILE_X; }
static inline int num_tiles_y(int H) { return (H + LG_TILE_Y - 1) / LG_TILE_Y; }

// ---------------------------------------------------------------- stage entry points (host side)
struct ForwardArgs {
    int P, D, M, C;
    const float* background;
    int W, H;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;
    bool prefiltered;
    bool antialiasing;
    bool debug;
};

int launch_preprocess(const ForwardArgs& a, GeometryState& g, ImageState& img, int* radii, cudaStream_t stream);
// binning (binning.cu): tile scan (ranges, cursors, num_rendered, launch order), then scatter + per-tile sort into a
// binning buffer of `capacity` entries (no-ops on the device when num_rendered exceeds it)
int launch_tile_scan(int W, int H, ImageState& img, bool debug, cudaStream_t stream);
int launch_binning(int P, int capacity, int W, int H, const GeometryState& g, const int* radii, BinningState& b,
                   ImageState& img, bool debug, cudaStream_t stream);
int launch_rebuild_keys(int W, int H, const GeometryState& g, const BinningState& b, const ImageState& img,
                        unsigned long long* keys_out, cudaStream_t stream);
int launch_blend_forward(int C, int W, int H, int capacity, const GeometryState& g, const BinningState& b, ImageState& img,
                         const float* features, const float* background, float* out_color, float* out_invdepth,
                         bool debug, cudaStream_t stream);
int launch_blend_count(int W, int H, const GeometryState& g, const BinningState& b, const ImageState& img,
                       unsigned long long* counts, cudaStream_t stream);
int launch_blend_backward(int P, int C, int W, int H, const GeometryState& g, const BinningState& b,
                          const ImageState& img, const float* features, const float* background,
                          const float* dL_dpix, const float* dL_dinvdepth_pix, float* grad_scratch, bool debug,
                          cudaStream_t stream);

struct BackwardArgs {
    int P, D, M, C;
    int W, H;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float tan_fovx, tan_fovy;
    float focal_x, focal_y;
    bool antialiasing;
    bool has_invdepth;
    const float* raw_rot_norm;  // non-null: parameter gradients w.r.t. the raw (pre-activation) parameters
    bool accumulate;  // add to dL_dmean3D / dL_dsh / dL_dopacity / dL_dscale / dL_drot instead of overwriting them
    float* dL_dmean2D;
    float* dL_dconic;
    float* dL_dopacity;
    float* dL_dcolor;
    float* dL_dinvdepth;
    float* dL_dmean3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscale;
    float* dL_drot;
};
int launch_preprocess_backward(const BackwardArgs& a, const GeometryState& g, const int* radii, bool debug,
                               cudaStream_t stream);

// radix sort (radix_sort.cu).  Input in (keys_a, vals_a); both buffer pairs are clobbered; the sorted result lands
// in (keys_b, vals_b) when the number of 8-bit passes is odd, else in (keys_a, vals_a) (*result_in_b says which).
size_t radix_sort_temp_bytes(size_t n, int key_bytes);
int radix_sort_num_passes(int begin_bit, int end_bit);
int radix_sort_pairs_u64(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n,
                         int begin_bit, int end_bit, char* temp, size_t temp_bytes, bool debug, cudaStream_t stream,
                         bool* result_in_b);
int radix_sort_pairs_u32(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n,
                         int begin_bit, int end_bit, char* temp, size_t temp_bytes, bool debug, cudaStream_t stream,
                         bool* result_in_b);
uint32_t* radix_sort_hist_ptr(char* temp);  // [pass * 256 + digit] counters a producer kernel may fill itself ...
int radix_sort_clear(char* temp, size_t n, int passes, cudaStream_t stream);  // ... after this, and then call:
int radix_sort_pairs_u32_prehist(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n,
                                 int begin_bit, int end_bit, char* temp, size_t temp_bytes, bool debug,
                                 cudaStream_t stream, bool* result_in_b);
int launch_point_offsets(int P, const GeometryState& g, uint32_t* offsets_out, cudaStream_t stream);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t stream);

}  // namespace lg

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// Explicitly rounded fp32 ops: never re-contracted by nvcc, so the bit-exact contract (read off the
// reference's sm_100a PTX) survives any restructuring of the surrounding code.
#define F_MUL(a, b) __fmul_rn((a), (b))
#define F_ADD(a, b) __fadd_rn((a), (b))
#define F_SUB(a, b) __fsub_rn((a), (b))
#define F_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define F_DIV(a, b) __fdiv_rn((a), (b))
#define F_RCP(a) __frcp_rn((a))
#define F_SQRT(a) __fsqrt_rn((a))

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// getRect (DGR/cuda_rasterizer/auxiliary.h:45-55) with BLOCK_X = BLOCK_Y = 16: the divisions by 16 are exact
// multiplies by 0.0625 and `p + r + 15` is evaluated as ((p + r) + 16) - 1 (reference PTX).
__device__ __forceinline__ void lg_get_rect(float px, float py, int radius, int grid_x, int grid_y, uint32_t& x0,
                                            uint32_t& y0, uint32_t& x1, uint32_t& y1) {
    const float r = (float)radius;
    x0 = min((uint32_t)grid_x, (uint32_t)max(0, __float2int_rz(F_MUL(F_SUB(px, r), 0.0625f))));
    y0 = min((uint32_t)grid_y, (uint32_t)max(0, __float2int_rz(F_MUL(F_SUB(py, r), 0.0625f))));
    x1 = min((uint32_t)grid_x,
             (uint32_t)max(0, __float2int_rz(F_MUL(F_ADD(F_ADD(F_ADD(px, r), 16.0f), -1.0f), 0.0625f))));
    y1 = min((uint32_t)grid_y,
             (uint32_t)max(0, __float2int_rz(F_MUL(F_ADD(F_ADD(F_ADD(py, r), 16.0f), -1.0f), 0.0625f))));
}

// Which 4x4 pixel sub-patches of a 16x16 tile can a list entry contribute to at all?
// alpha = o * exp(power) >= 1/255 needs q := a dx^2 + 2 b dx dy + c dy^2 = -2 power <= 2 ln(255 o).  For every
// sub-patch the exact minimum of the convex quadratic q over its rectangle (pixel centres x0 + 4 col .. + 3,
// y0 + 4 r .. + 3, taken as a continuous box, so a lower bound of q on the pixels) is compared with that threshold: if
// the mean lies in the box the minimum is 0, otherwise it sits on the box edge(s) facing the mean, where q is a 1-D
// parabola whose minimiser is clamped to the edge.  Bit (r*4 + col) of the result is clear only if every pixel of that
// sub-patch fails the reference's alpha test (forward.cu:358-361), with a margin (1 % in alpha, 0.1 % in q) that is
// four orders of magnitude above the fp32 rounding of `power`; skipping such an entry for those pixels therefore never
// changes a decision the reference takes.  Degenerate inputs (non-positive or NaN a, c; NaN opacity) keep all bits set.
// The blend forward ORs bits 2j and 2j + 1 into the bit of 8x4 patch j (one list per warp); the backward walks one list
// per half-warp: on the benchmark scene the trips per warp — the longer of its two lists — are 21 % fewer than the
// entries that reach the warp's 8x4 patch.
__device__ __forceinline__ unsigned lg_subpatch_mask16(float mx, float my, float4 conic_opacity, float tx0, float ty0) {
    const float a = conic_opacity.x, b = conic_opacity.y, c = conic_opacity.z, o = conic_opacity.w;
    if (o <= 0.0f) return 0u;                                       // alpha <= 0 on every pixel
    if (!(a > 0.0f) || !(c > 0.0f) || !(o > 0.0f)) return 0xffffu;
    const float qmax = 2.002f * (__logf(255.0f * o) + 0.01f);       // negative when o < 1/255: nothing survives
    const float nb_over_c = -b * __fdividef(1.0f, c), nb_over_a = -b * __fdividef(1.0f, a);
    float xlo[4], xhi[4], ax2[4], bx2[4], ty[4];
    bool in_x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        xlo[k] = tx0 + 4.0f * k - mx;
        xhi[k] = xlo[k] + 3.0f;
        in_x[k] = xlo[k] <= 0.0f && xhi[k] >= 0.0f;
        const float ex = xlo[k] > 0.0f ? xlo[k] : xhi[k];  // offset of the box edge facing the mean
        ax2[k] = a * ex * ex;
        bx2[k] = 2.0f * b * ex;
        ty[k] = nb_over_c * ex;                            // unconstrained minimiser of q along that edge
    }
    unsigned m = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const float ylo = ty0 + 4.0f * r - my, yhi = ylo + 3.0f;
        const bool in_y = ylo <= 0.0f && yhi >= 0.0f;
        const float ey = ylo > 0.0f ? ylo : yhi;
        const float cy2 = c * ey * ey, by2 = 2.0f * b * ey, tx = nb_over_a * ey;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float dy = fminf(fmaxf(ty[k], ylo), yhi);
            const float q1 = in_x[k] ? 3.0e38f : fmaf(dy, fmaf(c, dy, bx2[k]), ax2[k]);
            const float dx = fminf(fmaxf(tx, xlo[k]), xhi[k]);
            const float q2 = in_y ? 3.0e38f : fmaf(dx, fmaf(a, dx, by2), cy2);
            const float qmin = (in_x[k] && in_y) ? 0.0f : fminf(q1, q2);
            m |= (qmin > qmax) ? 0u : (1u << (r * 4 + k));
        }
    }
    return m;
}

typedef uint16_t lg_slot_t;  // byte offset of a staged list entry inside its batch

// A warp's 32 Gaussians own a contiguous block of 32 rows x M3 floats of a (P, M3) array.  Copy `count` floats of that
// block between global memory and a shared-memory tile whose row stride is M3 | 1 (odd: bank-conflict-free row
// walks), with fully coalesced 128-byte global accesses and LG_ROW_BATCH independent loads in flight per lane.
// Element e of the block lives at tile[e + e / M3] when M3 is even (row = M3 + 1) and at tile[e] when it is odd.
// M3C is the compile-time row length (48 = SH degree 3, the standard case: the division becomes a multiply) or 0
// for a run-time M3.
#define LG_ROW_BATCH 12
template <int M3C>
__device__ __forceinline__ int lg_tile_index(int e, int M3) {
    if constexpr (M3C > 0) return (M3C & 1) ? e : e + e / M3C;
    else return (M3 & 1) ? e : e + e / M3;
}
// Full-warp fast path for the standard row length 48 (SH degree 3): element e = lane + 32 u of the block lives at
// e + e / 48 = lane + 32 u + floor(2u / 3) + (u % 3 == 1 && lane >= 16), i.e. at a compile-time offset from one of two
// per-lane bases, so every copy is one LDG/STG plus one STS/LDS with an immediate offset and no address arithmetic.
__device__ __forceinline__ int lg_tile_index48(int u, unsigned lane) {
    return (int)lane + 32 * u + (2 * u) / 3 + ((u % 3 == 1) ? (int)(lane >> 4) : 0);
}
template <int M3C>
__device__ __forceinline__ void lg_warp_rows_to_tile(const float* __restrict__ src, float* tile, int M3, int count,
                                                     unsigned lane) {
    if (M3C == 48 && count == 32 * 48) {
#pragma unroll
        for (int u0 = 0; u0 < 48; u0 += LG_ROW_BATCH) {
            float v[LG_ROW_BATCH];
#pragma unroll
            for (int u = 0; u < LG_ROW_BATCH; u++) v[u] = __ldg(src + lane + 32 * (u0 + u));
#pragma unroll
            for (int u = 0; u < LG_ROW_BATCH; u++) tile[lg_tile_index48(u0 + u, lane)] = v[u];
        }
        return;
    }
    for (int e0 = (int)lane; e0 < count; e0 += 32 * LG_ROW_BATCH) {
        float v[LG_ROW_BATCH];
#pragma unroll
        for (int u = 0; u < LG_ROW_BATCH; u++) {
            const int e = e0 + 32 * u;
            v[u] = e < count ? __ldg(src + e) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < LG_ROW_BATCH; u++) {
            const int e = e0 + 32 * u;
            if (e < count) tile[lg_tile_index<M3C>(e, M3)] = v[u];
        }
    }
}
template <int M3C>
__device__ __forceinline__ void lg_warp_tile_to_rows(float* __restrict__ dst, const float* tile, int M3, int count,
                                                     unsigned lane, bool accumulate) {
    if (M3C == 48 && count == 32 * 48) {
        if (accumulate) {
#pragma unroll
            for (int u0 = 0; u0 < 48; u0 += LG_ROW_BATCH) {
                float v[LG_ROW_BATCH];
#pragma unroll
                for (int u = 0; u < LG_ROW_BATCH; u++) v[u] = dst[lane + 32 * (u0 + u)];
#pragma unroll
                for (int u = 0; u < LG_ROW_BATCH; u++)
                    dst[lane + 32 * (u0 + u)] = v[u] + tile[lg_tile_index48(u0 + u, lane)];
            }
        } else {
#pragma unroll
            for (int u = 0; u < 48; u++) dst[lane + 32 * u] = tile[lg_tile_index48(u, lane)];
        }
        return;
    }
    if (accumulate) {
        for (int e0 = (int)lane; e0 < count; e0 += 32 * LG_ROW_BATCH) {
            float v[LG_ROW_BATCH];
#pragma unroll
            for (int u = 0; u < LG_ROW_BATCH; u++) {
                const int e = e0 + 32 * u;
                v[u] = e < count ? dst[e] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < LG_ROW_BATCH; u++) {
                const int e = e0 + 32 * u;
                if (e < count) dst[e] = v[u] + tile[lg_tile_index<M3C>(e, M3)];
            }
        }
    } else {
        for (int e = (int)lane; e < count; e += 32) dst[e] = tile[lg_tile_index<M3C>(e, M3)];
    }
}

// Decoupled look-back of a single-pass prefix sum, executed by one full warp of the block holding ticket `tile`.
// Descriptor = (flag << 32) | value; flag 0 = not published, 1 = the block's own total, 2 = inclusive prefix.
// Publishes this block's total, adds up the predecessors' descriptors back to the nearest inclusive prefix, publishes
// the block's inclusive prefix and returns its exclusive prefix to every lane.  Each lane fetches LG_LB_DEPTH
// descriptors per round trip (32 * LG_LB_DEPTH predecessors per memory latency): when a whole wave of blocks starts
// together nobody holds an inclusive prefix yet and block t must add ~t totals, so the width of the window, not the
// arithmetic, sets the length of the chain.  Blocks take their tickets in launch order, so every predecessor is
// running or finished and will publish (no deadlock).
#define LG_LB_DEPTH 4
__device__ __forceinline__ uint32_t lg_lookback_exclusive(volatile unsigned long long* st, uint32_t tile,
                                                          uint32_t block_total, unsigned lane) {
    if (tile == 0) {
        if (lane == 0) st[0] = (2ull << 32) | block_total;
        return 0u;
    }
    if (lane == 0) st[tile] = (1ull << 32) | block_total;
    uint32_t exclusive = 0;
    int base = (int)tile - 1;
    while (true) {
        unsigned long long d[LG_LB_DEPTH];
#pragma unroll
        for (int k = 0; k < LG_LB_DEPTH; k++) {
            const int j = base - (int)lane - 32 * k;
            d[k] = 2ull << 32;  // virtual predecessor of block 0: inclusive prefix 0
            if (j >= 0) d[k] = st[j];
        }
        bool finished = false;
#pragma unroll
        for (int k = 0; k < LG_LB_DEPTH; k++) {
            if (!finished) {
                const int j = base - (int)lane - 32 * k;
                while ((d[k] >> 32) == 0ull) d[k] = st[j];
                const uint32_t flag = (uint32_t)(d[k] >> 32), val = (uint32_t)d[k];
                const unsigned done_mask = __ballot_sync(0xffffffffu, flag == 2u);
                uint32_t contrib = val;
                if (done_mask) {
                    const int first = __ffs(done_mask) - 1;  // nearest predecessor holding an inclusive prefix
                    contrib = lane <= (unsigned)first ? val : 0u;
                    finished = true;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                exclusive += contrib;
            }
        }
        if (finished) break;
        base -= 32 * LG_LB_DEPTH;
    }
    if (lane == 0) st[tile] = (2ull << 32) | (unsigned long long)(exclusive + block_total);
    return exclusive;
}

// Tile launch order for the sort / blend kernels: bucket the T tiles by key (1024 buckets scaled to the largest key)
// and list them heaviest bucket first (longest-processing-time-first: the hardware hands out blocks in blockIdx
// order).  Executed by one whole block of THREADS threads; s_cnt = 1024 words, s_warp = 33 words of shared memory.
template <int THREADS>
__device__ __forceinline__ void lg_bucket_order(int T, const uint32_t* keys, uint32_t* __restrict__ order,
                                                uint32_t* s_cnt, uint32_t* s_warp) {
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    for (int k = tid; k < 1024; k += THREADS) s_cnt[k] = 0;
    if (tid == 0) s_warp[32] = 1;
    __syncthreads();
    uint32_t m = 0;
    for (int t = tid; t < T; t += THREADS) m = max(m, keys[t]);
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) atomicMax(&s_warp[32], m);
    __syncthreads();
    const float scale = 1023.0f / (float)s_warp[32];
    for (int t = tid; t < T; t += THREADS) atomicAdd(&s_cnt[1023 - min((int)((float)keys[t] * scale), 1023)], 1u);
    __syncthreads();
    // exclusive scan of the 1024 bucket counts: thread t owns buckets [t * PER, (t + 1) * PER)
    constexpr int PER = 1024 / THREADS;
    uint32_t c[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        c[k] = s_cnt[tid * PER + k];
        sum += c[k];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t run = incl - sum;
    for (int w = 0; w < (int)warp; w++) run += s_warp[w];
#pragma unroll
    for (int k = 0; k < PER; k++) {
        s_cnt[tid * PER + k] = run;
        run += c[k];
    }
    __syncthreads();
    for (int t = tid; t < T; t += THREADS) {
        const uint32_t pos = atomicAdd(&s_cnt[1023 - min((int)((float)keys[t] * scale), 1023)], 1u);
        order[pos] = (uint32_t)t;
    }
}

// ---- bulk asynchronous copies (TMA, 1-D) and the mbarrier that tracks them.  Used where a Gaussian's data is one
// contiguous, 16-byte-sized row (the 192-byte SH row): the copy engine moves the row into shared memory while the thread
// does its projection math, instead of 48 LDG + 48 STS per thread through registers.
__device__ __forceinline__ uint32_t lg_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lg_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lg_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lg_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void lg_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lg_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void lg_mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lg_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lg_mbar_wait(uint64_t* bar, unsigned phase) {
    const uint32_t addr = lg_smem_addr(bar);
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(phase) : "memory");
    } while (!done);
}
// global -> shared, completion counted in bytes on `bar` (dst, src and bytes multiples of 16)
__device__ __forceinline__ void lg_bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(lg_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(lg_smem_addr(bar)) : "memory");
}
// shared -> global (plain store, or element-wise fp32 add into global memory), tracked by the thread's bulk group
__device__ __forceinline__ void lg_bulk_store(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(lg_smem_addr(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lg_bulk_reduce_add_f32(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(lg_smem_addr(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lg_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void lg_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (before a shared -> global bulk copy)
__device__ __forceinline__ void lg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// SH degree 3: a Gaussian's 16 coefficients x 3 channels = one 192-byte row.  The bulk-copy instruction is issued by
// one lane at a time (UBLKCP takes warp-uniform operands: per-lane copies become a 32-trip loop), so the rows travel
// in PAIRS — the rows of lanes 2p and 2p+1 are adjacent in global memory: one 384-byte copy per even lane — into
// shared memory at a pair stride of 400 bytes: with 25 sixteen-byte chunks per pair, chunk j of eight consecutive
// lanes falls into eight different bank groups (25 p + 12 q + j mod 8, p = 0..3, q = 0..1), so the per-lane float4 reads
// and writes of the rows are conflict-free.
#define LG_SH_ROW_FLOATS 48
#define LG_SH_PAIR_FLOATS 100
__device__ __forceinline__ float* lg_sh_row(float* base, unsigned tid) {
    return base + (size_t)(tid >> 1) * LG_SH_PAIR_FLOATS + (tid & 1u) * LG_SH_ROW_FLOATS;
}

// 128-bit read-only streaming load
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }
#endif
