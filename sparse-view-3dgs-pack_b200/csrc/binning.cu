// binning.cu — tile binning.  Produces what duplicateWithKeys + cub::DeviceRadixSort::SortPairs + cudaMemset +
// identifyTileRanges produce in the reference (DGR/cuda_rasterizer/rasterizer_impl.cu:70-138, 292-321): the list of
// Gaussian ids ordered by (tile, depth bits, Gaussian id) and the [begin, end) range of every tile.
//
// The reference sorts R = sum(tiles_touched) 64-bit (tile | depth) keys.  The same total order is reached with far
// less traffic by a two-level LSD scheme:
//   1. order the P Gaussians by their 32 depth bits (stable, so equal depths stay in ascending-id order);
//   2. emit one (tile id, Gaussian id) pair per overlap while walking the Gaussians in that order, so that inside
//      every tile the pairs already appear in (depth, id) order;
//   3. stable-sort the R pairs by the tile id alone (ceil(log2 T) = 12..13 bits: two 8-bit passes).
// Stable LSD sorting by the minor key first and the major key second is exactly a sort by (tile, depth, id) — the
// order the reference's stable 44..45-bit sort yields (a Gaussian appears at most once per tile, so the order is
// total and every correct sort reproduces the reference list bit for bit).  P-sized passes replace four of the six
// R-sized passes and the R-sized passes move 8 bytes per pair instead of 12.
#include "common.cuh"

namespace lg {

// bits needed for the tile id: the reference's getHigherMsb (rasterizer_impl.cu:35-50) == 32 - clz(n) for n >= 1
// (tests/test_host_logic.py checks the equality against a line-by-line restatement).
int higher_msb(uint32_t n) {
    int b = 0;
    while (n) { b++; n >>= 1; }
    return b < 1 ? 1 : b;
}

#define EMIT_BLOCK 256
#define EMIT_IPT 4   // Gaussians per thread
#define EMIT_TILE (EMIT_BLOCK * EMIT_IPT)
#define EMIT_CAP 5120 // pairs staged in shared memory per round (40 KB)
#define EMIT_COOP 8  // rectangles of more tiles than this are written by the whole warp (coalesced stores)

// Walks the Gaussians in depth order.  A thread takes EMIT_IPT consecutive entries of `order`, learns where their
// pairs start from a fused exclusive scan of the tile counts in that order (block scan + decoupled look-back over
// blocks taken in ticket order), and writes one (tile, id) pair per tile of each rectangle, row-major as the
// reference's double loop (rasterizer_impl.cu:93-108; the order inside one Gaussian is irrelevant to the result,
// its tiles being distinct).  The block also accumulates the digit histograms of the tile ids it writes, which the
// tile-id sort that follows would otherwise need a pass of its own for.
__global__ void __launch_bounds__(EMIT_BLOCK) emit_pairs_kernel(int P, const uint32_t* __restrict__ order,
                                                                const uint32_t* __restrict__ tiles_touched,
                                                                const uint32_t* __restrict__ rect_packed,
                                                                const float2* __restrict__ xy,
                                                                const int* __restrict__ radii,
                                                                uint32_t* __restrict__ tile_keys,
                                                                uint32_t* __restrict__ ids, int grid_x, int grid_y,
                                                                unsigned long long* scan_state, uint32_t* ticket,
                                                                uint32_t* tile_hist, uint32_t mask0, uint32_t mask1) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_sums[EMIT_BLOCK / 32];
    __shared__ uint32_t s_block_prefix, s_block_total;
    __shared__ uint32_t s_keys[EMIT_CAP], s_ids[EMIT_CAP];  // the block's pairs at their block-local offsets
    __shared__ uint32_t s_hist[2 * 256];  // digit histograms of the tile ids emitted by this block (two 8-bit places)
    s_hist[threadIdx.x] = 0;
    s_hist[256 + threadIdx.x] = 0;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int i0 = (int)(tile * EMIT_TILE + threadIdx.x * EMIT_IPT);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

    // per item: Gaussian id and its tile rectangle packed as xs = x0 | x1 << 16, ys = y0 | y1 << 16
    uint32_t id[EMIT_IPT], xs[EMIT_IPT], ys[EMIT_IPT];
#pragma unroll
    for (int u = 0; u < EMIT_IPT; u++) id[u] = (i0 + u < P) ? order[i0 + u] : 0u;
    uint32_t mine = 0;
#pragma unroll
    for (int u = 0; u < EMIT_IPT; u++) {
        xs[u] = ys[u] = 0;
        if (i0 + u < P) {
            if (rect_packed) {  // one 4-byte gather gives both the count and the rectangle
                const uint32_t r = rect_packed[id[u]];
                xs[u] = (r & 255u) | (((r >> 16) & 255u) << 16);
                ys[u] = ((r >> 8) & 255u) | ((r >> 24) << 16);
            } else if (tiles_touched[id[u]] > 0) {
                const float2 p = xy[id[u]];
                uint32_t x0, y0, x1, y1;
                lg_get_rect(p.x, p.y, radii[id[u]], grid_x, grid_y, x0, y0, x1, y1);
                xs[u] = x0 | (x1 << 16);
                ys[u] = y0 | (y1 << 16);
            }
        }
        mine += ((xs[u] >> 16) - (xs[u] & 0xffffu)) * ((ys[u] >> 16) - (ys[u] & 0xffffu));
    }
    // ---- exclusive scan of the tile counts over the depth-ordered Gaussians
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) s_warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t ws = lane < EMIT_BLOCK / 32 ? s_warp_sums[lane] : 0u;
        uint32_t wincl = ws;
#pragma unroll
        for (int o = 1; o < EMIT_BLOCK / 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, o);
            if (lane >= (unsigned)o) wincl += t;
        }
        if (lane < EMIT_BLOCK / 32) s_warp_sums[lane] = wincl - ws;  // exclusive warp offsets
        const uint32_t block_total = __shfl_sync(0xffffffffu, wincl, EMIT_BLOCK / 32 - 1);
        const uint32_t exclusive = lg_lookback_exclusive(scan_state, tile, block_total, lane);
        if (lane == 0) {
            s_block_prefix = exclusive;
            s_block_total = block_total;
        }
    }
    __syncthreads();
    // ---- emission.  Every thread owns a run of consecutive output slots, but runs are short (3.6 pairs per Gaussian
    // on the benchmark scene), so writing them straight to global memory touches ~11 sectors per store instruction.
    // The block's pairs are instead laid out in shared memory at their block-local offsets and streamed out with
    // fully coalesced stores, EMIT_CAP pairs per round (one round for all but the densest blocks).
    const uint32_t block_base = s_block_prefix;
    const uint32_t block_total = s_block_total;
    const uint32_t my_first = s_warp_sums[warp] + incl - mine;  // block-local offset of this thread's first pair
    for (uint32_t chunk = 0; chunk < block_total; chunk += EMIT_CAP) {
        uint32_t off = my_first;
#pragma unroll
        for (int u = 0; u < EMIT_IPT; u++) {
            const uint32_t x0 = xs[u] & 0xffffu, x1 = xs[u] >> 16, y0 = ys[u] & 0xffffu, y1 = ys[u] >> 16;
            const uint32_t nu = (x1 - x0) * (y1 - y0);
            const bool touches = nu > 0 && off < chunk + EMIT_CAP && off + nu > chunk;
            if (touches && nu <= EMIT_COOP) {
                uint32_t o = off;
                for (uint32_t y = y0; y < y1; y++)
                    for (uint32_t x = x0; x < x1; x++, o++) {
                        const uint32_t q = o - chunk;  // wraps for o < chunk
                        if (q < EMIT_CAP) {
                            s_keys[q] = y * (uint32_t)grid_x + x;
                            s_ids[q] = id[u];
                        }
                    }
            }
            unsigned big = __ballot_sync(0xffffffffu, touches && nu > EMIT_COOP);
            while (big) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
                const uint32_t bw = __shfl_sync(0xffffffffu, x1, src) - bx0;
                const uint32_t bn = __shfl_sync(0xffffffffu, nu, src), boff = __shfl_sync(0xffffffffu, off, src);
                const uint32_t bid = __shfl_sync(0xffffffffu, id[u], src);
                for (uint32_t k = lane; k < bn; k += 32) {
                    const uint32_t q = boff + k - chunk;
                    if (q < EMIT_CAP) {
                        s_keys[q] = (by0 + k / bw) * (uint32_t)grid_x + bx0 + k % bw;
                        s_ids[q] = bid;
                    }
                }
            }
            off += nu;
        }
        __syncthreads();
        const uint32_t count = min((uint32_t)EMIT_CAP, block_total - chunk);
        for (uint32_t q = threadIdx.x; q < count; q += EMIT_BLOCK) {
            const uint32_t t = s_keys[q];
            tile_keys[block_base + chunk + q] = t;
            ids[block_base + chunk + q] = s_ids[q];
            atomicAdd(&s_hist[t & mask0], 1u);
            atomicAdd(&s_hist[256 + ((t >> 8) & mask1)], 1u);
        }
        __syncthreads();
    }
    {
        const uint32_t c0 = s_hist[threadIdx.x], c1 = s_hist[256 + threadIdx.x];
        if (c0) atomicAdd(tile_hist + threadIdx.x, c0);
        if (c1) atomicAdd(tile_hist + 256 + threadIdx.x, c1);
    }
}

// identifyTileRanges (rasterizer_impl.cu:116-138) on the sorted tile ids; tiles with no entries keep (0,0).
// Four consecutive ids per thread (one 16-byte load + the id before them).
__global__ void __launch_bounds__(256) tile_ranges_kernel(uint32_t L, const uint32_t* __restrict__ tile_keys,
                                                          uint2* __restrict__ ranges) {
    const uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (base >= L) return;
    uint32_t k[4];
    if (base + 4u <= L) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile_keys + base);
        k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) k[i] = base + i < L ? tile_keys[base + i] : 0u;
    }
    uint32_t prev = base == 0 ? 0u : tile_keys[base - 1];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t idx = base + i;
        if (idx < L) {
            const uint32_t cur = k[i];
            if (idx == 0) ranges[cur].x = 0;
            else if (cur != prev) {
                ranges[prev].y = idx;
                ranges[cur].x = idx;
            }
            if (idx == L - 1) ranges[cur].y = L;
            prev = cur;
        }
    }
}

// Tile launch order for the blend kernels: one block buckets the T tiles by key (1024 buckets scaled to the largest
// key) and lists them heaviest bucket first.  keys[t * key_stride + key_offset] with (stride 2, offset: y - x computed
// here) for ranges, (stride 1) for tile_neff.
__global__ void __launch_bounds__(1024) tile_order_kernel(int T, const uint32_t* __restrict__ keys, int key_stride,
                                                          uint32_t* __restrict__ order) {
    __shared__ uint32_t s_cnt[1024];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    s_cnt[tid] = 0;
    if (tid == 0) s_max = 1;
    __syncthreads();
    auto key_of = [&](int t) -> uint32_t {
        return key_stride == 2 ? keys[2 * t + 1] - keys[2 * t] : keys[t];
    };
    uint32_t m = 0;
    for (int t = tid; t < T; t += 1024) m = max(m, key_of(t));
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) atomicMax(&s_max, m);
    __syncthreads();
    const float scale = 1023.0f / (float)s_max;
    // bucket 0 = heaviest
    for (int t = tid; t < T; t += 1024) atomicAdd(&s_cnt[1023 - min((int)((float)key_of(t) * scale), 1023)], 1u);
    __syncthreads();
    // exclusive scan of the 1024 bucket counts
    const uint32_t c = s_cnt[tid];
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((int)lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], z = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, z, o);
            if ((int)lane >= o) z += y;
        }
        s_warp[lane] = z - w;
    }
    __syncthreads();
    s_cnt[tid] = x - c + s_warp[warp];
    __syncthreads();
    for (int t = tid; t < T; t += 1024) {
        const uint32_t pos = atomicAdd(&s_cnt[1023 - min((int)((float)key_of(t) * scale), 1023)], 1u);
        order[pos] = (uint32_t)t;
    }
}

int launch_tile_order(int T, const uint32_t* keys, int key_stride, uint32_t* order, cudaStream_t stream) {
    tile_order_kernel<<<1, 1024, 0, stream>>>(T, keys, key_stride, order);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

// inspection only (lg_state_read "point_list_keys"): the reference's sorted 64-bit keys
__global__ void __launch_bounds__(256) rebuild_keys_kernel(uint32_t L, const uint32_t* __restrict__ tile_keys,
                                                           const uint32_t* __restrict__ point_list,
                                                           const float* __restrict__ depths,
                                                           unsigned long long* __restrict__ keys) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    keys[idx] = ((unsigned long long)tile_keys[idx] << 32) | __float_as_uint(depths[point_list[idx]]);
}

// step 1: Gaussian ids in (depth bits, id) order -> g.depth_ids[0] (four 8-bit passes: the result is back in buffer 0)
int launch_depth_order(int P, GeometryState& g, bool debug, cudaStream_t stream) {
    bool in_b = false;
    // the digit histograms were accumulated by the preprocess kernel
    int rc = radix_sort_pairs_u32_prehist(g.depth_keys[0], g.depth_keys[1], g.depth_ids[0], g.depth_ids[1], (size_t)P, 0,
                                          32, g.sort_temp, g.sort_temp_bytes, debug, stream, &in_b);
    if (rc != LG_OK) return rc;
    if (in_b) {
        set_error("depth ordering: unexpected pass parity");
        return LG_ERR_UNSUPPORTED;
    }
    return LG_OK;
}

// steps 2 and 3 + tile ranges
int launch_binning(int P, int R, int W, int H, const GeometryState& g, const int* radii, BinningState& b,
                   ImageState& img, bool debug, cudaStream_t stream) {
    const int gx = num_tiles_x(W), gy = num_tiles_y(H);
    const int T = gx * gy;
    LG_CUDA(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)T, stream));
    if (R <= 0) return launch_tile_order(T, reinterpret_cast<const uint32_t*>(img.ranges), 2, img.tile_order, stream);
    const int end_bit = higher_msb((uint32_t)T);
    const int passes = radix_sort_num_passes(0, end_bit);
    // ping-pong so that the sorted list ends in tile_keys / point_list
    uint32_t* ka = (passes & 1) ? b.tile_keys_unsorted : b.tile_keys;
    uint32_t* kb = (passes & 1) ? b.tile_keys : b.tile_keys_unsorted;
    uint32_t* va = (passes & 1) ? b.point_list_unsorted : b.point_list;
    uint32_t* vb = (passes & 1) ? b.point_list : b.point_list_unsorted;
    const int blocks = (P + EMIT_TILE - 1) / EMIT_TILE;
    LG_CUDA(cudaMemsetAsync(g.emit_scan_state, 0, sizeof(unsigned long long) * (size_t)blocks, stream));
    bool in_b = false;
    int rc;
    // the emission kernel accumulates the digit histograms of the tile ids it writes (no separate histogram pass);
    // beyond 16 tile-id bits (more than 65535 tiles) the generic path recomputes them
    rc = radix_sort_clear(b.sort_temp, (size_t)R, passes, stream);
    if (rc != LG_OK) return rc;
    const uint32_t mask0 = (1u << (end_bit < 8 ? end_bit : 8)) - 1u;
    const uint32_t mask1 = end_bit > 8 ? (1u << (end_bit - 8 < 8 ? end_bit - 8 : 8)) - 1u : 0u;
    emit_pairs_kernel<<<blocks, EMIT_BLOCK, 0, stream>>>(
        P, g.depth_ids[0], g.tiles_touched, (gx <= 255 && gy <= 255) ? g.rect_packed : nullptr, g.means2D, radii, ka, va,
        gx, gy, g.emit_scan_state, g.counters + 2, radix_sort_hist_ptr(b.sort_temp), mask0, mask1);
    LG_LAUNCH_CHECK(debug, stream);
    if (passes <= 2)
        rc = radix_sort_pairs_u32_prehist(ka, kb, va, vb, (size_t)R, 0, end_bit, b.sort_temp, b.sort_temp_bytes, debug,
                                          stream, &in_b);
    else
        rc = radix_sort_pairs_u32(ka, kb, va, vb, (size_t)R, 0, end_bit, b.sort_temp, b.sort_temp_bytes, debug, stream,
                                  &in_b);
    if (rc != LG_OK) return rc;
    tile_ranges_kernel<<<((R + 3) / 4 + 255) / 256, 256, 0, stream>>>((uint32_t)R, b.tile_keys, img.ranges);
    LG_LAUNCH_CHECK(debug, stream);
    return launch_tile_order(T, reinterpret_cast<const uint32_t*>(img.ranges), 2, img.tile_order, stream);
}

int launch_rebuild_keys(int R, const GeometryState& g, const BinningState& b, unsigned long long* keys_out,
                        cudaStream_t stream) {
    if (R <= 0) return LG_OK;
    rebuild_keys_kernel<<<(R + 255) / 256, 256, 0, stream>>>((uint32_t)R, b.tile_keys, b.point_list, g.depths, keys_out);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

}  // namespace lg
