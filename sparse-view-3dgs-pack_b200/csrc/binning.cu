// binning.cu — tile binning.  Produces what duplicateWithKeys + cub::DeviceRadixSort::SortPairs + cudaMemset +
// identifyTileRanges produce in the reference (DGR/cuda_rasterizer/rasterizer_impl.cu:70-138, 292-321): the list of
// Gaussian ids ordered by (tile, depth bits, Gaussian id) and the [begin, end) range of every tile.
//
// The reference sorts R = sum(tiles_touched) 64-bit (tile | depth) keys globally (6 passes over 12-byte pairs).
// The major key only has T <= a few thousand values, so here it is never sorted at all:
//   1. COUNT    the preprocess kernel adds 1 to tile_count[t] for every tile t of a Gaussian's rectangle;
//   2. SCAN     one block turns the counts into the tile ranges (the reference's identifyTileRanges output), the
//               per-tile write cursors, the total R and the longest-list-first launch order of the tiles;
//   3. SCATTER  every (Gaussian, tile) overlap takes a slot of its tile's range with one atomic on the tile cursor and
//               stores (depth bits, Gaussian id) there — the tile's entries are now contiguous, in arbitrary order;
//   4. SORT     one block per tile sorts its own entries by (depth bits, id) in shared memory (LSD radix over the
//               depth digits that actually vary inside the tile; equal depths are put in ascending-id order) and writes
//               the tile's slice of the id list.
// (tile, depth bits, id) is a total order — a Gaussian appears at most once per tile — so the list is the
// reference's bit for bit whatever order the atomics of step 3 resolved in.  Per-overlap traffic: 8 B written + 8 B
// read + 4 B written, all L2-resident, against 8 + 6 x 24 B in the reference; no P-sized or R-sized global sort pass
// is left.
#include "common.cuh"

namespace lg {

// bits needed to hold values < n (>= 1)
int higher_msb(uint32_t n) {
    int b = 0;
    while (n) { b++; n >>= 1; }
    return b < 1 ? 1 : b;
}

// ------------------------------------------------------------------------------------------------ 2. SCAN
// One block: exclusive scan of the per-tile counts -> ranges / cursors / total, then the tile launch order (1024
// buckets scaled to the longest list, heaviest first: the hardware hands out blocks in blockIdx order, so the long
// lists of the image centre are started first and do not form the tail).  Tiles with no entries keep (begin, begin),
// which the blend kernels treat like the reference's (0, 0): an empty range.
__global__ void __launch_bounds__(1024) tile_scan_kernel(int T, uint32_t* __restrict__ tile_ctr,
                                                         uint2* __restrict__ ranges, uint32_t* __restrict__ order,
                                                         uint32_t* __restrict__ counters) {
    __shared__ uint32_t s_cnt[1024];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_running, s_max;
    const int tid = threadIdx.x;
    const unsigned lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) { s_running = 0; s_max = 1; }
    s_cnt[tid] = 0;
    __syncthreads();
    uint32_t m = 0;
    for (int base = 0; base < T; base += 1024) {
        const int t = base + tid;
        const uint32_t v = t < T ? tile_ctr[(size_t)t * LG_CTR_STRIDE] : 0u;
        m = max(m, v);
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = s_running;
        for (unsigned w = 0; w < warp; w++) before += s_warp[w];
        if (t < T) {
            const uint32_t begin = before + incl - v;
            ranges[t] = v ? make_uint2(begin, begin + v) : make_uint2(0u, 0u);  // empty tiles as in the reference
            tile_ctr[(size_t)t * LG_CTR_STRIDE + 1] = begin;  // write cursor of the scatter step
        }
        __syncthreads();
        if (tid == 1023) s_running = before + incl;
        __syncthreads();
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) atomicMax(&s_max, m);
    if (tid == 0) {
        counters[1] = s_running;  // num_rendered
        counters[4] = 0;          // tile-sort overflow / error word
    }
    __syncthreads();
    if (tid == 0) counters[5] = s_max;  // longest list: tells the large-list sort launch whether it has any work
    const float scale = 1023.0f / (float)s_max;
    for (int t = tid; t < T; t += 1024) atomicAdd(&s_cnt[1023 - min((int)((float)(ranges[t].y - ranges[t].x) * scale), 1023)], 1u);
    __syncthreads();
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
        const uint32_t pos = atomicAdd(&s_cnt[1023 - min((int)((float)(ranges[t].y - ranges[t].x) * scale), 1023)], 1u);
        order[pos] = (uint32_t)t;
    }
}

// ------------------------------------------------------------------------------------------------ 3. SCATTER
#define SCATTER_BLOCK 256
#define SCATTER_COOP 8  // rectangles of more tiles than this are written by the whole warp

// One thread per Gaussian; a slot of the tile's range per overlap (rasterizer_impl.cu:93-108 writes the same pairs at
// scan offsets; the order inside a tile is settled by the sort that follows).  Does nothing when the capacity the
// caller's binning buffer was sized for turned out too small (the host then re-runs this step with a larger buffer).
__global__ void __launch_bounds__(SCATTER_BLOCK) scatter_pairs_kernel(int P, const uint32_t* __restrict__ tiles_touched,
                                                                      const uint32_t* __restrict__ rect_packed,
                                                                      const float2* __restrict__ xy,
                                                                      const int* __restrict__ radii,
                                                                      const float* __restrict__ depths,
                                                                      uint32_t* __restrict__ tile_ctr,
                                                                      uint2* __restrict__ pairs, int grid_x, int grid_y,
                                                                      const uint32_t* __restrict__ counters,
                                                                      uint32_t capacity) {
    if (counters[1] > capacity) return;
    const int idx = blockIdx.x * SCATTER_BLOCK + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0, key = 0;
    if (idx < P) {
        if (rect_packed) {  // 0 for a Gaussian that emits nothing
            const uint32_t r = rect_packed[idx];
            x0 = r & 255u; y0 = (r >> 8) & 255u; x1 = (r >> 16) & 255u; y1 = r >> 24;
        } else if (tiles_touched[idx] > 0) {
            const float2 p = xy[idx];
            lg_get_rect(p.x, p.y, radii[idx], grid_x, grid_y, x0, y0, x1, y1);
        }
    }
    const uint32_t n = (x1 - x0) * (y1 - y0);
    if (n > 0) key = __float_as_uint(depths[idx]);
    if (n > 0 && n <= SCATTER_COOP) {
        // all of the rectangle's cursor atomics are issued before the first result is used: up to SCATTER_COOP L2 round
        // trips in flight per thread instead of one after the other
        uint32_t slot[SCATTER_COOP];
        uint32_t x = x0, y = y0;
#pragma unroll
        for (int k = 0; k < SCATTER_COOP; k++) {
            slot[k] = 0xffffffffu;
            if ((uint32_t)k < n) slot[k] = atomicAdd(&tile_ctr[(size_t)(y * (uint32_t)grid_x + x) * LG_CTR_STRIDE + 1], 1u);
            if (++x == x1) { x = x0; y++; }
        }
#pragma unroll
        for (int k = 0; k < SCATTER_COOP; k++)
            if ((uint32_t)k < n) pairs[slot[k]] = make_uint2(key, (uint32_t)idx);
    }
    unsigned big = __ballot_sync(0xffffffffu, n > SCATTER_COOP);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const uint32_t bw = __shfl_sync(0xffffffffu, x1, src) - bx0;
        const uint32_t bn = __shfl_sync(0xffffffffu, n, src);
        const uint32_t bkey = __shfl_sync(0xffffffffu, key, src);
        const uint32_t bid = (uint32_t)(idx - (int)lane + src);
        for (uint32_t k = lane; k < bn; k += 32) {
            const uint32_t slot = atomicAdd(&tile_ctr[(size_t)((by0 + k / bw) * (uint32_t)grid_x + bx0 + k % bw) * LG_CTR_STRIDE + 1], 1u);
            pairs[slot] = make_uint2(bkey, bid);
        }
    }
}

// The same step with the block's overlaps first counted per tile in shared memory: one global atomic per (block, tile)
// reserves a run of slots for all of the block's entries of that tile (a 4096-Gaussian block holds ~6 entries per tile
// on the benchmark scene, so six times fewer global atomics than above), and the entries then take consecutive slots
// of their run with shared-memory atomics.  Needs 8 bytes of shared memory per tile.
#ifndef SCATTER_AGGREGATE
#define SCATTER_AGGREGATE 1
#endif
#define SCATTER_AGG_BLOCK 1024
#define SCATTER_AGG_IPT 4
template <bool PLACE>
__device__ __forceinline__ void scatter_walk(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1, uint32_t n, uint32_t key,
                                             uint32_t id, int grid_x, unsigned lane, uint32_t* s_cnt,
                                             const uint32_t* s_base, uint2* __restrict__ pairs) {
    if (n > 0 && n <= SCATTER_COOP) {
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = x0; x < x1; x++) {
                const uint32_t t = y * (uint32_t)grid_x + x;
                const uint32_t r = atomicAdd(&s_cnt[t], 1u);
                if (PLACE) pairs[s_base[t] + r] = make_uint2(key, id);
            }
    }
    unsigned big = __ballot_sync(0xffffffffu, n > SCATTER_COOP);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const uint32_t bw = __shfl_sync(0xffffffffu, x1, src) - bx0;
        const uint32_t bn = __shfl_sync(0xffffffffu, n, src);
        const uint32_t bkey = __shfl_sync(0xffffffffu, key, src), bid = __shfl_sync(0xffffffffu, id, src);
        for (uint32_t k = lane; k < bn; k += 32) {
            const uint32_t t = (by0 + k / bw) * (uint32_t)grid_x + bx0 + k % bw;
            const uint32_t r = atomicAdd(&s_cnt[t], 1u);
            if (PLACE) pairs[s_base[t] + r] = make_uint2(bkey, bid);
        }
    }
}

__global__ void __launch_bounds__(SCATTER_AGG_BLOCK) scatter_pairs_agg_kernel(
    int P, int T, const uint32_t* __restrict__ rect_packed, const float* __restrict__ depths,
    uint32_t* __restrict__ tile_ctr, uint2* __restrict__ pairs, int grid_x, const uint32_t* __restrict__ counters,
    uint32_t capacity) {
    extern __shared__ uint32_t s_scatter[];
    if (counters[1] > capacity) return;
    uint32_t* s_cnt = s_scatter;
    uint32_t* s_base = s_scatter + T;
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    for (int t = tid; t < T; t += SCATTER_AGG_BLOCK) s_cnt[t] = 0;
    uint32_t x0[SCATTER_AGG_IPT], y0[SCATTER_AGG_IPT], x1[SCATTER_AGG_IPT], y1[SCATTER_AGG_IPT], key[SCATTER_AGG_IPT];
    const int first = blockIdx.x * (SCATTER_AGG_BLOCK * SCATTER_AGG_IPT) + (int)tid;
#pragma unroll
    for (int u = 0; u < SCATTER_AGG_IPT; u++) {
        const int idx = first + u * SCATTER_AGG_BLOCK;
        const uint32_t r = idx < P ? rect_packed[idx] : 0u;
        x0[u] = r & 255u; y0[u] = (r >> 8) & 255u; x1[u] = (r >> 16) & 255u; y1[u] = r >> 24;
        key[u] = (x1[u] - x0[u]) * (y1[u] - y0[u]) > 0 ? __float_as_uint(depths[idx]) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SCATTER_AGG_IPT; u++)
        scatter_walk<false>(x0[u], y0[u], x1[u], y1[u], (x1[u] - x0[u]) * (y1[u] - y0[u]), key[u],
                            (uint32_t)(first + u * SCATTER_AGG_BLOCK), grid_x, lane, s_cnt, s_base, pairs);
    __syncthreads();
    for (int t = tid; t < T; t += SCATTER_AGG_BLOCK) {
        const uint32_t c = s_cnt[t];
        if (c) s_base[t] = atomicAdd(&tile_ctr[(size_t)t * LG_CTR_STRIDE + 1], c);
        s_cnt[t] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < SCATTER_AGG_IPT; u++)
        scatter_walk<true>(x0[u], y0[u], x1[u], y1[u], (x1[u] - x0[u]) * (y1[u] - y0[u]), key[u],
                           (uint32_t)(first + u * SCATTER_AGG_BLOCK), grid_x, lane, s_cnt, s_base, pairs);
}

// ------------------------------------------------------------------------------------------------ 4. SORT
#ifndef TS_MIN_BLOCKS
#define TS_MIN_BLOCKS 4
#endif
constexpr int TS_RADIX = 256;
constexpr int TS_MAX_IPT = 16;
constexpr int TS_MAX_GROUPS = 256;     // groups of 3+ entries queued per tile for the warps to rank
constexpr int TS_MAX_GROUP_LEN = 128;  // longest group ranked that way (4 entries per lane)

template <int THREADS>
struct TsSmem {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int MAXIPT = THREADS >= 1024 ? 8 : TS_MAX_IPT;  // entries per thread held in registers
    static constexpr int HALF = THREADS * MAXIPT;                    // longest list sorted in one go (= items held)
    // longest list the block takes: 256 threads sort two halves one after the other — the first goes back to global
    // memory sorted, the second stays in shared memory — and merge them on the way out.  Holding both halves in shared
    // memory (64 KB of items) limited the kernel to 2 blocks per SM; 32 KB gives 3 (then registers limit).
    static constexpr int CAP = THREADS >= 1024 ? HALF : 2 * HALF;
    uint32_t warp_hist[WARPS][TS_RADIX];  // per warp and digit: running count, then exclusive prefix over warps
    uint32_t excl[TS_RADIX];              // per digit: first slot inside the chunk
    uint32_t base[TS_RADIX];              // multi-chunk passes: running global offset of every digit
    uint32_t warp_sums[32];
    uint32_t red[8];                      // [0] OR, [1] AND of the keys, [2] full-sort fallback, [3] min, [4] max, [5] #groups
    uint32_t grp_start[TS_MAX_GROUPS];    // long groups left to the whole block (ts_fix_groups)
    uint32_t grp_len[TS_MAX_GROUPS];
    uint2 items[HALF];
};

template <int SEL>
__device__ __forceinline__ uint32_t ts_key(const uint2& v) { return SEL == 0 ? v.x : v.y; }

// Stable ranking of the block's chunk by one 8-bit digit.  Items are held warp-striped: warp w owns elements
// [w * 32 * IPT, (w + 1) * 32 * IPT) of the chunk and item i of a lane is element w * 32 * IPT + i * 32 + lane, so
// (warp, i, lane) order is element order.  On return rank[i] is the item's position among the warp's items of the same
// digit, s.warp_hist[w][d] the number of such items in earlier warps, and (threads < 256) the return value the
// chunk's count of digit `tid`.  The lanes sharing a digit are found with one ballot per digit bit (independent of
// each other, so they pipeline; MATCH.ANY is far slower on sm_100a).
template <int THREADS, int IPT, int SEL>
__device__ __forceinline__ uint32_t ts_rank(TsSmem<THREADS>& s, const uint2 (&it)[IPT], uint32_t n, int shift,
                                            uint32_t (&rank)[IPT], uint32_t sub = 0u) {
    constexpr int WARPS = THREADS / 32;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t warp_base = warp * 32u * IPT;
    for (int k = tid; k < WARPS * TS_RADIX; k += THREADS) (&s.warp_hist[0][0])[k] = 0;
    __syncthreads();
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        const bool valid = g < n;
        const uint32_t d = ((ts_key<SEL>(it[i]) - sub) >> shift) & 255u;
        unsigned p = __ballot_sync(0xffffffffu, valid);
        if (p == 0u) { rank[i] = 0; continue; }  // warp-uniform: nothing left in this warp's slice
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const bool bit = (d >> b) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, bit);
            p &= bit ? m : ~m;
        }
        const unsigned before = p & lt_mask;
        uint32_t pre = 0;
        if (valid) pre = s.warp_hist[warp][d];
        __syncwarp();
        if (valid && before == 0) s.warp_hist[warp][d] = pre + __popc(p);
        __syncwarp();
        rank[i] = pre + __popc(before);
    }
    __syncthreads();
    uint32_t count = 0;
    if (tid < TS_RADIX) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const uint32_t c = s.warp_hist[w][tid];
            s.warp_hist[w][tid] = count;
            count += c;
        }
    }
    return count;
}

// exclusive scan over the 256 digits of `count` (held by threads < 256) -> s.excl; ends with a block barrier
template <int THREADS>
__device__ __forceinline__ void ts_digit_scan(TsSmem<THREADS>& s, uint32_t count, uint32_t* dst, uint32_t offset) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (tid < TS_RADIX && lane == 31) s.warp_sums[warp] = incl;
    __syncthreads();
    if (tid < TS_RADIX) {
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < TS_RADIX / 32; w++) wbase += (w < (int)warp) ? s.warp_sums[w] : 0u;
        dst[tid] = offset + wbase + incl - count;
    }
    __syncthreads();
}

// one in-shared-memory pass: items (registers, warp-striped) -> s.items in digit order -> back into the registers
template <int THREADS, int IPT, int SEL>
__device__ __forceinline__ void ts_pass_smem(TsSmem<THREADS>& s, uint2* items, uint2 (&it)[IPT], uint32_t n, int shift,
                                             uint32_t sub = 0u) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t warp_base = warp * 32u * IPT;
    uint32_t rank[IPT];
    const uint32_t count = ts_rank<THREADS, IPT, SEL>(s, it, n, shift, rank, sub);
    ts_digit_scan<THREADS>(s, count, s.excl, 0u);
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        if (g < n) {
            const uint32_t d = ((ts_key<SEL>(it[i]) - sub) >> shift) & 255u;
            items[s.excl[d] + s.warp_hist[warp][d] + rank[i]] = it[i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        if (g < n) it[i] = items[g];
    }
}

// The reference's order inside a tile is (depth bits, Gaussian id).  After the window passes the list is ordered by
// (depth - min) >> lo only: entries of one group (equal window value — usually one entry, sometimes two) sit next to
// each other in arbitrary order.  The thread that finds the start of a group orders a pair on the spot and queues
// anything longer (equal depths: coincident points, a scene clipped to a box seen face-on); the warps of the block
// then take the queued groups in turn and rank them by counting — every lane counts the group members that precede
// its entry, reading the group as warp-wide broadcasts — which keeps the barrier behind this step short (a serial
// insertion sort of a 30-entry group by one thread, 6000 dependent shared-memory cycles, held up the other 255).
// Returns true when a group is longer than TS_MAX_GROUP_LEN or the queue overflows: the caller then falls back to the
// full LSD sort.
template <int THREADS>
__device__ __forceinline__ bool ts_fix_groups(TsSmem<THREADS>& s, uint2* items, uint32_t n, uint32_t sub, int lo) {
    constexpr int WARPS = THREADS / 32;
    constexpr int PER_LANE = TS_MAX_GROUP_LEN / 32;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t q = threadIdx.x; q + 1 < n; q += THREADS) {
        const uint2 me = items[q], next = items[q + 1];
        const uint32_t k = (me.x - sub) >> lo;
        if (((next.x - sub) >> lo) != k || (q > 0 && ((items[q - 1].x - sub) >> lo) == k)) continue;
        uint32_t L = 2;
        while (q + L < n && L <= TS_MAX_GROUP_LEN && ((items[q + L].x - sub) >> lo) == k) L++;
        if (L == 2) {
            if (me.x > next.x || (me.x == next.x && me.y > next.y)) {
                items[q] = next;
                items[q + 1] = me;
            }
            continue;
        }
        const uint32_t slot = L <= TS_MAX_GROUP_LEN ? atomicAdd(&s.red[5], 1u) : (uint32_t)TS_MAX_GROUPS;
        if (slot < (uint32_t)TS_MAX_GROUPS) {
            s.grp_start[slot] = q;
            s.grp_len[slot] = L;
        } else {
            s.red[2] = 1u;
        }
    }
    __syncthreads();
    if (s.red[2] != 0u) return true;
    const uint32_t groups = s.red[5];
    for (uint32_t gi = warp; gi < groups; gi += WARPS) {
        const uint32_t start = s.grp_start[gi], L = s.grp_len[gi];
        uint2 mine[PER_LANE];
        uint32_t where[PER_LANE];
#pragma unroll
        for (int u = 0; u < PER_LANE; u++) {
            const uint32_t j = lane + 32u * u;
            mine[u] = j < L ? items[start + j] : make_uint2(0xffffffffu, 0xffffffffu);
            where[u] = 0;
        }
        for (uint32_t i = 0; i < L; i++) {
            const uint2 w = items[start + i];  // same address for the whole warp: a broadcast
#pragma unroll
            for (int u = 0; u < PER_LANE; u++) {
                if (u * 32u < L) where[u] += (w.x < mine[u].x || (w.x == mine[u].x && w.y < mine[u].y)) ? 1u : 0u;
            }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < PER_LANE; u++) {
            const uint32_t j = lane + 32u * u;
            if (j < L) items[start + where[u]] = mine[u];
        }
    }
    __syncthreads();
    return false;
}

// Sort of one list that fits the block's shared memory.  Fast path: two 8-bit ballot-ranked passes over a 16-bit
// window of the depth keys — the top 16 bits of (key - min), i.e. 32769..65536 distinct values across the tile's
// depth range — after which the few entries that share a window value are put right locally (ts_fix_groups).
// (Shared-memory atomics would make a counting sort trivial, but ATOMS retires one lane every two cycles per SM: measured
// 4x slower than nine ballots per entry and digit.)
// Fallback (long runs of equal / nearly equal depths): full stable LSD sort, id digits first, then every depth digit
// that varies inside the tile.
template <int THREADS, int IPT>
__device__ __noinline__ void ts_sort_in_smem(TsSmem<THREADS>& s, uint2* items, const uint2* __restrict__ src,
                                                uint32_t n, uint32_t* __restrict__ out, int id_bits) {
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t warp_base = warp * 32u * IPT;
    uint2 it[IPT];
    uint32_t vor = 0u, vand = 0xffffffffu, vmin = 0xffffffffu, vmax = 0u;
#pragma unroll
    for (int i = 0; i < IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        it[i] = make_uint2(0u, 0u);
        if (g < n) {
            it[i] = src[g];
            vor |= it[i].x;
            vand &= it[i].x;
            vmin = min(vmin, it[i].x);
            vmax = max(vmax, it[i].x);
        }
    }
    vor = __reduce_or_sync(0xffffffffu, vor);
    vand = __reduce_and_sync(0xffffffffu, vand);
    vmin = __reduce_min_sync(0xffffffffu, vmin);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if (lane == 0) {
        atomicOr(&s.red[0], vor);
        atomicAnd(&s.red[1], vand);
        atomicMin(&s.red[3], vmin);
        atomicMax(&s.red[4], vmax);
    }
    __syncthreads();
    const uint32_t kmin = s.red[3], span = s.red[4] - kmin;
    int lo = 0;
    if (span == 0u) {  // all depths equal (or a single entry): one group
#pragma unroll
        for (int i = 0; i < IPT; i++) {
            const uint32_t g = warp_base + i * 32 + lane;
            if (g < n) items[g] = it[i];
        }
        __syncthreads();
    } else {
        lo = max(0, (31 - __clz(span)) - 15);
        ts_pass_smem<THREADS, IPT, 0>(s, items, it, n, lo, kmin);
        if ((span >> lo) > 255u) ts_pass_smem<THREADS, IPT, 0>(s, items, it, n, lo + 8, kmin);
    }
    if (ts_fix_groups<THREADS>(s, items, n, kmin, lo)) {
        const uint32_t varying = s.red[0] ^ s.red[1];  // depth bits that differ inside this tile
#pragma unroll
        for (int i = 0; i < IPT; i++) {
            const uint32_t g = warp_base + i * 32 + lane;
            if (g < n) it[i] = items[g];
        }
#pragma unroll 1
        for (int shift = 0; shift < id_bits; shift += 8) ts_pass_smem<THREADS, IPT, 1>(s, items, it, n, shift);
#pragma unroll 1
        for (int shift = 0; shift < 32; shift += 8) {
            if (((varying >> shift) & 255u) == 0u) continue;  // a digit every key shares: nothing to reorder
            ts_pass_smem<THREADS, IPT, 0>(s, items, it, n, shift);
        }
        __syncthreads();
    }
    if (out != nullptr)
        for (uint32_t q = tid; q < n; q += THREADS) out[q] = items[q].y;
}

// Lists of HALF < n <= 2 * HALF entries (256-thread blocks): the two halves of the scattered list are sorted one
// after the other; the first goes back, sorted, to the tile's slice of the second pair buffer (plain global loads and
// stores of the same block, ordered by the block barrier; it stays in L2), the second remains in shared memory, and
// the two are merged on the way out.  Every thread produces a
// run of consecutive output slots: a merge-path binary search finds how many entries of each half precede its run,
// then it merges serially (keys are (depth bits, id) pairs: distinct, so the merge is unambiguous).
__device__ __forceinline__ bool ts_pair_less(const uint2& a, const uint2& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); }

template <int THREADS>
__device__ __forceinline__ void ts_merge_out(const uint2* A, uint32_t nA, const uint2* B, uint32_t nB,
                                             uint32_t* out) {
    const uint32_t n = nA + nB;
    const uint32_t per = (n + THREADS - 1) / THREADS;
    const uint32_t d0 = min(threadIdx.x * per, n), d1 = min(d0 + per, n);
    // largest i with i + j = d0 such that the first i entries of A and j of B are the d0 smallest
    uint32_t lo = d0 > nB ? d0 - nB : 0u, hi = min(d0, nA);
    while (lo < hi) {
        const uint32_t i = (lo + hi) >> 1;   // candidate: take i from A, d0 - i from B; too few from A if A[i] < B[d0-i-1]
        if (ts_pair_less(A[i], B[d0 - i - 1])) lo = i + 1;
        else hi = i;
    }
    uint32_t i = lo, j = d0 - lo;
    uint2 a = i < nA ? A[i] : make_uint2(0xffffffffu, 0xffffffffu);
    uint2 b = j < nB ? B[j] : make_uint2(0xffffffffu, 0xffffffffu);
    for (uint32_t k = d0; k < d1; k++) {
        const bool take_a = j >= nB || (i < nA && ts_pair_less(a, b));
        out[k] = take_a ? a.y : b.y;
        if (take_a) { i++; a = i < nA ? A[i] : make_uint2(0xffffffffu, 0xffffffffu); }
        else { j++; b = j < nB ? B[j] : make_uint2(0xffffffffu, 0xffffffffu); }
    }
}

// Lists longer than the shared-memory capacity: the same LSD passes streamed through global memory, chunk by chunk,
// between the tile's slices of the two pair buffers (id digits first, then the varying depth digits: no tie fix-up).
template <int THREADS>
__device__ void ts_sort_streamed(TsSmem<THREADS>& s, uint2* a, uint2* b, uint32_t n, uint32_t* __restrict__ out,
                                 int id_bits) {
    constexpr int IPT = TsSmem<THREADS>::MAXIPT;
    constexpr uint32_t CAP = TsSmem<THREADS>::CAP;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t warp_base = warp * 32u * IPT;
    uint32_t vor = 0u, vand = 0xffffffffu;
    for (uint32_t g = tid; g < n; g += THREADS) {
        const uint32_t k = a[g].x;
        vor |= k;
        vand &= k;
    }
    vor = __reduce_or_sync(0xffffffffu, vor);
    vand = __reduce_and_sync(0xffffffffu, vand);
    if (lane == 0) {
        atomicOr(&s.red[0], vor);
        atomicAnd(&s.red[1], vand);
    }
    __syncthreads();
    const uint32_t varying = s.red[0] ^ s.red[1];
    const int id_passes = (id_bits + 7) / 8;
    uint2* src = a;
    uint2* dst = b;
#pragma unroll 1
    for (int pass = 0; pass < id_passes + 4; pass++) {
        const bool by_id = pass < id_passes;
        const int shift = by_id ? 8 * pass : 8 * (pass - id_passes);
        if (!by_id && ((varying >> shift) & 255u) == 0u) continue;
        // digit totals of the whole list -> where every digit's run starts
        if (tid < TS_RADIX) s.base[tid] = 0;
        __syncthreads();
        for (uint32_t g = tid; g < n; g += THREADS) {
            const uint2 v = src[g];
            atomicAdd(&s.base[((by_id ? v.y : v.x) >> shift) & 255u], 1u);
        }
        __syncthreads();
        {
            const uint32_t c = tid < TS_RADIX ? s.base[tid] : 0u;
            __syncthreads();
            ts_digit_scan<THREADS>(s, c, s.base, 0u);
        }
        for (uint32_t c0 = 0; c0 < n; c0 += CAP) {
            const uint32_t cn = min(CAP, n - c0);
            uint2 it[IPT];
            uint32_t rank[IPT];
#pragma unroll
            for (int i = 0; i < IPT; i++) {
                const uint32_t g = warp_base + i * 32 + lane;
                it[i] = g < cn ? src[c0 + g] : make_uint2(0u, 0u);
            }
            const uint32_t count = by_id ? ts_rank<THREADS, IPT, 1>(s, it, cn, shift, rank)
                                         : ts_rank<THREADS, IPT, 0>(s, it, cn, shift, rank);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < IPT; i++) {
                const uint32_t g = warp_base + i * 32 + lane;
                if (g < cn) {
                    const uint32_t d = ((by_id ? it[i].y : it[i].x) >> shift) & 255u;
                    dst[s.base[d] + s.warp_hist[warp][d] + rank[i]] = it[i];
                }
            }
            __syncthreads();
            if (tid < TS_RADIX) s.base[tid] += count;
            __syncthreads();
        }
        uint2* t = src; src = dst; dst = t;
        __threadfence_block();
        __syncthreads();
    }
    for (uint32_t q = tid; q < n; q += THREADS) out[q] = src[q].y;
}

// One block per tile, longest lists first.  LARGE = false: 256 threads, lists of 1..8192 entries (above 4096 as two
// sorted halves merged on the way out); LARGE = true: a few persistent 1024-thread blocks for the lists above 8192,
// streamed through global memory.
template <int THREADS, bool LARGE>
__global__ void __launch_bounds__(THREADS, LARGE ? 1 : TS_MIN_BLOCKS) tile_sort_kernel(const uint2* __restrict__ ranges,
                                                            const uint32_t* __restrict__ order, uint2* pairs,
                                                            uint2* pairs_alt, uint32_t* __restrict__ point_list,
                                                            int id_bits, const uint32_t* __restrict__ counters,
                                                            uint32_t capacity, int T) {
    extern __shared__ __align__(16) unsigned char ts_smem_raw[];
    TsSmem<THREADS>& s = *reinterpret_cast<TsSmem<THREADS>*>(ts_smem_raw);
    if (counters[1] > capacity) return;
    constexpr uint32_t SMALL_CAP = TsSmem<256>::CAP;  // what the 256-thread blocks take (two halves + merge)
    if (LARGE && counters[5] <= SMALL_CAP) return;    // no list for this launch (the usual case): nothing to walk
    // LARGE: a few persistent blocks walk the launch order and pick out the long lists (empty 1024-thread blocks with
    // 82 KB of shared memory cost ~6 us per wave to schedule, so one block per tile is not an option here)
    for (uint32_t slot = blockIdx.x; slot < (uint32_t)T; slot += gridDim.x) {
    const uint32_t tile = order[slot];
    const uint2 range = ranges[tile];
    const uint32_t n = range.y - range.x;
    if (LARGE ? (n <= SMALL_CAP) : (n == 0 || n > SMALL_CAP)) continue;
    __syncthreads();  // the previous list of this block is finished with the shared state
    if (threadIdx.x == 0) {
        s.red[0] = 0u;
        s.red[1] = 0xffffffffu;
        s.red[2] = 0u;
        s.red[3] = 0xffffffffu;
        s.red[4] = 0u;
        s.red[5] = 0u;
    }
    __syncthreads();
    const uint2* src = pairs + range.x;
    uint32_t* out = point_list + range.x;
    if constexpr (LARGE) {
        if (n > TsSmem<THREADS>::CAP) {
            ts_sort_streamed<THREADS>(s, pairs + range.x, pairs_alt + range.x, n, out, id_bits);
            continue;
        }
    }
    if constexpr (LARGE) {
        ts_sort_in_smem<THREADS, TsSmem<THREADS>::MAXIPT>(s, s.items, src, n, out, id_bits);
    } else {
        constexpr uint32_t HALF = TsSmem<THREADS>::HALF;
        if (n > HALF) {
            const uint32_t nA = n / 2, nB = n - nA;  // both <= HALF
            ts_sort_in_smem<THREADS, 16>(s, s.items, src, nA, nullptr, id_bits);
            __syncthreads();
            uint2* sortedA = pairs_alt + range.x;   // this tile's slice of the second pair buffer (L2-resident)
            for (uint32_t q = threadIdx.x; q < nA; q += THREADS) sortedA[q] = s.items[q];
            if (threadIdx.x == 0) {   // fresh reduction state for the second half
                s.red[0] = 0u; s.red[1] = 0xffffffffu; s.red[2] = 0u; s.red[3] = 0xffffffffu; s.red[4] = 0u; s.red[5] = 0u;
            }
            __syncthreads();
            ts_sort_in_smem<THREADS, 16>(s, s.items, src + nA, nB, nullptr, id_bits);
            __syncthreads();   // also orders the block's global writes of sortedA before the reads below
            ts_merge_out<THREADS>(sortedA, nA, s.items, nB, out);
            continue;
        }
        const uint32_t per_thread = (n + THREADS - 1) / THREADS;
        if (per_thread <= 2) ts_sort_in_smem<THREADS, 2>(s, s.items, src, n, out, id_bits);
        else if (per_thread <= 4) ts_sort_in_smem<THREADS, 4>(s, s.items, src, n, out, id_bits);
        else if (per_thread <= 8) ts_sort_in_smem<THREADS, 8>(s, s.items, src, n, out, id_bits);
        else ts_sort_in_smem<THREADS, 16>(s, s.items, src, n, out, id_bits);
    }
    }
}

// inspection only (lg_state_read "point_list_keys"): the reference's sorted 64-bit keys, one block per tile
__global__ void __launch_bounds__(256) rebuild_keys_kernel(const uint2* __restrict__ ranges,
                                                           const uint32_t* __restrict__ point_list,
                                                           const float* __restrict__ depths,
                                                           unsigned long long* __restrict__ keys) {
    const uint2 r = ranges[blockIdx.x];
    for (uint32_t i = r.x + threadIdx.x; i < r.y; i += 256)
        keys[i] = ((unsigned long long)blockIdx.x << 32) | __float_as_uint(depths[point_list[i]]);
}

// step 2 (queued right behind the preprocess kernel; needs nothing from the host)
int launch_tile_scan(int W, int H, ImageState& img, bool debug, cudaStream_t stream) {
    const int T = num_tiles_x(W) * num_tiles_y(H);
    tile_scan_kernel<<<1, 1024, 0, stream>>>(T, img.tile_ctr, img.ranges, img.tile_order, img.counters);
    LG_LAUNCH_CHECK(debug, stream);
    return LG_OK;
}

// steps 3 and 4 for a binning buffer of `capacity` entries (kernels are no-ops if num_rendered exceeds it)
int launch_binning(int P, int capacity, int W, int H, const GeometryState& g, const int* radii, BinningState& b,
                   ImageState& img, bool debug, cudaStream_t stream) {
    const int gx = num_tiles_x(W), gy = num_tiles_y(H);
    const int T = gx * gy;
    if (capacity <= 0 || P <= 0) return LG_OK;
    const size_t agg_smem = 2 * sizeof(uint32_t) * (size_t)T;
    if (SCATTER_AGGREGATE && gx <= 255 && gy <= 255 && agg_smem <= 160 * 1024 && P >= 64 * 1024) {
        LG_CUDA(cudaFuncSetAttribute(scatter_pairs_agg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem));
        const int per_block = SCATTER_AGG_BLOCK * SCATTER_AGG_IPT;
        scatter_pairs_agg_kernel<<<(P + per_block - 1) / per_block, SCATTER_AGG_BLOCK, agg_smem, stream>>>(
            P, T, g.rect_packed, g.depths, img.tile_ctr, b.pairs, gx, img.counters, (uint32_t)capacity);
    } else {
        scatter_pairs_kernel<<<(P + SCATTER_BLOCK - 1) / SCATTER_BLOCK, SCATTER_BLOCK, 0, stream>>>(
            P, g.tiles_touched, (gx <= 255 && gy <= 255) ? g.rect_packed : nullptr, g.means2D, radii, g.depths,
            img.tile_ctr, b.pairs, gx, gy, img.counters, (uint32_t)capacity);
    }
    LG_LAUNCH_CHECK(debug, stream);
    const int id_bits = higher_msb((uint32_t)(P > 1 ? P - 1 : 1));
    const size_t smem_small = sizeof(TsSmem<256>), smem_large = sizeof(TsSmem<1024>);
    LG_CUDA(cudaFuncSetAttribute(tile_sort_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small));
    LG_CUDA(cudaFuncSetAttribute(tile_sort_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_large));
    tile_sort_kernel<256, false><<<T, 256, smem_small, stream>>>(img.ranges, img.tile_order, b.pairs, b.pairs_alt,
                                                                b.point_list, id_bits, img.counters, (uint32_t)capacity, T);
    LG_LAUNCH_CHECK(debug, stream);
    if ((size_t)capacity > (size_t)TsSmem<256>::CAP) {
        tile_sort_kernel<1024, true><<<min(T, LG_NUM_SMS), 1024, smem_large, stream>>>(
            img.ranges, img.tile_order, b.pairs, b.pairs_alt, b.point_list, id_bits, img.counters, (uint32_t)capacity, T);
        LG_LAUNCH_CHECK(debug, stream);
    }
    return LG_OK;
}

int launch_rebuild_keys(int W, int H, const GeometryState& g, const BinningState& b, const ImageState& img,
                        unsigned long long* keys_out, cudaStream_t stream) {
    const int T = num_tiles_x(W) * num_tiles_y(H);
    rebuild_keys_kernel<<<T, 256, 0, stream>>>(img.ranges, b.point_list, g.depths, keys_out);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

}  // namespace lg
