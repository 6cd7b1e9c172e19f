// radix_sort.cu — stable LSD radix sort of (key, u32 value) pairs, onesweep style: one up-front histogram of
// every digit place, then ONE read + ONE write of the pairs per 8-bit digit, with the per-tile digit offsets
// resolved by a decoupled look-back chain instead of a separate scan pass.
// Replaces cub::DeviceRadixSort::SortPairs at KNN/simple_knn.cu:211-214 (u32 Morton codes of distCUDA2).  The
// rasterizer's tile binning (DGR/cuda_rasterizer/rasterizer_impl.cu:306-311) no longer sorts globally: csrc/binning.cu.
#include "common.cuh"

namespace lg {

constexpr int RS_BITS = 8;
constexpr int RS_RADIX = 1 << RS_BITS;
constexpr int RS_THREADS = 256;  // == RS_RADIX: thread d owns digit d in the scan / look-back steps
constexpr int RS_WARPS = RS_THREADS / 32;
// items per thread.  Measured on B200 (1 M and 3.6 M pairs): pass time ~ 12 us + 26 ns per tile, i.e. it grows with
// the NUMBER of tiles (2: 62/220 us, 4: 40/122, 8: 30/77, 16: 20/53) — when a whole wave of tiles starts together
// every tile has to add up the aggregates of all its predecessors in the wave, so the look-back traffic is quadratic
// in the wave size and fat tiles win.
#ifndef RS_IPT
#define RS_IPT 16
#endif
constexpr int RS_TILE = RS_THREADS * RS_IPT;
constexpr int RS_MAX_PASSES = 8;
constexpr int RS_LB_WIN = 16;     // group descriptors fetched per round trip
constexpr int RS_GROUP = 32;      // tiles per look-back group    // look-back descriptors fetched per round trip

constexpr uint32_t LB_PARTIAL = 1u << 30;
constexpr uint32_t LB_INCLUSIVE = 2u << 30;
constexpr uint32_t LB_FLAGS = 3u << 30;
constexpr uint32_t LB_VALUE = ~LB_FLAGS;

// ------------------------------------------------------------------ up-front histogram of all digit places
template <typename K>
__global__ void __launch_bounds__(256) rs_histogram_kernel(const K* __restrict__ keys, uint32_t n, int begin_bit,
                                                           int end_bit, int passes, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[RS_MAX_PASSES * RS_RADIX];
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const K key = keys[i];
#pragma unroll
        for (int p = 0; p < RS_MAX_PASSES; p++) {
            if (p < passes) {
                const int shift = begin_bit + p * RS_BITS;
                const int nb = min(RS_BITS, end_bit - shift);
                const uint32_t d = (uint32_t)(key >> shift) & ((1u << nb) - 1u);
                atomicAdd(&s_hist[p * RS_RADIX + d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RS_RADIX; i += blockDim.x) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// ------------------------------------------------------------------ one digit pass
template <typename K>
struct RsSmem {
    uint32_t warp_hist[RS_WARPS][RS_RADIX];
    uint32_t tile_excl[RS_RADIX];
    uint32_t digit_off[RS_RADIX];
    uint32_t warp_sums[RS_WARPS];
    uint32_t hist_sums[RS_WARPS];
    uint32_t tile_id;
    uint32_t vals[RS_TILE];
    K keys[RS_TILE];
};

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) rs_onesweep_kernel(const K* __restrict__ keys_in, K* __restrict__ keys_out,
                                                                 const uint32_t* __restrict__ vals_in,
                                                                 uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                                 uint32_t digit_mask,
                                                                 const uint32_t* __restrict__ global_hist,
                                                                 volatile uint32_t* lookback,
                                                                 volatile uint32_t* group_desc, uint32_t* ticket) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem<K>& s = *reinterpret_cast<RsSmem<K>*>(rs_smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid == 0) s.tile_id = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) s.warp_hist[w][tid] = 0;
    // exclusive scan of this digit place's 256-bin histogram = where every digit's run starts in the output (each
    // block redoes this 256-element scan instead of a separate launch between the histogram and the passes)
    const uint32_t hist_count = global_hist[tid];
    uint32_t hist_incl = hist_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, hist_incl, o);
        if (lane >= (unsigned)o) hist_incl += t;
    }
    if (lane == 31) s.hist_sums[warp] = hist_incl;
    __syncthreads();
    uint32_t global_offset = hist_incl - hist_count;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) global_offset += (w < (int)warp) ? s.hist_sums[w] : 0u;
    const uint32_t tile = s.tile_id;
    const uint32_t tile_base = tile * RS_TILE;
    const uint32_t warp_base = tile_base + warp * (32 * RS_IPT);

    // warp-striped load: item i of this lane is element warp_base + i*32 + lane (order = (warp, i, lane))
    K key[RS_IPT];
    uint32_t val[RS_IPT];
#pragma unroll
    for (int i = 0; i < RS_IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        if (g < n) {
            key[i] = keys_in[g];
            val[i] = vals_in[g];
        } else {
            key[i] = (K)0;
            val[i] = 0;
        }
    }

    // rank items inside the warp by digit, in element order (stable).  The lanes sharing an item's digit are found
    // with one ballot per digit bit (independent of each other and of the running counts, so all of them pipeline;
    // MATCH.ANY costs several hundred cycles per call on sm_100a and serialised this loop); only the running per-warp
    // digit counts are updated in order.
    uint32_t rank[RS_IPT];
    unsigned peers[RS_IPT];
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
        unsigned p = __ballot_sync(0xffffffffu, g < n);
#pragma unroll
        for (int b = 0; b < RS_BITS; b++) {
            const bool bit = (d >> b) & 1u;
            const unsigned m = __ballot_sync(0xffffffffu, bit);
            p &= bit ? m : ~m;
        }
        peers[i] = p;
    }
#pragma unroll
    for (int i = 0; i < RS_IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        const bool valid = g < n;
        const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
        const unsigned before = peers[i] & lt_mask;
        uint32_t pre = 0;
        if (valid) pre = s.warp_hist[warp][d];
        __syncwarp();
        if (valid && before == 0) s.warp_hist[warp][d] = pre + __popc(peers[i]);
        __syncwarp();
        rank[i] = pre + __popc(before);
    }
    __syncthreads();

    // thread d: exclusive scan of digit d over warps -> tile count
    uint32_t tile_count = 0;
    {
        const uint32_t d = tid;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t c = s.warp_hist[w][d];
            s.warp_hist[w][d] = tile_count;
            tile_count += c;
        }
    }
    // publish this tile's digit count as early as possible
    lookback[(size_t)tile * RS_RADIX + tid] = tile_count | LB_PARTIAL;

    // exclusive scan of tile_count over digits (in-tile bucket starts)
    uint32_t incl = tile_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) s.warp_sums[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) wbase += (w < (int)warp) ? s.warp_sums[w] : 0u;
    const uint32_t excl_in_tile = wbase + incl - tile_count;
    s.tile_excl[tid] = excl_in_tile;
    __syncthreads();

    // scatter into shared memory in tile-sorted order (before the look-back: the registers holding the items are
    // free while the descriptor loads are in flight)
#pragma unroll
    for (int i = 0; i < RS_IPT; i++) {
        const uint32_t g = warp_base + i * 32 + lane;
        if (g < n) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & digit_mask;
            const uint32_t q = s.tile_excl[d] + s.warp_hist[warp][d] + rank[i];
            s.keys[q] = key[i];
            s.vals[q] = val[i];
        }
    }

    // Two-level decoupled look-back for digit `tid`.  When a whole wave of W tiles starts together nobody holds an
    // inclusive prefix yet, and a flat look-back makes tile t add up the totals of ~t predecessors: O(W^2) descriptor
    // traffic per wave, which is what bounded this pass (measured: time ~ 12 us + 26 ns per tile).  Tiles are
    // therefore grouped by RS_GROUP consecutive tickets:
    //   - a tile adds the totals of the (< RS_GROUP) earlier tiles of its own group (one window of independent loads);
    //   - the last tile of a group publishes the group total, looks back over the earlier GROUP descriptors to the
    //     nearest inclusive one, and publishes the group's inclusive prefix;
    //   - every other tile takes the inclusive prefix of the preceding group.
    // Tickets are taken in launch order, so everything a tile waits for belongs to a running or finished tile.
    uint32_t prev = 0;
    {
        const uint32_t grp = tile / RS_GROUP, l = tile % RS_GROUP;
#pragma unroll
        for (int w0 = 0; w0 < RS_GROUP - 1; w0 += RS_LB_WIN) {  // windows of RS_LB_WIN independent descriptor loads
            if ((uint32_t)w0 < l) {
                uint32_t v[RS_LB_WIN];
#pragma unroll
                for (int w = 0; w < RS_LB_WIN; w++) {
                    v[w] = LB_PARTIAL + 0u;
                    if ((uint32_t)(w0 + w) < l) v[w] = lookback[(size_t)(tile - 1u - (w0 + w)) * RS_RADIX + tid];
                }
#pragma unroll
                for (int w = 0; w < RS_LB_WIN; w++) {
                    if ((uint32_t)(w0 + w) < l) {
                        while ((v[w] & LB_FLAGS) == 0u) v[w] = lookback[(size_t)(tile - 1u - (w0 + w)) * RS_RADIX + tid];
                        prev += v[w] & LB_VALUE;
                    }
                }
            }
        }
        if (l == RS_GROUP - 1) {  // group closer
            const uint32_t group_total = prev + tile_count;
            uint32_t before = 0;
            if (grp > 0) {
                group_desc[(size_t)grp * RS_RADIX + tid] = (group_total & LB_VALUE) | LB_PARTIAL;
                int j = (int)grp - 1;
                bool done = false;
                while (!done) {
                    uint32_t g[RS_LB_WIN];
#pragma unroll
                    for (int w = 0; w < RS_LB_WIN; w++) {
                        g[w] = LB_INCLUSIVE + 0u;
                        if (j - w >= 0) g[w] = group_desc[(size_t)(j - w) * RS_RADIX + tid];
                    }
                    int consumed = 0;
                    bool stalled = false;
#pragma unroll
                    for (int w = 0; w < RS_LB_WIN; w++) {
                        if (!done && !stalled) {
                            if ((g[w] & LB_FLAGS) == 0u) stalled = true;  // not published yet: poll again from here
                            else {
                                before += g[w] & LB_VALUE;
                                consumed++;
                                if (g[w] & LB_INCLUSIVE) done = true;
                            }
                        }
                    }
                    j -= consumed;
                }
            }
            group_desc[(size_t)grp * RS_RADIX + tid] = ((before + group_total) & LB_VALUE) | LB_INCLUSIVE;
            prev += before;
        } else if (grp > 0) {
            uint32_t g;
            do { g = group_desc[(size_t)(grp - 1u) * RS_RADIX + tid]; } while ((g & LB_INCLUSIVE) == 0u);
            prev += g & LB_VALUE;
        }
    }
    s.digit_off[tid] = global_offset + prev - excl_in_tile;
    __syncthreads();

    // coalesced write-out: consecutive threads hold consecutive slots of a digit run
    const uint32_t tile_n = min((uint32_t)RS_TILE, n - tile_base);
#pragma unroll
    for (int i = 0; i < RS_IPT; i++) {
        const uint32_t q = i * RS_THREADS + tid;
        if (q < tile_n) {
            const K k = s.keys[q];
            const uint32_t d = (uint32_t)(k >> shift) & digit_mask;
            const uint32_t dst = s.digit_off[d] + q;
            keys_out[dst] = k;
            vals_out[dst] = s.vals[q];
        }
    }
}

// ------------------------------------------------------------------ host driver
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t radix_sort_temp_bytes(size_t n, int key_bytes) {
    (void)key_bytes;
    const size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    size_t b = 0;
    b += align_up(sizeof(uint32_t) * RS_MAX_PASSES * RS_RADIX, 128);          // histograms / global offsets
    b += align_up(sizeof(uint32_t) * 32, 128);                                // tickets (one per pass)
    b += align_up(sizeof(uint32_t) * RS_MAX_PASSES * tiles * RS_RADIX, 128);  // look-back descriptors (tiles)
    b += align_up(sizeof(uint32_t) * RS_MAX_PASSES * ((tiles + RS_GROUP - 1) / RS_GROUP) * RS_RADIX, 128);  // (groups)
    return b + 128;
}

int radix_sort_num_passes(int begin_bit, int end_bit) { return (end_bit - begin_bit + RS_BITS - 1) / RS_BITS; }

// The digit histograms live at the start of the scratch buffer: counter [pass * 256 + digit].  A producer kernel may
// fill them itself (after radix_sort_clear) and then call the *_prehist entry point, which skips the histogram pass.
uint32_t* radix_sort_hist_ptr(char* temp) {
    char* p = temp;
    uint32_t* hist;
    carve(p, hist, RS_MAX_PASSES * RS_RADIX);
    return hist;
}

int radix_sort_clear(char* temp, size_t n, int passes, cudaStream_t stream) {
    const size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    char* p = temp;
    uint32_t* hist;
    uint32_t* tickets;
    uint32_t* lookback;
    carve(p, hist, RS_MAX_PASSES * RS_RADIX);
    carve(p, tickets, 32);
    carve(p, lookback, RS_MAX_PASSES * tiles * RS_RADIX);
    const size_t groups = (tiles + RS_GROUP - 1) / RS_GROUP;
    uint32_t* group_desc;
    carve(p, group_desc, RS_MAX_PASSES * groups * RS_RADIX);
    const size_t zero_bytes = (size_t)((char*)(lookback + (size_t)passes * tiles * RS_RADIX) - (char*)hist);
    LG_CUDA(cudaMemsetAsync(hist, 0, zero_bytes, stream));
    LG_CUDA(cudaMemsetAsync(group_desc, 0, sizeof(uint32_t) * (size_t)passes * groups * RS_RADIX, stream));
    return LG_OK;
}

// Sorts n pairs on key bits [begin_bit, end_bit).  Input in (keys_a, vals_a); buffers are ping-ponged and BOTH are
// clobbered.  The sorted result lands in (keys_b, vals_b) when the pass count is odd, else in (keys_a, vals_a);
// *result_in_b tells which.
template <typename K>
static int radix_sort_pairs_impl(K* keys_a, K* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n, int begin_bit,
                                 int end_bit, char* temp, size_t temp_bytes, bool debug, cudaStream_t stream,
                                 bool* result_in_b, bool have_hist) {
    const int passes = radix_sort_num_passes(begin_bit, end_bit);
    *result_in_b = (passes & 1) != 0;
    if (n == 0 || passes == 0) {
        *result_in_b = false;
        return LG_OK;
    }
    if (passes > RS_MAX_PASSES) {
        set_error("radix sort: %d passes requested, max %d", passes, RS_MAX_PASSES);
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (n >= (size_t)LB_VALUE) {
        set_error("radix sort: n=%zu exceeds 2^30-1", n);
        return LG_ERR_UNSUPPORTED;
    }
    if (temp_bytes < radix_sort_temp_bytes(n, sizeof(K))) {
        set_error("radix sort: temp buffer too small");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const size_t tiles = (n + RS_TILE - 1) / RS_TILE;
    char* p = temp;
    uint32_t* hist;
    uint32_t* tickets;
    uint32_t* lookback;
    carve(p, hist, RS_MAX_PASSES * RS_RADIX);
    carve(p, tickets, 32);
    carve(p, lookback, RS_MAX_PASSES * tiles * RS_RADIX);
    const size_t groups = (tiles + RS_GROUP - 1) / RS_GROUP;
    uint32_t* group_desc;
    carve(p, group_desc, RS_MAX_PASSES * groups * RS_RADIX);
    if (!have_hist) {
        int rc = radix_sort_clear(temp, n, passes, stream);
        if (rc != LG_OK) return rc;
        const int hist_blocks = (int)min((size_t)LG_NUM_SMS * 8, (n + 255) / 256);
        rs_histogram_kernel<K><<<hist_blocks, 256, 0, stream>>>(keys_a, (uint32_t)n, begin_bit, end_bit, passes, hist);
        LG_LAUNCH_CHECK(debug, stream);
    }
    const size_t smem = sizeof(RsSmem<K>);
    // per call: the attribute belongs to the current device's context, and the call costs about a microsecond
    LG_CUDA(cudaFuncSetAttribute(rs_onesweep_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    K* src_k = keys_a;
    K* dst_k = keys_b;
    uint32_t* src_v = vals_a;
    uint32_t* dst_v = vals_b;
    for (int pass = 0; pass < passes; pass++) {
        const int shift = begin_bit + pass * RS_BITS;
        const int nb = min(RS_BITS, end_bit - shift);
        rs_onesweep_kernel<K><<<(unsigned)tiles, RS_THREADS, smem, stream>>>(
            src_k, dst_k, src_v, dst_v, (uint32_t)n, shift, (1u << nb) - 1u, hist + pass * RS_RADIX,
            lookback + (size_t)pass * tiles * RS_RADIX, group_desc + (size_t)pass * groups * RS_RADIX, tickets + pass);
        LG_LAUNCH_CHECK(debug, stream);
        K* tk = src_k; src_k = dst_k; dst_k = tk;
        uint32_t* tv = src_v; src_v = dst_v; dst_v = tv;
    }
    return LG_OK;
}

int radix_sort_pairs_u64(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n,
                         int begin_bit, int end_bit, char* temp, size_t temp_bytes, bool debug, cudaStream_t stream,
                         bool* result_in_b) {
    return radix_sort_pairs_impl<unsigned long long>(reinterpret_cast<unsigned long long*>(keys_a),
                                                     reinterpret_cast<unsigned long long*>(keys_b), vals_a, vals_b, n,
                                                     begin_bit, end_bit, temp, temp_bytes, debug, stream, result_in_b,
                                                     false);
}

int radix_sort_pairs_u32(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n,
                         int begin_bit, int end_bit, char* temp, size_t temp_bytes, bool debug, cudaStream_t stream,
                         bool* result_in_b) {
    return radix_sort_pairs_impl<uint32_t>(keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes,
                                           debug, stream, result_in_b, false);
}

int radix_sort_pairs_u32_prehist(uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, size_t n,
                                 int begin_bit, int end_bit, char* temp, size_t temp_bytes, bool debug,
                                 cudaStream_t stream, bool* result_in_b) {
    return radix_sort_pairs_impl<uint32_t>(keys_a, keys_b, vals_a, vals_b, n, begin_bit, end_bit, temp, temp_bytes,
                                           debug, stream, result_in_b, true);
}

}  // namespace lg
