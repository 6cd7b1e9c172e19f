// binning.cu — tile binning: (tile | depth) key emission, stable radix sort, tile-range identification.
// Replaces duplicateWithKeys + cub::DeviceRadixSort::SortPairs + cudaMemset + identifyTileRanges
// (DGR/cuda_rasterizer/rasterizer_impl.cu:70-138, 292-321).
#include "common.cuh"

namespace lg {

// bits needed for the tile id: the reference's getHigherMsb (rasterizer_impl.cu:35-50) == 32 - clz(n) for n >= 1
// (tests/test_host_logic.py checks the equality against a line-by-line restatement).
int higher_msb(uint32_t n) {
    int b = 0;
    while (n) { b++; n >>= 1; }
    return b < 1 ? 1 : b;
}

// One (key, value) per Gaussian/tile overlap, row-major over the Gaussian's tile rectangle, at
// offsets[idx-1] .. offsets[idx]-1 — the order the reference's per-thread double loop produces.  Rectangles of
// more than EMIT_COOP tiles are written cooperatively by the whole warp (coalesced 8-byte stores) instead of by
// one thread.
#define EMIT_COOP 8
__global__ void __launch_bounds__(256) emit_keys_kernel(int P, const float2* __restrict__ xy,
                                                        const float* __restrict__ depths,
                                                        const uint32_t* __restrict__ offsets,
                                                        const int* __restrict__ radii, unsigned long long* __restrict__ keys,
                                                        uint32_t* __restrict__ vals, int grid_x, int grid_y) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t x0 = 0, y0 = 0, x1 = 0, y1 = 0, off = 0, n = 0, dbits = 0;
    if (idx < P) {
        const int r = radii[idx];
        if (r > 0) {
            const float2 p = xy[idx];
            lg_get_rect(p.x, p.y, r, grid_x, grid_y, x0, y0, x1, y1);
            n = (x1 - x0) * (y1 - y0);
            off = idx == 0 ? 0u : offsets[idx - 1];
            dbits = __float_as_uint(depths[idx]);
        }
    }
    if (n > 0 && n <= EMIT_COOP) {
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = x0; x < x1; x++) {
                keys[off] = ((unsigned long long)(y * (uint32_t)grid_x + x) << 32) | dbits;
                vals[off] = (uint32_t)idx;
                off++;
            }
    }
    unsigned big = __ballot_sync(0xffffffffu, n > EMIT_COOP);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const uint32_t bw = __shfl_sync(0xffffffffu, x1, src) - bx0;
        const uint32_t bn = __shfl_sync(0xffffffffu, n, src), boff = __shfl_sync(0xffffffffu, off, src);
        const uint32_t bd = __shfl_sync(0xffffffffu, dbits, src);
        const uint32_t bidx = (uint32_t)(idx - (int)lane + src);
        for (uint32_t i = lane; i < bn; i += 32) {
            const uint32_t y = by0 + i / bw, x = bx0 + i % bw;
            keys[boff + i] = ((unsigned long long)(y * (uint32_t)grid_x + x) << 32) | bd;
            vals[boff + i] = bidx;
        }
    }
}

// identifyTileRanges (rasterizer_impl.cu:116-138); tiles with no entries keep (0,0) from the memset.
__global__ void __launch_bounds__(256) tile_ranges_kernel(uint32_t L, const unsigned long long* __restrict__ keys,
                                                          uint2* __restrict__ ranges) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) ranges[cur].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

int launch_binning(int P, int R, int W, int H, const GeometryState& g, const int* radii, BinningState& b,
                   ImageState& img, bool debug, cudaStream_t stream) {
    const int gx = num_tiles_x(W), gy = num_tiles_y(H);
    const int T = gx * gy;
    LG_CUDA(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)T, stream));
    if (R <= 0) return LG_OK;
    const int end_bit = 32 + higher_msb((uint32_t)T);
    const int passes = radix_sort_num_passes(0, end_bit);
    // ping-pong so that the sorted list ends in point_list_keys / point_list
    uint64_t* ka = (passes & 1) ? b.point_list_keys_unsorted : b.point_list_keys;
    uint64_t* kb = (passes & 1) ? b.point_list_keys : b.point_list_keys_unsorted;
    uint32_t* va = (passes & 1) ? b.point_list_unsorted : b.point_list;
    uint32_t* vb = (passes & 1) ? b.point_list : b.point_list_unsorted;
    emit_keys_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g.means2D, g.depths, g.point_offsets, radii,
                                                          reinterpret_cast<unsigned long long*>(ka), va, gx, gy);
    LG_LAUNCH_CHECK(debug, stream);
    bool in_b = false;
    int rc = radix_sort_pairs_u64(ka, kb, va, vb, (size_t)R, 0, end_bit, b.sort_temp, b.sort_temp_bytes, debug, stream,
                                  &in_b);
    if (rc != LG_OK) return rc;
    tile_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>((uint32_t)R,
                                                            reinterpret_cast<unsigned long long*>(b.point_list_keys),
                                                            img.ranges);
    LG_LAUNCH_CHECK(debug, stream);
    return LG_OK;
}

}  // namespace lg
