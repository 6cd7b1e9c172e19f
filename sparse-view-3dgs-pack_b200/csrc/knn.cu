// knn.cu — distCUDA2: mean squared distance to the 3 nearest neighbours of every point.
// Replaces SimpleKNN::knn and its kernels (KNN/simple_knn.cu:46-222): bbox reduction (init {0,0,0}, so the box
// always contains the origin — quirk Q6), 30-bit Morton codes, radix sort, per-1024-point boxes, exact 3-NN by
// box pruning.  Differences: no host round-trips (the bbox stays on the device), no cudaMalloc per call, the
// library's own radix sort, and candidate boxes are staged in shared memory once per block (the 1024 threads of a
// block are Morton-neighbours and want the same boxes) instead of being gathered from global memory per thread.
// The result is the exact 3-NN, so it is independent of traversal order.
#include "common.cuh"
#include <float.h>

namespace lg {

#define KNN_BOX 1024

struct KnnWorkspace {
    float* bbox;            // min xyz, max xyz (6) + pad
    float* partial;         // per-block partial min/max (6 per block)
    uint32_t* morton_a;
    uint32_t* morton_b;
    uint32_t* index_a;
    uint32_t* index_b;
    float* boxes;           // 6 floats per box
    char* sort_temp;
    size_t sort_temp_bytes;
    static KnnWorkspace from_chunk(char*& chunk, size_t P) {
        KnnWorkspace w;
        const size_t nblk = (P + 1023) / 1024, nbox = (P + KNN_BOX - 1) / KNN_BOX;
        carve(chunk, w.bbox, 8);
        carve(chunk, w.partial, 6 * nblk);
        carve(chunk, w.morton_a, P);
        carve(chunk, w.morton_b, P);
        carve(chunk, w.index_a, P);
        carve(chunk, w.index_b, P);
        carve(chunk, w.boxes, 6 * nbox);
        w.sort_temp_bytes = radix_sort_temp_bytes(P, 4);
        carve(chunk, w.sort_temp, w.sort_temp_bytes);
        return w;
    }
};

__device__ __forceinline__ float block_reduce_minmax(float v, bool is_min, float* s_tmp) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_min ? fminf(v, t) : fmaxf(v, t);
    }
    __syncthreads();
    if (lane == 0) s_tmp[warp] = v;
    __syncthreads();
    const unsigned nw = blockDim.x >> 5;
    v = lane < nw ? s_tmp[lane] : (is_min ? FLT_MAX : -FLT_MAX);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_min ? fminf(v, t) : fmaxf(v, t);
    }
    return v;
}

__global__ void __launch_bounds__(1024) knn_bbox_partial_kernel(int P, const float* __restrict__ pts,
                                                                float* __restrict__ partial) {
    __shared__ float s_tmp[32];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float p[3] = {0.f, 0.f, 0.f};  // neutral w.r.t. the reference's {0,0,0} init of both reductions
    if (idx < P) { p[0] = pts[3 * (size_t)idx]; p[1] = pts[3 * (size_t)idx + 1]; p[2] = pts[3 * (size_t)idx + 2]; }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float mn = block_reduce_minmax(p[k], true, s_tmp);
        const float mx = block_reduce_minmax(p[k], false, s_tmp);
        if (threadIdx.x == 0) { partial[6 * blockIdx.x + k] = mn; partial[6 * blockIdx.x + 3 + k] = mx; }
    }
}

__global__ void __launch_bounds__(1024) knn_bbox_final_kernel(int nblk, const float* __restrict__ partial,
                                                              float* __restrict__ bbox) {
    __shared__ float s_tmp[32];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float mn = 0.f, mx = 0.f;  // init {0,0,0} (simple_knn.cu:192)
        for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
            mn = fminf(mn, partial[6 * b + k]);
            mx = fmaxf(mx, partial[6 * b + 3 + k]);
        }
        mn = block_reduce_minmax(mn, true, s_tmp);
        mx = block_reduce_minmax(mx, false, s_tmp);
        if (threadIdx.x == 0) { bbox[k] = mn; bbox[3 + k] = mx; }
    }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(int P, const float* __restrict__ pts,
                                                         const float* __restrict__ bbox, uint32_t* __restrict__ codes,
                                                         uint32_t* __restrict__ index) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float mnx = bbox[0], mny = bbox[1], mnz = bbox[2], mxx = bbox[3], mxy = bbox[4], mxz = bbox[5];
    const float x = pts[3 * (size_t)idx], y = pts[3 * (size_t)idx + 1], z = pts[3 * (size_t)idx + 2];
    // coord2Morton (simple_knn.cu:55-62): float -> uint32 conversion of ((c - min) / (max - min)) * 1023
    const uint32_t mx = prep_morton((uint32_t)(((x - mnx) / (mxx - mnx)) * 1023.0f));
    const uint32_t my = prep_morton((uint32_t)(((y - mny) / (mxy - mny)) * 1023.0f));
    const uint32_t mz = prep_morton((uint32_t)(((z - mnz) / (mxz - mnz)) * 1023.0f));
    codes[idx] = mx | (my << 1) | (mz << 2);
    index[idx] = (uint32_t)idx;
}

// boxMinMax (simple_knn.cu:79-118)
__global__ void __launch_bounds__(KNN_BOX) knn_box_minmax_kernel(uint32_t P, const float* __restrict__ pts,
                                                                 const uint32_t* __restrict__ index,
                                                                 float* __restrict__ boxes) {
    __shared__ float s_tmp[32];
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (idx < P) {
        const size_t i = index[idx];
#pragma unroll
        for (int k = 0; k < 3; k++) mn[k] = mx[k] = pts[3 * i + k];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float a = block_reduce_minmax(mn[k], true, s_tmp);
        const float b = block_reduce_minmax(mx[k], false, s_tmp);
        if (threadIdx.x == 0) { boxes[6 * blockIdx.x + k] = a; boxes[6 * blockIdx.x + 3 + k] = b; }
    }
}

__device__ __forceinline__ void update_kbest3(float px, float py, float pz, float qx, float qy, float qz, float* knn) {
    const float dx = qx - px, dy = qy - py, dz = qz - pz;
    float dist = dx * dx + dy * dy + dz * dz;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (knn[j] > dist) {
            const float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
    }
}

__device__ __forceinline__ float dist_box_point(const float* box, float px, float py, float pz) {
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (px < box[0] || px > box[3]) dx = fminf(fabsf(px - box[0]), fabsf(px - box[3]));
    if (py < box[1] || py > box[4]) dy = fminf(fabsf(py - box[1]), fabsf(py - box[4]));
    if (pz < box[2] || pz > box[5]) dz = fminf(fabsf(pz - box[2]), fabsf(pz - box[5]));
    return dx * dx + dy * dy + dz * dz;
}

// boxMeanDist (simple_knn.cu:148-184), block-cooperative: a candidate box is staged in shared memory when at
// least one thread of the block cannot prune it.
__global__ void __launch_bounds__(KNN_BOX) knn_box_mean_dist_kernel(uint32_t P, const float* __restrict__ pts,
                                                                    const uint32_t* __restrict__ index,
                                                                    const float* __restrict__ boxes,
                                                                    float* __restrict__ dists) {
    __shared__ float s_x[KNN_BOX], s_y[KNN_BOX], s_z[KNN_BOX];
    __shared__ float s_box[6];
    const int idx = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const bool active = (uint32_t)idx < P;
    const int num_boxes = (int)((P + KNN_BOX - 1) / KNN_BOX);
    float px = 0.f, py = 0.f, pz = 0.f;
    uint32_t my_index = 0;
    if (active) {
        my_index = index[idx];
        px = pts[3 * (size_t)my_index]; py = pts[3 * (size_t)my_index + 1]; pz = pts[3 * (size_t)my_index + 2];
    }
    // own box first: doubles as the staging for the +-3 Morton-neighbour seed
    s_x[threadIdx.x] = px; s_y[threadIdx.x] = py; s_z[threadIdx.x] = pz;
    __syncthreads();
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    if (active) {
        const int lo = max(0, idx - 3), hi = min((int)P - 1, idx + 3);
        for (int i = lo; i <= hi; i++) {
            if (i == idx) continue;
            const int local = i - (int)(blockIdx.x * blockDim.x);
            float qx, qy, qz;
            if (local >= 0 && local < KNN_BOX) { qx = s_x[local]; qy = s_y[local]; qz = s_z[local]; }
            else { const size_t q = index[i]; qx = pts[3 * q]; qy = pts[3 * q + 1]; qz = pts[3 * q + 2]; }
            update_kbest3(px, py, pz, qx, qy, qz, best);
        }
    }
    const float reject = best[2];
    best[0] = best[1] = best[2] = FLT_MAX;

    for (int b = 0; b < num_boxes; b++) {
        __syncthreads();  // previous box fully consumed
        if (threadIdx.x < 6) s_box[threadIdx.x] = boxes[6 * b + threadIdx.x];
        __syncthreads();
        bool want = false;
        if (active) {
            const float d = dist_box_point(s_box, px, py, pz);
            want = !(d > reject || d > best[2]);
        }
        if (!__syncthreads_or(want)) continue;
        const uint32_t src = (uint32_t)b * KNN_BOX + threadIdx.x;
        if (src < P) {
            const size_t q = index[src];
            s_x[threadIdx.x] = pts[3 * q]; s_y[threadIdx.x] = pts[3 * q + 1]; s_z[threadIdx.x] = pts[3 * q + 2];
        }
        __syncthreads();
        if (want) {
            const int cnt = (int)min((uint32_t)KNN_BOX, P - (uint32_t)b * KNN_BOX);
            const int self = idx - b * KNN_BOX;
            for (int i = 0; i < cnt; i++) {
                if (i == self) continue;
                update_kbest3(px, py, pz, s_x[i], s_y[i], s_z[i], best);
            }
        }
    }
    if (active) dists[my_index] = (best[0] + best[1] + best[2]) / 3.0f;
}

}  // namespace lg

using namespace lg;

extern "C" size_t lg_knn_workspace_bytes(int P) {
    char* p = nullptr;
    KnnWorkspace::from_chunk(p, (size_t)(P < 0 ? 0 : P));
    return (size_t)p + 128;
}

extern "C" int lg_knn_mean_dist2(int P, const float* points, float* mean_dists, char* workspace,
                                 size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (P < 0 || (P > 0 && (!points || !mean_dists || !workspace))) {
        set_error("lg_knn_mean_dist2: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return LG_OK;
    if (workspace_bytes < lg_knn_workspace_bytes(P)) {
        set_error("lg_knn_mean_dist2: workspace too small (%zu < %zu)", workspace_bytes, lg_knn_workspace_bytes(P));
        return LG_ERR_INVALID_ARGUMENT;
    }
    char* p = workspace;
    KnnWorkspace w = KnnWorkspace::from_chunk(p, (size_t)P);
    const int nblk = (P + 1023) / 1024, nbox = (P + KNN_BOX - 1) / KNN_BOX;
    knn_bbox_partial_kernel<<<nblk, 1024, 0, stream>>>(P, points, w.partial);
    LG_LAUNCH_CHECK(false, stream);
    knn_bbox_final_kernel<<<1, 1024, 0, stream>>>(nblk, w.partial, w.bbox);
    LG_LAUNCH_CHECK(false, stream);
    knn_morton_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, w.bbox, w.morton_a, w.index_a);
    LG_LAUNCH_CHECK(false, stream);
    bool in_b = false;
    int rc = radix_sort_pairs_u32(w.morton_a, w.morton_b, w.index_a, w.index_b, (size_t)P, 0, 32, w.sort_temp,
                                  w.sort_temp_bytes, false, stream, &in_b);
    if (rc != LG_OK) return rc;
    const uint32_t* sorted_index = in_b ? w.index_b : w.index_a;
    knn_box_minmax_kernel<<<nbox, KNN_BOX, 0, stream>>>((uint32_t)P, points, sorted_index, w.boxes);
    LG_LAUNCH_CHECK(false, stream);
    knn_box_mean_dist_kernel<<<nbox, KNN_BOX, 0, stream>>>((uint32_t)P, points, sorted_index, w.boxes, mean_dists);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}
