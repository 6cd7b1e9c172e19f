// peer.cu — the data-parallel exchange step as ONE kernel over NVLink peer memory: gradient reduce-scatter + Adam +
// parameter all-gather, fused.
//
// The path's only exchange is "sum the 236 B/Gaussian gradient bucket over ranks, then apply the same
// Adam update on every rank".  With NCCL that is an all-reduce (every rank receives the whole summed bucket) followed
// by a full-size Adam pass on every rank (28 B of HBM traffic per element, N times over).  Here every rank owns a
// contiguous 1/N shard of the flat buffer and one kernel per rank
//   1. loads its shard of the gradient bucket from all N ranks' buffers with 16-byte peer loads and sums them in rank
//      order (the reduce-scatter),
//   2. applies the Adam update to its shard — moments are only ever touched by their owner (ZeRO-1 style: the moment
//      traffic and arithmetic drop by N),
//   3. stores the updated parameters into all N ranks' parameter buffers with 16-byte peer stores (the all-gather).
// Per rank that moves (N-1)/N of the bucket in and out over NVLink — what an all-reduce moves — with the optimizer
// pass riding along for free, and every element is computed by exactly one rank, so the replicas stay bit-identical
// by construction.  Ordering across GPUs comes from lg_peer_barrier: a flag exchange with system-scope
// release/acquire before (all buckets complete) and after (all parameter shards delivered) the kernel.
//
// Buffers live in cudaMalloc memory exported with CUDA IPC (lg_peer_alloc/export/open) — one process per GPU, no
// NVSHMEM in this image.  A barrier that does not complete within ~30 s raises LG_ERR_CUDA on the host side at the
// next lg_peer_check instead of hanging the GPU.
#include "common.cuh"

namespace lg {

#define PEER_MAX 8
struct PeerPtrs {
    void* p[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned* addr, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* addr) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

// thread t: tell rank t that `rank` reached `epoch`, then wait until rank t has told us the same.
// flags.p[q] = rank q's flag array (PEER_MAX words); word r of it is written by rank r only.
__global__ void peer_barrier_kernel(PeerPtrs flags, int rank, int world, unsigned epoch, unsigned* status) {
    const int t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    st_release_sys((unsigned*)flags.p[t] + rank, epoch);
    const unsigned* mine = (const unsigned*)flags.p[rank] + t;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
        if (clock64() - t0 > 60000000000ll) {  // ~30 s at 1.9 GHz: a peer died or never arrived
            atomicExch(status, 1u);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

#define PEER_MAX_SEGMENTS 16
struct PeerAdam {
    long long end[PEER_MAX_SEGMENTS];
    float step_a[PEER_MAX_SEGMENTS], step_b[PEER_MAX_SEGMENTS];
    int width[PEER_MAX_SEGMENTS], split[PEER_MAX_SEGMENTS];
    int count;
    float beta1, beta2, eps, inv_sqrt_bc2, grad_scale;
};

// Adam on the four elements of float4 number q (segment boundaries and split-row widths are multiples of 4, so the
// four share their segment and row: one segment search per float4)
__device__ __forceinline__ float4 peer_adam4(float4 g, float4& m, float4& v, float4 p, long long q, const PeerAdam& a) {
    const long long i = 4 * q;
    int s = 0;
    while (s + 1 < a.count && i >= a.end[s]) s++;
    float st[4] = {a.step_a[s], a.step_a[s], a.step_a[s], a.step_a[s]};
    if (a.width[s] > 1) {
        const int col = (int)((i - (s ? a.end[s - 1] : 0)) % a.width[s]);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (col + k >= a.split[s]) st[k] = a.step_b[s];
    }
    float gv[4] = {g.x, g.y, g.z, g.w}, mv[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
    float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float gs = gv[k] * a.grad_scale;
        mv[k] = a.beta1 * mv[k] + (1.0f - a.beta1) * gs;
        vv[k] = a.beta2 * vv[k] + (1.0f - a.beta2) * gs * gs;
        pv[k] = pv[k] - st[k] * (mv[k] / (sqrtf(vv[k]) * a.inv_sqrt_bc2 + a.eps));
    }
    m = make_float4(mv[0], mv[1], mv[2], mv[3]);
    v = make_float4(vv[0], vv[1], vv[2], vv[3]);
    return make_float4(pv[0], pv[1], pv[2], pv[3]);
}

// lo4 .. hi4: this rank's shard in float4 units.  WORLD and UNROLL are template parameters so that the
// WORLD x UNROLL 16-byte peer loads of a trip are issued back to back before the first add: NVLink round trips are
// several microseconds, so the kernel lives on bytes in flight.
template <int WORLD>
struct PeerUnroll {
    static constexpr int value = 1;  // measured at N = 2: 4 trips in flight per thread are slower (0.42 vs 0.36 ms), the
                                     // all-reduce already runs at ~75 % of the NVLink rate in each direction
};

__device__ __forceinline__ float4 ld_stream(const float4* p) { return *p; }   // cache hints measured: no gain
__device__ __forceinline__ void st_stream(float4* p, float4 v) { *p = v; }

template <int WORLD>
__global__ void __launch_bounds__(256, 2) peer_reduce_adam_kernel(PeerPtrs grads, PeerPtrs params, int rank,
                                                               float4* __restrict__ exp_avg,
                                                               float4* __restrict__ exp_avg_sq, long long lo4,
                                                               long long hi4, PeerAdam a,
                                                               const unsigned* __restrict__ status) {
    if (*status) return;  // a barrier in front of this kernel timed out
    constexpr int U = PeerUnroll<WORLD>::value;
    const float4* src[WORLD];
    float4* dst[WORLD];
#pragma unroll
    for (int r = 0; r < WORLD; r++) {
        src[r] = (const float4*)grads.p[r];                 // summed in rank order on every rank
        dst[r] = (float4*)params.p[(rank + r) % WORLD];     // stores start at home, then walk the ring
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q0 = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < hi4; q0 += U * stride) {
        float4 g[U][WORLD], m[U], v[U], p[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long q = q0 + u * stride;
            if (q < hi4) {
#pragma unroll
                for (int r = 0; r < WORLD; r++) g[u][r] = ld_stream(src[r] + q);
                m[u] = exp_avg[q];
                v[u] = exp_avg_sq[q];
                p[u] = dst[0][q];
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long q = q0 + u * stride;
            if (q >= hi4) break;
            float4 s = g[u][0];
#pragma unroll
            for (int r = 1; r < WORLD; r++) { s.x += g[u][r].x; s.y += g[u][r].y; s.z += g[u][r].z; s.w += g[u][r].w; }
            const float4 pn = peer_adam4(s, m[u], v[u], p[u], q, a);
            exp_avg[q] = m[u];
            exp_avg_sq[q] = v[u];
#pragma unroll
            for (int r = 0; r < WORLD; r++) st_stream(dst[r] + q, pn);
        }
    }
}

// reduce-only variant (no optimizer): every rank ends up with the summed bucket, like an all-reduce
template <int WORLD>
__global__ void __launch_bounds__(256, 2) peer_allreduce_kernel(PeerPtrs grads, int rank, long long lo4, long long hi4,
                                                             float grad_scale, const unsigned* __restrict__ status) {
    if (*status) return;
    constexpr int U = PeerUnroll<WORLD>::value;
    const float4* src[WORLD];
    float4* dst[WORLD];
#pragma unroll
    for (int r = 0; r < WORLD; r++) {
        src[r] = (const float4*)grads.p[r];
        dst[r] = (float4*)grads.p[(rank + r) % WORLD];
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q0 = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < hi4; q0 += U * stride) {
        float4 g[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long q = q0 + u * stride;
            if (q < hi4) {
#pragma unroll
                for (int r = 0; r < WORLD; r++) g[u][r] = ld_stream(src[r] + q);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long q = q0 + u * stride;
            if (q >= hi4) break;
            float4 s = g[u][0];
#pragma unroll
            for (int r = 1; r < WORLD; r++) { s.x += g[u][r].x; s.y += g[u][r].y; s.z += g[u][r].z; s.w += g[u][r].w; }
            s.x *= grad_scale; s.y *= grad_scale; s.z *= grad_scale; s.w *= grad_scale;
#pragma unroll
            for (int r = 0; r < WORLD; r++) st_stream(dst[r] + q, s);
        }
    }
}

// ---- NVLS variants: the buffers are additionally mapped through a multicast object (NVSwitch), so ONE
// multimem.ld_reduce returns the sum over all ranks (added inside the switch) and ONE multimem.st delivers the result
// to all ranks.  Per rank and direction the links then carry about one bucket per step instead of 2 (N-1)/N buckets.
__device__ __forceinline__ float4 mc_ld_reduce_add(const float4* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#ifndef MC_UNROLL
#define MC_UNROLL 4  // multimem.ld_reduce round trips through the switch: keep several in flight per thread
#endif
__global__ void __launch_bounds__(256, 2) peer_reduce_adam_mc_kernel(const float4* __restrict__ mc_grad,
                                                                     float4* __restrict__ mc_param,
                                                                     const float4* __restrict__ local_param,
                                                                     float4* __restrict__ exp_avg,
                                                                     float4* __restrict__ exp_avg_sq, long long lo4,
                                                                     long long hi4, PeerAdam a,
                                                                     const unsigned* __restrict__ status) {
    if (*status) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q0 = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < hi4; q0 += MC_UNROLL * stride) {
        float4 s[MC_UNROLL], m[MC_UNROLL], v[MC_UNROLL], p[MC_UNROLL];
#pragma unroll
        for (int u = 0; u < MC_UNROLL; u++) {
            const long long q = q0 + u * stride;
            if (q < hi4) {
                s[u] = mc_ld_reduce_add(mc_grad + q);
                m[u] = exp_avg[q];
                v[u] = exp_avg_sq[q];
                p[u] = local_param[q];
            }
        }
#pragma unroll
        for (int u = 0; u < MC_UNROLL; u++) {
            const long long q = q0 + u * stride;
            if (q >= hi4) break;
            const float4 pn = peer_adam4(s[u], m[u], v[u], p[u], q, a);
            exp_avg[q] = m[u];
            exp_avg_sq[q] = v[u];
            mc_st(mc_param + q, pn);
        }
    }
}

__global__ void __launch_bounds__(256, 2) peer_allreduce_mc_kernel(float4* __restrict__ mc_grad, long long lo4,
                                                                   long long hi4, float grad_scale,
                                                                   const unsigned* __restrict__ status) {
    if (*status) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q0 = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < hi4; q0 += MC_UNROLL * stride) {
        float4 s[MC_UNROLL];
#pragma unroll
        for (int u = 0; u < MC_UNROLL; u++)
            if (q0 + u * stride < hi4) s[u] = mc_ld_reduce_add(mc_grad + q0 + u * stride);
#pragma unroll
        for (int u = 0; u < MC_UNROLL; u++) {
            const long long q = q0 + u * stride;
            if (q >= hi4) break;
            s[u].x *= grad_scale; s[u].y *= grad_scale; s[u].z *= grad_scale; s[u].w *= grad_scale;
            mc_st(mc_grad + q, s[u]);
        }
    }
}

// One word per device, set by a timed-out barrier.  The exchange kernels queued behind that barrier read it and do
// nothing: sums over peers that never arrived must not reach the parameters (the error surfaces at lg_peer_check).
#define PEER_MAX_DEVICES 32
static unsigned* g_peer_status[PEER_MAX_DEVICES] = {};

static int peer_status_word(unsigned** out) {
    int dev = 0;
    LG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= PEER_MAX_DEVICES) {
        set_error("peer exchange: device ordinal %d out of range", dev);
        return LG_ERR_UNSUPPORTED;
    }
    if (!g_peer_status[dev]) {
        LG_CUDA(cudaMalloc(&g_peer_status[dev], sizeof(unsigned)));
        LG_CUDA(cudaMemset(g_peer_status[dev], 0, sizeof(unsigned)));
    }
    *out = g_peer_status[dev];
    return LG_OK;
}

static void shard_of(long long n4, int rank, int world, long long* lo4, long long* hi4) {
    *lo4 = n4 * rank / world;
    *hi4 = n4 * (rank + 1) / world;
}

}  // namespace lg

using namespace lg;

static int peer_fill_adam(PeerAdam& a, long long n, int num_segments, const long long* segment_ends, const float* lrs,
                          const float* lrs_b, const int* row_width, const int* row_split, float beta1, float beta2,
                          float eps, int step, float grad_scale);

extern "C" int lg_peer_alloc(size_t bytes, void** dev_ptr) {
    if (!dev_ptr || bytes == 0) {
        set_error("lg_peer_alloc: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    LG_CUDA(cudaMalloc(dev_ptr, bytes));
    LG_CUDA(cudaMemset(*dev_ptr, 0, bytes));
    LG_CUDA(cudaDeviceSynchronize());  // the zeros must be in place before a peer can see the handle
    return LG_OK;
}

extern "C" int lg_peer_free(void* dev_ptr) {
    if (dev_ptr) LG_CUDA(cudaFree(dev_ptr));
    return LG_OK;
}

extern "C" int lg_peer_export(void* dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64) {
        set_error("lg_peer_export: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    LG_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, 64);
    return LG_OK;
}

extern "C" int lg_peer_open(const unsigned char* handle64, void** mapped) {
    if (!handle64 || !mapped) {
        set_error("lg_peer_open: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    LG_CUDA(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
    return LG_OK;
}

extern "C" int lg_peer_close(void* mapped) {
    if (mapped) LG_CUDA(cudaIpcCloseMemHandle(mapped));
    return LG_OK;
}

extern "C" int lg_peer_barrier(int rank, int world, void* const* flag_ptrs, unsigned epoch, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !flag_ptrs) {
        set_error("lg_peer_barrier: invalid arguments (1 <= world <= %d)", PEER_MAX);
        return LG_ERR_INVALID_ARGUMENT;
    }
    PeerPtrs f;
    for (int r = 0; r < world; r++) f.p[r] = flag_ptrs[r];
    unsigned* status = nullptr;
    int rc = peer_status_word(&status);
    if (rc != LG_OK) return rc;
    peer_barrier_kernel<<<1, 32, 0, stream>>>(f, rank, world, epoch, status);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_peer_check(void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int dev = 0;
    LG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= PEER_MAX_DEVICES || !g_peer_status[dev]) return LG_OK;
    unsigned s = 0;
    LG_CUDA(cudaMemcpyAsync(&s, g_peer_status[dev], sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    LG_CUDA(cudaStreamSynchronize(stream));
    if (s) {
        LG_CUDA(cudaMemsetAsync(g_peer_status[dev], 0, sizeof(unsigned), stream));  // reported once
        set_error("lg_peer_barrier timed out: a peer rank did not reach the exchange step");
        return LG_ERR_CUDA;
    }
    return LG_OK;
}

static int peer_grid(long long n4) {
    static int per_sm = -1;
    if (per_sm < 0) {
        const char* e = getenv("LGDWT_PEER_BLOCKS_PER_SM");
        per_sm = e ? atoi(e) : 16;  // measured at N = 2: 4 / 8 / 16 blocks per SM -> 0.389 / 0.380 / 0.370 ms
        if (per_sm < 1) per_sm = 1;
    }
    const long long want = (n4 + 255) / 256, cap = (long long)LG_NUM_SMS * per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

extern "C" int lg_peer_reduce_adam(int rank, int world, void* const* grad_ptrs, void* const* param_ptrs,
                                   float* exp_avg, float* exp_avg_sq, long long n, int num_segments,
                                   const long long* segment_ends, const float* lrs, const float* lrs_b,
                                   const int* row_width, const int* row_split, float beta1, float beta2, float eps,
                                   int step, float grad_scale, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !grad_ptrs || !param_ptrs || !exp_avg ||
        !exp_avg_sq || n < 0 || (n & 3) || num_segments < 1 || num_segments > PEER_MAX_SEGMENTS || !segment_ends ||
        !lrs || step < 1) {
        set_error("lg_peer_reduce_adam: invalid arguments (n must be a multiple of 4, world <= %d)", PEER_MAX);
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return LG_OK;
    PeerAdam a;
    {
        int rc = peer_fill_adam(a, n, num_segments, segment_ends, lrs, lrs_b, row_width, row_split, beta1, beta2, eps,
                                step, grad_scale);
        if (rc != LG_OK) return rc;
    }
    PeerPtrs g, p;
    for (int r = 0; r < world; r++) { g.p[r] = grad_ptrs[r]; p.p[r] = param_ptrs[r]; }
    long long lo4, hi4;
    shard_of(n / 4, rank, world, &lo4, &hi4);
    const int blocks = peer_grid(hi4 - lo4);
    unsigned* status = nullptr;
    const int rc_status = peer_status_word(&status);
    if (rc_status != LG_OK) return rc_status;
#define PEER_LAUNCH(W) case W: peer_reduce_adam_kernel<W><<<blocks, 256, 0, stream>>>(g, p, rank, (float4*)exp_avg, \
                                                                                     (float4*)exp_avg_sq, lo4, hi4, a, status); break
    switch (world) {
        PEER_LAUNCH(1); PEER_LAUNCH(2); PEER_LAUNCH(3); PEER_LAUNCH(4); PEER_LAUNCH(5); PEER_LAUNCH(6); PEER_LAUNCH(7);
        PEER_LAUNCH(8);
        default: set_error("lg_peer_reduce_adam: world sizes 1..8 are built (one NVSwitch node)"); return LG_ERR_UNSUPPORTED;
    }
#undef PEER_LAUNCH
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_peer_allreduce(int rank, int world, void* const* grad_ptrs, long long n, float grad_scale,
                                 void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (world < 1 || world > PEER_MAX || rank < 0 || rank >= world || !grad_ptrs || n < 0 || (n & 3)) {
        set_error("lg_peer_allreduce: invalid arguments (n must be a multiple of 4, world <= %d)", PEER_MAX);
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return LG_OK;
    PeerPtrs g;
    for (int r = 0; r < world; r++) g.p[r] = grad_ptrs[r];
    long long lo4, hi4;
    shard_of(n / 4, rank, world, &lo4, &hi4);
    const int blocks = peer_grid(hi4 - lo4);
    unsigned* status = nullptr;
    const int rc_status = peer_status_word(&status);
    if (rc_status != LG_OK) return rc_status;
#define PEER_LAUNCH(W) case W: peer_allreduce_kernel<W><<<blocks, 256, 0, stream>>>(g, rank, lo4, hi4, grad_scale, status); break
    switch (world) {
        PEER_LAUNCH(1); PEER_LAUNCH(2); PEER_LAUNCH(3); PEER_LAUNCH(4); PEER_LAUNCH(5); PEER_LAUNCH(6); PEER_LAUNCH(7);
        PEER_LAUNCH(8);
        default: set_error("lg_peer_allreduce: world sizes 1..8 are built (one NVSwitch node)"); return LG_ERR_UNSUPPORTED;
    }
#undef PEER_LAUNCH
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

static int peer_fill_adam(PeerAdam& a, long long n, int num_segments, const long long* segment_ends, const float* lrs,
                          const float* lrs_b, const int* row_width, const int* row_split, float beta1, float beta2,
                          float eps, int step, float grad_scale) {
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    a.count = num_segments;
    for (int s = 0; s < num_segments; s++) {
        a.end[s] = segment_ends[s];
        a.step_a[s] = (float)((double)lrs[s] / bc1);
        a.step_b[s] = (float)((double)(lrs_b ? lrs_b[s] : lrs[s]) / bc1);
        a.width[s] = row_width ? row_width[s] : 1;
        a.split[s] = row_split ? row_split[s] : 0;
    }
    if (a.end[num_segments - 1] != n) {
        set_error("peer Adam: the last segment must end at n");
        return LG_ERR_INVALID_ARGUMENT;
    }
    for (int s = 0; s < num_segments; s++)
        if (a.end[s] % 4 != 0 || (a.width[s] > 1 && a.width[s] % 4 != 0)) {
            set_error("peer Adam: segment ends and split-row widths must be multiples of 4 elements");
            return LG_ERR_INVALID_ARGUMENT;
        }
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2)); a.grad_scale = grad_scale;
    return LG_OK;
}

extern "C" int lg_peer_reduce_adam_mc(int rank, int world, const void* mc_grad, void* mc_param, const float* local_param,
                                      float* exp_avg, float* exp_avg_sq, long long n, int num_segments,
                                      const long long* segment_ends, const float* lrs, const float* lrs_b,
                                      const int* row_width, const int* row_split, float beta1, float beta2, float eps,
                                      int step, float grad_scale, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (world < 1 || rank < 0 || rank >= world || !mc_grad || !mc_param || !local_param || !exp_avg || !exp_avg_sq ||
        n < 0 || (n & 3) || num_segments < 1 || num_segments > PEER_MAX_SEGMENTS || !segment_ends || !lrs || step < 1) {
        set_error("lg_peer_reduce_adam_mc: invalid arguments (n must be a multiple of 4)");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return LG_OK;
    PeerAdam a;
    int rc = peer_fill_adam(a, n, num_segments, segment_ends, lrs, lrs_b, row_width, row_split, beta1, beta2, eps, step,
                            grad_scale);
    if (rc != LG_OK) return rc;
    long long lo4, hi4;
    shard_of(n / 4, rank, world, &lo4, &hi4);
    unsigned* status = nullptr;
    rc = peer_status_word(&status);
    if (rc != LG_OK) return rc;
    peer_reduce_adam_mc_kernel<<<peer_grid(hi4 - lo4), 256, 0, stream>>>((const float4*)mc_grad, (float4*)mc_param,
                                                                        (const float4*)local_param, (float4*)exp_avg,
                                                                        (float4*)exp_avg_sq, lo4, hi4, a, status);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_peer_allreduce_mc(int rank, int world, void* mc_grad, long long n, float grad_scale, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (world < 1 || rank < 0 || rank >= world || !mc_grad || n < 0 || (n & 3)) {
        set_error("lg_peer_allreduce_mc: invalid arguments (n must be a multiple of 4)");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return LG_OK;
    long long lo4, hi4;
    shard_of(n / 4, rank, world, &lo4, &hi4);
    unsigned* status = nullptr;
    const int rc_status = peer_status_word(&status);
    if (rc_status != LG_OK) return rc_status;
    peer_allreduce_mc_kernel<<<peer_grid(hi4 - lo4), 256, 0, stream>>>((float4*)mc_grad, lo4, hi4, grad_scale, status);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}
