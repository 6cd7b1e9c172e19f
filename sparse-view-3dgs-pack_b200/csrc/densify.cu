// densify.cu — adaptive density control on the flat, field-major Gaussian buffer of the view-parallel trainer.
//
// Replaces, for the flat (F x P) layout (F = 11 + 3M floats per Gaussian: xyz 3 | SH 3M (f_dc then f_rest) |
// opacity 1 | scaling 3 | rotation 4; every field a contiguous (P, w) slab), the tensor surgery of
//   GaussianModel.densify_and_prune   LG/scene/gaussian_model.py:456-476
//     -> densify_and_clone            :437-454
//     -> densify_and_split            :411-435
//     -> prune_points / _prune_optimizer / cat_tensors_to_optimizer / densification_postfix  :331-409
//   GaussianModel.add_densification_stats :478-480  + the max_radii2D update of LG/train.py:268
//   GaussianModel.reset_opacity       :258-261 (+ replace_tensor_to_optimizer :316-329)
// The reference rebuilds all six parameter tensors and both Adam moments by boolean indexing + torch.cat three times
// per call (clone, split, prune) and a fourth time for the opacity / size pruning.  Here the whole call is one
// decision pass (36 B read per Gaussian), one scan and ONE gather pass that writes every surviving / new row of the
// parameters and both moments exactly once (3 x 4F bytes in, 3 x 4F bytes out per output row).
//
// Output order = the reference's: surviving originals (in order) | surviving clones (in order) | first split
// children of every split parent (in order) | second children.  Why that is what the reference produces: clones are
// appended after the P0 originals; split candidates are looked up in a gradient vector padded with zeros for the
// clones (:414-415), so a clone is never split; 2S children are appended as [all first samples, all second samples]
// (`repeat(N,1)`, :421-430); the S parents are pruned (:434-435); the final opacity / world-size pruning (:464-469)
// is a per-row predicate, so it commutes with the appends.  `max_radii2D > max_screen_size` is evaluated on the
// zeros that densification_postfix just wrote (:404-409) and is therefore never true — kept as the reference has it
// (SURVEY.md App. D).  The unit normals of the split (`torch.normal(0, std)` = randn * std, :421-423) are an INPUT
// (`eps`, row k*S + j for the k-th child of the j-th split parent), drawn by the caller from a generator shared by all
// data-parallel ranks so that every replica performs the identical edit.
#include "common.cuh"

namespace lg {

#define DN_BLOCK 1024
enum { DN_KEEP = 0, DN_CLONE = 1, DN_SPLIT = 2, DN_CHILD = 3 };
// src_index entry: source row | kind << 30 (kind 0 original, 1 clone, 2 first child, 3 second child)

struct DensifyCfg {
    float max_grad, min_opacity, dense_extent, big_extent;  // thresholds already rounded to fp32 on the host
    int use_size;                                             // max_screen_size is not None
};

// per-Gaussian decisions; bit0 keep original, bit1 emit clone, bit2 split parent, bit3 emit children
__device__ __forceinline__ unsigned densify_flags(int i, const float* __restrict__ scaling,
                                                  const float* __restrict__ opacity, const float* __restrict__ accum,
                                                  const float* __restrict__ denom, const DensifyCfg& c) {
    float g = accum[i] / denom[i];  // gaussian_model.py:457
    if (g != g) g = 0.0f;           // :458
    const float s0 = expf(scaling[3 * i + 0]), s1 = expf(scaling[3 * i + 1]), s2 = expf(scaling[3 * i + 2]);
    const float smax = fmaxf(s0, fmaxf(s1, s2));
    const bool hot = fabsf(g) >= c.max_grad;                  // :439 (norm over a length-1 axis) / :416
    const bool clone = hot && smax <= c.dense_extent;         // :440-441
    const bool split = hot && smax > c.dense_extent;          // :417-418
    const float o = 1.0f / (1.0f + expf(-opacity[i]));        // torch.sigmoid
    const bool faint = o < c.min_opacity;                     // :464
    const bool prune_self = faint || (c.use_size && smax > c.big_extent);  // :466-468
    // children carry log(exp(s) / 1.6) (:425; torch's CUDA division by a host scalar multiplies by its reciprocal)
    const float inv = 1.0f / 1.6f;
    const float cmax = fmaxf(expf(logf(s0 * inv)), fmaxf(expf(logf(s1 * inv)), expf(logf(s2 * inv))));
    const bool prune_child = faint || (c.use_size && cmax > c.big_extent);
    unsigned f = 0;
    if (!split && !prune_self) f |= 1u << DN_KEEP;
    if (clone && !prune_self) f |= 1u << DN_CLONE;
    if (split) f |= 1u << DN_SPLIT;
    if (split && !prune_child) f |= 1u << DN_CHILD;
    return f;
}

__global__ void __launch_bounds__(DN_BLOCK) densify_flags_kernel(int P, const float* __restrict__ scaling,
                                                                 const float* __restrict__ opacity,
                                                                 const float* __restrict__ accum,
                                                                 const float* __restrict__ denom, DensifyCfg cfg,
                                                                 uint8_t* __restrict__ flags,
                                                                 uint32_t* __restrict__ block_counts, int nblocks) {
    const int i = blockIdx.x * DN_BLOCK + threadIdx.x;
    unsigned f = 0;
    if (i < P) {
        f = densify_flags(i, scaling, opacity, accum, denom, cfg);
        flags[i] = (uint8_t)f;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int n = __syncthreads_count((f >> k) & 1u);
        if (threadIdx.x == 0) block_counts[k * nblocks + blockIdx.x] = (uint32_t)n;
    }
}

// exclusive scan of the four per-block count rows (one block; nblocks <= a few thousand) + the four totals
__global__ void __launch_bounds__(1024) densify_scan_kernel(uint32_t* __restrict__ block_counts, int nblocks,
                                                            int* __restrict__ totals) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int k = 0; k < 4; k++) {
        uint32_t* row = block_counts + (size_t)k * nblocks;
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < nblocks; base += 1024) {
            const int j = base + (int)threadIdx.x;
            const uint32_t v = j < nblocks ? row[j] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if ((int)lane >= o) x += y;
            }
            if (lane == 31) s_warp[warp] = x;
            __syncthreads();
            if (warp == 0) {
                uint32_t w = s_warp[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                    if ((int)lane >= o) w += y;
                }
                s_warp[lane] = w;  // inclusive over warps
            }
            __syncthreads();
            const uint32_t carry = s_carry;
            const uint32_t incl = x + (warp ? s_warp[warp - 1] : 0u);
            if (j < nblocks) row[j] = carry + incl - v;
            __syncthreads();
            if (threadIdx.x == 1023) s_carry = carry + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) totals[k] = (int)s_carry;
        __syncthreads();
    }
}

// destination rows of every Gaussian's outputs
__global__ void __launch_bounds__(DN_BLOCK) densify_index_kernel(int P, const uint8_t* __restrict__ flags,
                                                                 const uint32_t* __restrict__ block_offsets, int nblocks,
                                                                 const int* __restrict__ totals,
                                                                 uint32_t* __restrict__ src_index,
                                                                 uint32_t* __restrict__ eps_row) {
    __shared__ uint32_t s_warp[4][32];
    const int i = blockIdx.x * DN_BLOCK + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned f = i < P ? flags[i] : 0u;
    uint32_t rank[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned b = __ballot_sync(0xffffffffu, (f >> k) & 1u);
        rank[k] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) s_warp[k][warp] = __popc(b);
    }
    __syncthreads();
    if (warp < 4) {
        uint32_t w = s_warp[warp][lane], x = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        s_warp[warp][lane] = x - w;  // exclusive over warps
    }
    __syncthreads();
    if (i >= P) return;
#pragma unroll
    for (int k = 0; k < 4; k++) rank[k] += s_warp[k][warp] + block_offsets[k * nblocks + blockIdx.x];
    const uint32_t n_keep = (uint32_t)totals[DN_KEEP], n_clone = (uint32_t)totals[DN_CLONE];
    const uint32_t n_split = (uint32_t)totals[DN_SPLIT], n_child = (uint32_t)totals[DN_CHILD];
    if (f & (1u << DN_KEEP)) src_index[rank[DN_KEEP]] = (uint32_t)i;
    if (f & (1u << DN_CLONE)) src_index[n_keep + rank[DN_CLONE]] = (uint32_t)i | (1u << 30);
    if (f & (1u << DN_CHILD)) {
        const uint32_t d0 = n_keep + n_clone + rank[DN_CHILD], d1 = d0 + n_child;
        src_index[d0] = (uint32_t)i | (2u << 30);
        src_index[d1] = (uint32_t)i | (3u << 30);
        eps_row[d0] = rank[DN_SPLIT];            // first sample of the j-th split parent
        eps_row[d1] = n_split + rank[DN_SPLIT];  // second sample (repeat(N,1) order, :421-423)
    }
}

// One field slab: out row d <- in row src(d); moments zeroed for every new row; split children get the sampled
// position and the shrunk log-scale.  One thread per output float: writes coalesced, reads coalesced wherever the
// survivors are contiguous (the source rows of a kind are ascending).
enum { FIELD_XYZ = 0, FIELD_PLAIN = 1, FIELD_SCALING = 2 };
template <int WIDTH_CT, int KIND>
__global__ void __launch_bounds__(256) densify_apply_kernel(
    long long n_out, int width_rt, const uint32_t* __restrict__ src_index, const uint32_t* __restrict__ eps_row,
    const float* __restrict__ eps, const float* __restrict__ in_p, const float* __restrict__ in_m,
    const float* __restrict__ in_v, float* __restrict__ out_p, float* __restrict__ out_m, float* __restrict__ out_v,
    const float* __restrict__ in_scaling, const float* __restrict__ in_rotation) {
    const int width = WIDTH_CT > 0 ? WIDTH_CT : width_rt;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_out; e += stride) {
        const uint32_t d = (uint32_t)(e / width);
        const int c = (int)(e - (long long)d * width);
        const uint32_t si = src_index[d];
        const uint32_t src = si & 0x3fffffffu, kind = si >> 30;
        const size_t se = (size_t)src * width + c;
        float p = in_p[se];
        float m = 0.0f, v = 0.0f;
        if (kind == 0) {
            m = in_m[se];
            v = in_v[se];
        } else if (kind >= 2) {
            if (KIND == FIELD_SCALING) {
                p = logf(expf(p) * (1.0f / 1.6f));  // :425
            } else if (KIND == FIELD_XYZ) {
                // :421-424  samples = eps * exp(scaling); R = build_rotation(rotation) (utils/general_utils.py:78-99)
                const float* ep = eps + (size_t)eps_row[d] * 3;
                const float sx = ep[0] * expf(in_scaling[3 * (size_t)src + 0]);
                const float sy = ep[1] * expf(in_scaling[3 * (size_t)src + 1]);
                const float sz = ep[2] * expf(in_scaling[3 * (size_t)src + 2]);
                const float4 q4 = *reinterpret_cast<const float4*>(in_rotation + 4 * (size_t)src);
                const float norm = sqrtf(q4.x * q4.x + q4.y * q4.y + q4.z * q4.z + q4.w * q4.w);
                const float r = q4.x / norm, x = q4.y / norm, y = q4.z / norm, z = q4.w / norm;
                float r0, r1, r2;
                if (c == 0) {
                    r0 = 1.0f - 2.0f * (y * y + z * z), r1 = 2.0f * (x * y - r * z), r2 = 2.0f * (x * z + r * y);
                } else if (c == 1) {
                    r0 = 2.0f * (x * y + r * z), r1 = 1.0f - 2.0f * (x * x + z * z), r2 = 2.0f * (y * z - r * x);
                } else {
                    r0 = 2.0f * (x * z - r * y), r1 = 2.0f * (y * z + r * x), r2 = 1.0f - 2.0f * (x * x + y * y);
                }
                p = (r0 * sx + r1 * sy + r2 * sz) + p;
            }
        }
        out_p[e] = p;
        out_m[e] = m;
        out_v[e] = v;
    }
}

// add_densification_stats + max_radii2D update for the visible Gaussians of one view
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const float* __restrict__ grad2D,
                                                            const int* __restrict__ radii, float* __restrict__ accum,
                                                            float* __restrict__ denom, float* __restrict__ max_radii2D) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;  // visibility_filter = radii > 0
    const float gx = grad2D[3 * (size_t)i + 0], gy = grad2D[3 * (size_t)i + 1];
    accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.0f;
    max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
}

// reset_opacity: opacity <- inverse_sigmoid(min(sigmoid(opacity), 0.01)); Adam moments of the slab <- 0
__global__ void __launch_bounds__(256) reset_opacity_kernel(int P, float* __restrict__ opacity, float* __restrict__ m,
                                                            float* __restrict__ v) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const float o = fminf(1.0f / (1.0f + expf(-opacity[i])), 0.01f);
    opacity[i] = logf(o / (1.0f - o));  // utils/general_utils.py inverse_sigmoid
    m[i] = 0.0f;
    v[i] = 0.0f;
}

static inline int dn_blocks(int P) { return (P + DN_BLOCK - 1) / DN_BLOCK; }

}  // namespace lg

using namespace lg;

extern "C" size_t lg_densify_scratch_bytes(int P) {
    if (P < 0) return 0;
    const size_t nb = (size_t)dn_blocks(P > 0 ? P : 1);
    return 128 + (((size_t)P + 127) & ~(size_t)127) + 4 * nb * sizeof(uint32_t);
}

extern "C" int lg_densify_plan(int P, const float* scaling, const float* opacity, const float* grad_accum,
                               const float* denom, float max_grad, float min_opacity, float extent,
                               float percent_dense, float max_screen_size, uint32_t* src_index, uint32_t* eps_row,
                               int* totals_dev, void* scratch, size_t scratch_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (P < 0 || !totals_dev) {
        set_error("lg_densify_plan: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) {
        LG_CUDA(cudaMemsetAsync(totals_dev, 0, 4 * sizeof(int), stream));
        return LG_OK;
    }
    if (!scaling || !opacity || !grad_accum || !denom || !src_index || !eps_row || !scratch ||
        scratch_bytes < lg_densify_scratch_bytes(P)) {
        set_error("lg_densify_plan: null pointer or scratch smaller than lg_densify_scratch_bytes(P)");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P >= (1 << 30)) {
        set_error("lg_densify_plan: P must be below 2^30");
        return LG_ERR_UNSUPPORTED;
    }
    const int nb = dn_blocks(P);
    char* chunk = (char*)scratch;
    uint8_t* flags;
    uint32_t* block_counts;
    carve(chunk, flags, (size_t)P);
    carve(chunk, block_counts, (size_t)4 * nb, 16);
    DensifyCfg cfg;
    cfg.max_grad = max_grad;
    cfg.min_opacity = min_opacity;
    cfg.dense_extent = (float)((double)percent_dense * (double)extent);  // self.percent_dense * scene_extent
    cfg.big_extent = (float)(0.1 * (double)extent);                      // 0.1 * extent
    cfg.use_size = max_screen_size >= 0.0f;
    densify_flags_kernel<<<nb, DN_BLOCK, 0, stream>>>(P, scaling, opacity, grad_accum, denom, cfg, flags,
                                                      block_counts, nb);
    LG_LAUNCH_CHECK(false, stream);
    densify_scan_kernel<<<1, 1024, 0, stream>>>(block_counts, nb, totals_dev);
    LG_LAUNCH_CHECK(false, stream);
    densify_index_kernel<<<nb, DN_BLOCK, 0, stream>>>(P, flags, block_counts, nb, totals_dev, src_index, eps_row);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_densify_apply(int P, int P_stride, int P_new, int P_new_stride, int sh_floats, const float* data,
                                const float* exp_avg,
                                const float* exp_avg_sq, float* data_new, float* exp_avg_new, float* exp_avg_sq_new,
                                const uint32_t* src_index, const uint32_t* eps_row, const float* eps, int eps_rows,
                                void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (P < 0 || P_new < 0 || sh_floats < 0 || P_new > 2 * (long long)P || P_stride < P || P_new_stride < P_new) {
        set_error("lg_densify_apply: invalid sizes (P_new <= 2 P, strides >= row counts)");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P_new == 0) return LG_OK;
    if (!data || !exp_avg || !exp_avg_sq || !data_new || !exp_avg_new || !exp_avg_sq_new || !src_index || !eps_row ||
        (eps_rows > 0 && !eps)) {
        set_error("lg_densify_apply: null pointer");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const int widths[5] = {3, sh_floats, 1, 3, 4};
    const size_t sP = (size_t)P_stride, dP = (size_t)P_new_stride;  // slab k starts at (floats before it) * stride
    size_t off = 0;
    const float* in_scaling = data + (size_t)(3 + sh_floats + 1) * sP;
    const float* in_rotation = in_scaling + 3 * sP;
    for (int f = 0; f < 5; f++) {
        const int w = widths[f];
        if (w == 0) continue;
        const long long n_out = (long long)P_new * w;
        const int blocks = (int)((n_out + 255) / 256 < (long long)LG_NUM_SMS * 32 ? (n_out + 255) / 256
                                                                                   : (long long)LG_NUM_SMS * 32);
#define DN_ARGS n_out, w, src_index, eps_row, eps, data + off * sP, exp_avg + off * sP, exp_avg_sq + off * sP, \
                data_new + off * dP, exp_avg_new + off * dP, exp_avg_sq_new + off * dP, in_scaling, in_rotation
        if (f == 0) densify_apply_kernel<3, FIELD_XYZ><<<blocks, 256, 0, stream>>>(DN_ARGS);
        else if (f == 3) densify_apply_kernel<3, FIELD_SCALING><<<blocks, 256, 0, stream>>>(DN_ARGS);
        else if (w == 48) densify_apply_kernel<48, FIELD_PLAIN><<<blocks, 256, 0, stream>>>(DN_ARGS);
        else if (w == 4) densify_apply_kernel<4, FIELD_PLAIN><<<blocks, 256, 0, stream>>>(DN_ARGS);
        else if (w == 1) densify_apply_kernel<1, FIELD_PLAIN><<<blocks, 256, 0, stream>>>(DN_ARGS);
        else densify_apply_kernel<0, FIELD_PLAIN><<<blocks, 256, 0, stream>>>(DN_ARGS);
#undef DN_ARGS
        LG_LAUNCH_CHECK(false, stream);
        off += (size_t)w;
    }
    return LG_OK;
}

extern "C" int lg_densify_stats(int P, const float* grad2D, const int* radii, float* grad_accum, float* denom,
                                float* max_radii2D, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (P < 0 || (P > 0 && (!grad2D || !radii || !grad_accum || !denom || !max_radii2D))) {
        set_error("lg_densify_stats: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return LG_OK;
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, grad2D, radii, grad_accum, denom, max_radii2D);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_reset_opacity(int P, float* opacity, float* exp_avg, float* exp_avg_sq, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (P < 0 || (P > 0 && (!opacity || !exp_avg || !exp_avg_sq))) {
        set_error("lg_reset_opacity: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return LG_OK;
    reset_opacity_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, opacity, exp_avg, exp_avg_sq);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}
