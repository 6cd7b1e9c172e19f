// adam.cu — fused Adam update of the flat (59 x P) Gaussian-parameter buffer of the view-parallel trainer.
// One pass (read grad, exp_avg, exp_avg_sq, param; write exp_avg, exp_avg_sq, param = 28 B per element) replaces the
// ~10 elementwise launches of torch.optim.Adam's per-group update as the reference runs it (LG/train.py:278-288,
// optimiser groups LG/scene/gaussian_model.py:178-211).  Semantics = torch.optim.Adam (amsgrad off, no weight decay):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with one learning rate per contiguous segment (= parameter group) of the buffer.
#include "common.cuh"

namespace lg {

#define ADAM_MAX_SEGMENTS 16
struct AdamSegments {
    long long end[ADAM_MAX_SEGMENTS];  // exclusive end offset (elements) of every segment, ascending
    float step_size[ADAM_MAX_SEGMENTS];  // lr / (1 - b1^t)
    float step_size_b[ADAM_MAX_SEGMENTS];  // second rate of a split segment (columns >= split of every row)
    int width[ADAM_MAX_SEGMENTS], split[ADAM_MAX_SEGMENTS];  // width <= 1: plain segment
    int count;
};

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                   long long n, AdamSegments seg, float beta1, float beta2, float eps,
                                                   float inv_sqrt_bc2, float grad_scale) {
    // four independent elements per trip (all 16 loads issued before the first use) keep enough bytes in flight
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += 4 * stride) {
        float g[4], m[4], v[4], p[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * stride;
            if (i < n) {
                g[u] = grad[i];
                m[u] = exp_avg[i];
                v[u] = exp_avg_sq[i];
                p[u] = param[i];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = base + u * stride;
            if (i >= n) break;
            int s = 0;
            while (s + 1 < seg.count && i >= seg.end[s]) s++;
            float step = seg.step_size[s];
            if (seg.width[s] > 1) {
                const long long local = i - (s ? seg.end[s - 1] : 0);
                if ((int)(local % seg.width[s]) >= seg.split[s]) step = seg.step_size_b[s];
            }
            const float gs = g[u] * grad_scale;
            const float mn = beta1 * m[u] + (1.0f - beta1) * gs;
            const float vn = beta2 * v[u] + (1.0f - beta2) * gs * gs;
            exp_avg[i] = mn;
            exp_avg_sq[i] = vn;
            const float denom = sqrtf(vn) * inv_sqrt_bc2 + eps;
            param[i] = p[u] - step * (mn / denom);
        }
    }
}

// 16-byte variant: every segment boundary is a multiple of 4 elements (the trainer pads the slab stride to 4 rows), so
// the four elements of a float4 share their segment and, in a row-split segment whose width is a multiple of 4, their
// row.  One segment search and a quarter of the index arithmetic per element; two float4 per array in flight per trip.
__global__ void __launch_bounds__(256) adam_kernel_v4(float4* __restrict__ param, const float4* __restrict__ grad,
                                                      float4* __restrict__ exp_avg, float4* __restrict__ exp_avg_sq,
                                                      long long n4, AdamSegments seg, float beta1, float beta2,
                                                      float eps, float inv_sqrt_bc2, float grad_scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += 2 * stride) {
        float4 g[2], m[2], v[2], p[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long q = base + u * stride;
            if (q < n4) {
                g[u] = __ldcs(grad + q);
                m[u] = exp_avg[q];
                v[u] = exp_avg_sq[q];
                p[u] = param[q];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const long long q = base + u * stride;
            if (q >= n4) break;
            const long long i = 4 * q;
            int s = 0;
            while (s + 1 < seg.count && i >= seg.end[s]) s++;
            float st[4] = {seg.step_size[s], seg.step_size[s], seg.step_size[s], seg.step_size[s]};
            if (seg.width[s] > 1) {
                const int col = (int)((i - (s ? seg.end[s - 1] : 0)) % seg.width[s]);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (col + k >= seg.split[s]) st[k] = seg.step_size_b[s];
            }
            float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w}, mv[4] = {m[u].x, m[u].y, m[u].z, m[u].w};
            float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, pv[4] = {p[u].x, p[u].y, p[u].z, p[u].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float gs = gv[k] * grad_scale;
                mv[k] = beta1 * mv[k] + (1.0f - beta1) * gs;
                vv[k] = beta2 * vv[k] + (1.0f - beta2) * gs * gs;
                pv[k] = pv[k] - st[k] * (mv[k] / (sqrtf(vv[k]) * inv_sqrt_bc2 + eps));
            }
            exp_avg[q] = make_float4(mv[0], mv[1], mv[2], mv[3]);
            exp_avg_sq[q] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            param[q] = make_float4(pv[0], pv[1], pv[2], pv[3]);
        }
    }
}

}  // namespace lg

using namespace lg;

extern "C" int lg_adam_step_split(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                  int num_segments, const long long* segment_ends_host, const float* lrs_host,
                                  const float* lrs_b_host, const int* row_width_host, const int* row_split_host,
                                  float beta1, float beta2, float eps, int step, float grad_scale, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || num_segments < 1 || num_segments > ADAM_MAX_SEGMENTS ||
        !segment_ends_host || !lrs_host || step < 1) {
        set_error("lg_adam_step: invalid arguments (1..%d segments, step >= 1)", ADAM_MAX_SEGMENTS);
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return LG_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    AdamSegments seg;
    seg.count = num_segments;
    long long prev = 0;
    for (int s = 0; s < num_segments; s++) {
        if (segment_ends_host[s] < prev) {
            set_error("lg_adam_step: segment ends must be ascending");
            return LG_ERR_INVALID_ARGUMENT;
        }
        prev = seg.end[s] = segment_ends_host[s];
        seg.step_size[s] = (float)((double)lrs_host[s] / bc1);
        seg.width[s] = row_width_host ? row_width_host[s] : 1;
        seg.split[s] = row_split_host ? row_split_host[s] : 0;
        seg.step_size_b[s] = (float)((double)(lrs_b_host ? lrs_b_host[s] : lrs_host[s]) / bc1);
        if (seg.width[s] > 1 && (!lrs_b_host || seg.split[s] < 0 || seg.split[s] > seg.width[s] ||
                                 (seg.end[s] - (s ? seg.end[s - 1] : 0)) % seg.width[s] != 0)) {
            set_error("lg_adam_step_split: segment %d is not a whole number of rows of %d floats", s, seg.width[s]);
            return LG_ERR_INVALID_ARGUMENT;
        }
    }
    if (prev != n) {
        set_error("lg_adam_step: the last segment must end at n");
        return LG_ERR_INVALID_ARGUMENT;
    }
    bool vec = (n % 4 == 0) && ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15u) == 0);
    for (int s = 0; s < num_segments && vec; s++)
        vec = (seg.end[s] % 4 == 0) && (seg.width[s] <= 1 || seg.width[s] % 4 == 0);
    if (vec) {
        const long long n4 = n / 4;
        const int blocks = (int)((n4 + 255) / 256 < (long long)LG_NUM_SMS * 16 ? (n4 + 255) / 256 : (long long)LG_NUM_SMS * 16);
        adam_kernel_v4<<<blocks, 256, 0, stream>>>((float4*)param, (const float4*)grad, (float4*)exp_avg,
                                                   (float4*)exp_avg_sq, n4, seg, beta1, beta2, eps,
                                                   (float)(1.0 / sqrt(bc2)), grad_scale);
    } else {
        const int blocks = (int)((n + 255) / 256 < (long long)LG_NUM_SMS * 16 ? (n + 255) / 256 : (long long)LG_NUM_SMS * 16);
        adam_kernel<<<blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, seg, beta1, beta2, eps,
                                                (float)(1.0 / sqrt(bc2)), grad_scale);
    }
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}

extern "C" int lg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                            int num_segments, const long long* segment_ends_host, const float* lrs_host, float beta1,
                            float beta2, float eps, int step, float grad_scale, void* stream_v) {
    return lg_adam_step_split(param, grad, exp_avg, exp_avg_sq, n, num_segments, segment_ends_host, lrs_host, nullptr,
                              nullptr, nullptr, beta1, beta2, eps, step, grad_scale, stream_v);
}
