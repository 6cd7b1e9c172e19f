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
    int count;
};

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                   float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                   long long n, AdamSegments seg, float beta1, float beta2, float eps,
                                                   float inv_sqrt_bc2, float grad_scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int s = 0;
        while (s + 1 < seg.count && i >= seg.end[s]) s++;
        const float g = grad[i] * grad_scale;
        const float m = beta1 * exp_avg[i] + (1.0f - beta1) * g;
        const float v = beta2 * exp_avg_sq[i] + (1.0f - beta2) * g * g;
        exp_avg[i] = m;
        exp_avg_sq[i] = v;
        const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
        param[i] = param[i] - seg.step_size[s] * (m / denom);
    }
}

}  // namespace lg

using namespace lg;

extern "C" int lg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                            int num_segments, const long long* segment_ends_host, const float* lrs_host, float beta1,
                            float beta2, float eps, int step, float grad_scale, void* stream_v) {
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
    }
    if (prev != n) {
        set_error("lg_adam_step: the last segment must end at n");
        return LG_ERR_INVALID_ARGUMENT;
    }
    const int blocks = (int)((n + 255) / 256 < (long long)LG_NUM_SMS * 16 ? (n + 255) / 256 : (long long)LG_NUM_SMS * 16);
    adam_kernel<<<blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, seg, beta1, beta2, eps,
                                            (float)(1.0 / sqrt(bc2)), grad_scale);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}
