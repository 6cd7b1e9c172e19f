// activate.cu — the parameter activations of the trainer in one pass over the raw opacity / scaling / rotation slabs.
// Replaces get_opacity / get_scaling / get_rotation (LG/scene/gaussian_model.py:102-130; activations :36-50):
//   opacities = torch.sigmoid(raw)            = 1 / (1 + exp(-x))
//   scales    = torch.exp(raw)
//   rotations = torch.nn.functional.normalize = raw / max(|raw|_2, 1e-12)
// and keeps |raw rotation| for the chain rule that lg_rasterize_backward_raw applies in the preprocess-backward
// kernel, so neither direction launches a PyTorch elementwise op.  get_xyz is the xyz slab itself and get_features
// (cat(f_dc, f_rest), :121-124) is the SH slab itself: the flat buffer stores a Gaussian's 16 x 3 coefficients as
// one row, f_dc first, which is exactly the (P, M, 3) tensor the rasterizer takes.
// HBM-bound: 32 B in + 36 B out per Gaussian.
#include "common.cuh"

namespace lg {

__global__ void __launch_bounds__(256) activate_kernel(int P, const float* __restrict__ opacity_raw,
                                                       const float* __restrict__ scaling_raw,
                                                       const float4* __restrict__ rotation_raw,
                                                       float* __restrict__ opacities, float* __restrict__ scales,
                                                       float4* __restrict__ rotations, float* __restrict__ rot_norm) {
    const int stride = gridDim.x * blockDim.x;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    // scaling: 3P independent floats, coalesced
    for (int e = tid; e < 3 * P; e += stride) scales[e] = expf(scaling_raw[e]);
    for (int i = tid; i < P; i += stride) {
        opacities[i] = 1.0f / (1.0f + expf(-opacity_raw[i]));
        const float4 q = rotation_raw[i];
        const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
        const float d = fmaxf(n, 1e-12f);
        rotations[i] = make_float4(q.x / d, q.y / d, q.z / d, q.w / d);
        rot_norm[i] = n;
    }
}

}  // namespace lg

using namespace lg;

extern "C" int lg_activate_forward(int P, const float* opacity_raw, const float* scaling_raw,
                                   const float* rotation_raw, float* opacities, float* scales, float* rotations,
                                   float* rot_norm, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (P < 0 || (P > 0 && (!opacity_raw || !scaling_raw || !rotation_raw || !opacities || !scales || !rotations ||
                            !rot_norm))) {
        set_error("lg_activate_forward: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (((uintptr_t)rotation_raw | (uintptr_t)rotations) & 15u) {
        set_error("lg_activate_forward: rotation arrays must be 16-byte aligned");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return LG_OK;
    const int blocks = (P + 255) / 256 < LG_NUM_SMS * 8 ? (P + 255) / 256 : LG_NUM_SMS * 8;
    activate_kernel<<<blocks, 256, 0, stream>>>(P, opacity_raw, scaling_raw, (const float4*)rotation_raw, opacities,
                                                scales, (float4*)rotations, rot_norm);
    LG_LAUNCH_CHECK(false, stream);
    return LG_OK;
}
