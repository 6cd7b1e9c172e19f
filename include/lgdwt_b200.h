/*
 * lgdwt_b200.h — C ABI of the B200-native LGDWT-GS hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types. Every entry point names the
 * reference interface it replaces (paths relative to /root/reference, prefixes as in SURVEY.md:
 *   DGR/ = fs3dgs_benchmark/gaussian-splatting/submodules/diff-gaussian-rasterization/
 *   KNN/ = fs3dgs_benchmark/gaussian-splatting/submodules/simple-knn/
 *   LG/  = fs3dgs_benchmark/LGDWT-GS/ ).
 *
 * Conventions
 *   - all pointers except `*_ctx`, `num_rendered` and `out_host_*` are DEVICE pointers on the current device;
 *   - tensors are dense, row-major fp32 unless stated; indices / counters are 32-bit;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, what the reference uses);
 *   - every function returns LG_OK (0) or an error code; lg_last_error() returns the message of the calling
 *     thread's last failure.  There is no CPU fallback anywhere behind this ABI.
 */
#ifndef LGDWT_B200_H_INCLUDED
#define LGDWT_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LG_OK 0
#define LG_ERR_INVALID_ARGUMENT 1
#define LG_ERR_CUDA 2
#define LG_ERR_ALLOC 3
#define LG_ERR_UNSUPPORTED 4

#define LG_ABI_VERSION 1

/* Caller-owned resizable buffer: replaces std::function<char*(size_t)> (DGR/cuda_rasterizer/rasterizer.h:33-35,
 * DGR/rasterize_points.cu:27-33).  Must return a device pointer to >= bytes bytes (128-byte aligned), or NULL. */
typedef char* (*lg_alloc_fn)(void* ctx, size_t bytes);

int lg_abi_version(void);
const char* lg_last_error(void);

/* Sizes of the three opaque state buffers (layout private to the library).  Replaces
 * CudaRasterizer::required<GeometryState|ImageState|BinningState> (DGR/cuda_rasterizer/rasterizer_impl.h:64-71). */
size_t lg_geometry_state_bytes(int P, int channels);
size_t lg_image_state_bytes(int width, int height);
size_t lg_binning_state_bytes(int num_rendered, int width, int height);

/* Replaces CudaRasterizer::Rasterizer::forward (DGR/cuda_rasterizer/rasterizer.h:31-59,
 * rasterizer_impl.cu:198-341) as bound by RasterizeGaussiansCUDA (DGR/rasterize_points.cu:35-124).
 *   channels: 3 = the reference NUM_CHANNELS (DGR/cuda_rasterizer/config.h:15); 1..4 supported here.
 *             SH evaluation needs channels == 3; other counts need colors_precomp (rasterizer_impl.cu:244-247).
 *   shs (P,M,3) or NULL; colors_precomp (P,channels) or NULL; scales (P,3)+rotations (P,4) or cov3D_precomp (P,6).
 *   out_color (channels,H,W); out_invdepth (1,H,W) or NULL; radii (P) int32 or NULL.
 *   *num_rendered receives the number of (Gaussian, tile) instances (the reference's return value).  */
int lg_rasterize_forward(
    lg_alloc_fn geometry_alloc, void* geometry_ctx,
    lg_alloc_fn binning_alloc, void* binning_ctx,
    lg_alloc_fn image_alloc, void* image_ctx,
    int P, int D, int M, int channels,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* cam_pos,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    float* out_color,
    float* out_invdepth,
    int antialiasing,
    int* radii,
    int debug,
    void* stream,
    int* num_rendered);

/* The same call without the mid-forward host stall (extension; the reference blocks on a cudaMemcpy of num_rendered
 * before it can size the binning buffer, rasterizer_impl.cu:283-288).  capacity_hint > 0: the binning buffer is
 * requested for that many entries and the whole forward is queued at once; the host then waits only for the entry count
 * (an event behind the scan), and re-queues binning + blending at the exact size if the hint was too small.
 * capacity_hint <= 0: identical to lg_rasterize_forward.  *binning_capacity receives the entry count the binning state
 * was laid out for (>= *num_rendered): pass THAT as `R` to lg_rasterize_backward* / lg_state_read / lg_blend_work_count.
 * A good hint is 1.25 x the num_rendered of the previous call with the same scene and image size. */
int lg_rasterize_forward_hinted(
    lg_alloc_fn geometry_alloc, void* geometry_ctx,
    lg_alloc_fn binning_alloc, void* binning_ctx,
    lg_alloc_fn image_alloc, void* image_ctx,
    int P, int D, int M, int channels,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* cam_pos,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    float* out_color,
    float* out_invdepth,
    int antialiasing,
    int* radii,
    int debug,
    void* stream,
    int capacity_hint,
    int* num_rendered,
    int* binning_capacity);

/* Replaces CudaRasterizer::Rasterizer::backward (DGR/cuda_rasterizer/rasterizer.h:61-97,
 * rasterizer_impl.cu:345-450) as bound by RasterizeGaussiansBackwardCUDA (DGR/rasterize_points.cu:126-223).
 * All dL_* outputs are fully written by the call (zero for invisible Gaussians); the caller does NOT need to
 * zero them first (the reference requires torch::zeros, rasterize_points.cu:163-172).
 *   dL_dpix (channels,H,W); dL_dinvdepth_pix (1,H,W) or NULL.
 *   dL_dmean2D (P,3) [xy in NDC-scaled units, z = 0]; dL_dconic (P,4) [a,b,unused,c as in the reference];
 *   dL_dopacity (P); dL_dcolor (P,channels) — may be NULL unless colors_precomp is given; dL_dinvdepth (P) or NULL;
 *   dL_dmean3D (P,3); dL_dcov3D (P,6) — may be NULL unless cov3D_precomp is given (intermediate results nobody reads
 *   are then not written: 36 B/Gaussian at 3 channels); dL_dsh (P,M,3) or NULL; dL_dscale (P,3); dL_drot (P,4).  */
int lg_rasterize_backward(
    int P, int D, int M, int R, int channels,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* campos,
    float tan_fovx, float tan_fovy,
    const int* radii,
    char* geometry_state,
    char* binning_state,
    char* image_state,
    const float* dL_dpix,
    const float* dL_dinvdepth_pix,
    float* dL_dmean2D,
    float* dL_dconic,
    float* dL_dopacity,
    float* dL_dcolor,
    float* dL_dinvdepth,
    float* dL_dmean3D,
    float* dL_dcov3D,
    float* dL_dsh,
    float* dL_dscale,
    float* dL_drot,
    int antialiasing,
    int debug,
    void* stream);

/* Same call with gradient accumulation (no counterpart in the reference, which always overwrites freshly zeroed
 * tensors and leaves the summation over views to autograd): when `accumulate` is non-zero the PARAMETER gradients
 * dL_dmean3D, dL_dsh, dL_dopacity, dL_dscale and dL_drot are added to the values already in the buffers instead of
 * overwriting them, so a view batch sums its gradients in one flat bucket (the buffer a data-parallel step
 * all-reduces) without a separate read-modify-write pass per view.  The screen-space / per-stage outputs
 * (dL_dmean2D, dL_dconic, dL_dcolor, dL_dinvdepth, dL_dcov3D) are always overwritten. */
int lg_rasterize_backward_ex(
    int P, int D, int M, int R, int channels, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
    const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
    const int* radii, char* geometry_state, char* binning_state, char* image_state, const float* dL_dpix,
    const float* dL_dinvdepth_pix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
    int antialiasing, int debug, void* stream, int accumulate);

/* Same call for a trainer that keeps RAW (pre-activation) parameters (no counterpart in the reference, which leaves
 * the activation chain rule to autograd: sigmoid / exp / normalize of LG/scene/gaussian_model.py:36-50,102-130 and
 * their backward are ~10 P-sized PyTorch launches per view).  With raw_rot_norm != NULL (P floats = |raw rotation|,
 * written by lg_activate_forward) `opacities`, `scales`, `rotations` must be the activated values of the raw
 * parameters, and dL_dopacity / dL_dscale / dL_drot receive the gradients w.r.t. the RAW opacity / scaling /
 * rotation: d/draw_opacity = d/do * o(1-o); d/draw_scaling = d/ds * s; d/draw_rot = (g - q (q.g)) / max(|raw|,1e-12).
 * dL_dmean3D and dL_dsh are unchanged (those parameters have no activation).  NULL = lg_rasterize_backward_ex. */
int lg_rasterize_backward_raw(
    int P, int D, int M, int R, int channels, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
    const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
    const int* radii, char* geometry_state, char* binning_state, char* image_state, const float* dL_dpix,
    const float* dL_dinvdepth_pix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
    int antialiasing, int debug, void* stream, int accumulate, const float* raw_rot_norm);

/* Activations of the raw parameters in one pass (get_opacity / get_scaling / get_rotation,
 * LG/scene/gaussian_model.py:102-130): opacities = sigmoid(raw) = 1/(1+exp(-x)); scales = exp(raw);
 * rotations = raw / max(|raw|_2, 1e-12) (torch.nn.functional.normalize); rot_norm = |raw|_2.  32 B in, 36 B out per
 * Gaussian.  xyz and the SH coefficients have no activation and are used in place. */
int lg_activate_forward(int P, const float* opacity_raw, const float* scaling_raw, const float* rotation_raw,
                        float* opacities, float* scales, float* rotations, float* rot_norm, void* stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible (DGR/cuda_rasterizer/rasterizer.h:24-29,
 * rasterizer_impl.cu:54-66,141-153; torch glue DGR/rasterize_points.cu:225-244).  present: (P) uint8/bool. */
int lg_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                    uint8_t* present, void* stream);

/* Test/inspection access to the opaque state (what a `ref_probe` reads out of the reference buffers through
 * GeometryState/ImageState/BinningState::fromChunk, DGR/cuda_rasterizer/rasterizer_impl.cu:155-194).
 * Each call copies one named array into a caller-provided DEVICE buffer of the documented size:
 *   "depths" f32[P]  "means2D" f32[2P]  "cov3D" f32[6P]  "conic_opacity" f32[4P]  "rgb" f32[channels*P]
 *   "clamped" u8[3P]  "tiles_touched" u32[P]  "point_offsets" u32[P]
 *   "final_T" f32[W*H]  "n_contrib" u32[W*H]  "ranges" u32[2*T]
 *   "point_list" u32[R]  "point_list_keys" u64[R]                                                       */
int lg_state_read(const char* name, int P, int channels, int width, int height, int R,
                  const char* geometry_state, const char* binning_state, const char* image_state,
                  void* dst, size_t dst_bytes, void* stream);

/* Instrumentation (no counterpart in the reference, whose only timing is two torch.cuda.Events around an
 * iteration, LG/train.py:65-66,97,220): number of kernels this library has launched in this process, and optional
 * per-stage CUDA-event timing on the launching stream.  lg_stage_timing_enable(slots) arms a ring of `slots`
 * (<= 256, 0 disables) event sets; every lg_rasterize_forward call advances to the next slot and the following
 * lg_rasterize_backward records into the same one, so a run can be timed step by step without a host sync in the
 * loop.  Stage order of lg_stage_timing_read (milliseconds, -1 when the stage did not run in that slot):
 * preprocess, binning, blend_forward, blend_backward, preprocess_backward. */
/* counts_dev (device, 2 x u64): [0] tile-list entries evaluated over all pixels before early termination,
 * [1] (pixel, Gaussian) pairs that pass all blend tests — the work terms of the blend roofline. */
int lg_blend_work_count(int P, int channels, int width, int height, int R, const char* geometry_state,
                        const char* binning_state, const char* image_state, unsigned long long* counts_dev,
                        void* stream);
#define LG_NUM_STAGES 5
unsigned long long lg_launch_count(void);
/* On-box SIMT peaks the blend kernels are measured against (no counterpart in the reference; SURVEY.md §8d asks for
 * measured FFMA / MUFU / shared-memory ceilings).  Runs six short microbenchmarks on `stream`, best of 5 each:
 * peaks_out[0] FFMA TFLOP/s, [1] MUFU.EX2 G lane-ops/s, [2] MUFU.RCP G lane-ops/s, [3] broadcast LDS.128
 * G warp-instructions/s, [4] SHFL.BFLY G warp-instructions/s, [5] integer ALU (add / logic) G warp-instructions/s. */
int lg_simt_peaks(float* peaks_out, int n, void* stream);
int lg_stage_timing_enable(int slots);
/* time only one rasterizer call in `every` (default 1): an event record between two kernels costs ~3 us of stream
 * serialisation, ten per call; sampling keeps the measurement inside a timed region without paying that on every call */
int lg_stage_timing_sample(int every);
int lg_stage_timing_read(int slot, float* ms_out, int n);

/* Replaces SimpleKNN::knn (KNN/simple_knn.cu:186-222) as bound by distCUDA2 (KNN/spatial.cu:16-26):
 * mean squared distance to the 3 nearest neighbours.  points (P,3); mean_dists (P).
 * workspace: device scratch of >= lg_knn_workspace_bytes(P) bytes (the reference cudaMallocs internally). */
size_t lg_knn_workspace_bytes(int P);
int lg_knn_mean_dist2(int P, const float* points, float* mean_dists, char* workspace, size_t workspace_bytes,
                      void* stream);

/* Fused Haar-DWT loss.  Replaces the PyTorch op chain LG/train.py:131-180 built from
 * get_dwt_subbands (LG/utils/loss_utils.py:106-153), l1_loss (:40-41), compute_elf_map (:336-366) and
 * compute_patch_dwt_loss (:368-442) on one (1,C,H,W) render/GT pair.
 *   band_weights[8]: LL1,LH1,HL1,HH1,LL2,LH2,HL2,HH2 (LG/arguments/__init__.py:103-114).
 *   patch_size / percentile / patch_w_lh / patch_w_hl: LG/arguments/__init__.py:116-121; patch_size <= 0
 *   disables the patch term.
 *   out_losses (device, 12 floats): [0] weighted global DWT loss, [1] patch loss, [2..9] the eight unweighted
 *   band L1 means, [10] number of selected patches, [11] ELF threshold.
 *   patch_mask (device, u8[ceil-free floor(H/ps)*floor(W/ps)]) receives the ELF selection (1 = selected).
 *   workspace: device scratch of >= lg_dwt_workspace_bytes(C,H,W,patch_size) bytes.
 * The backward writes dL/dpred (C,H,W) for loss = g_dwt * out[0] + g_patch * out[1], using the mask and
 * selection count produced by the forward call on the same workspace.                                     */
size_t lg_dwt_workspace_bytes(int C, int H, int W, int patch_size);
int lg_dwt_loss_forward(const float* pred, const float* gt, int C, int H, int W,
                        const float* band_weights_host, int patch_size, double percentile,
                        float patch_w_lh, float patch_w_hl,
                        float* out_losses, uint8_t* patch_mask,
                        char* workspace, size_t workspace_bytes, void* stream);
int lg_dwt_loss_backward(const float* pred, const float* gt, int C, int H, int W,
                         const float* band_weights_host, int patch_size,
                         float patch_w_lh, float patch_w_hl,
                         const float* g_dwt_dev, const float* g_patch_dev,
                         const uint8_t* patch_mask, const float* out_losses,
                         float* dL_dpred, void* stream);

/* Fused Adam update of one flat fp32 parameter buffer (the (59 x P) buffer of the view-parallel trainer), one pass
 * instead of torch.optim.Adam's ~10 elementwise launches per group as the reference runs it (LG/train.py:278-288,
 * groups LG/scene/gaussian_model.py:178-211).  torch.optim.Adam semantics (no amsgrad / weight decay), bias
 * corrections from `step` (1-based), one learning rate per contiguous segment: segment s covers elements
 * [segment_ends[s-1], segment_ends[s]) (host arrays; the last end must equal n).  grad is read as grad * grad_scale
 * (e.g. 1 / views for the mean over a view batch). */
int lg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, int num_segments,
                 const long long* segment_ends, const float* lrs, float beta1, float beta2, float eps, int step,
                 float grad_scale, void* stream);

/* lg_adam_step with two learning rates inside a segment: segment s is a (rows, row_width[s]) array whose first
 * row_split[s] columns use lrs[s] and the rest lrs_b[s] (the SH slab (P, 16, 3): 3 f_dc floats at feature_lr, 45
 * f_rest floats at feature_lr / 20, LG/scene/gaussian_model.py:181-182).  row_width[s] <= 1 = a plain segment. */
int lg_adam_step_split(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                       int num_segments, const long long* segment_ends, const float* lrs, const float* lrs_b,
                       const int* row_width, const int* row_split, float beta1, float beta2, float eps, int step,
                       float grad_scale, void* stream);

/* The data-parallel exchange step over NVLink peer memory (no counterpart in the reference, which is single-GPU;
 * SURVEY.md §8e): gradient reduce-scatter + Adam + parameter all-gather as ONE kernel per rank (csrc/peer.cu).
 * Buffers are cudaMalloc allocations shared between the per-GPU processes with CUDA IPC:
 *   lg_peer_alloc/free     zero-filled device buffer that can be exported
 *   lg_peer_export         64-byte IPC handle of such a buffer (exchange it with any host-side all-gather)
 *   lg_peer_open/close     map a peer's buffer into this process (peer access enabled lazily)
 *   lg_peer_barrier        flag exchange between the `world` ranks on `stream` (flag_ptrs[q] = rank q's array of 16
 *                          u32 words (world <= 8), mapped here; `epoch` must grow by one per call); orders every rank's earlier
 *                          work on its stream before every rank's later work
 *   lg_peer_check          host-side: LG_ERR_CUDA if a barrier timed out (~30 s) since the last check
 *   lg_peer_reduce_adam    rank's shard = float4 range [n/4*rank/world, n/4*(rank+1)/world): g = sum_r grad_r (rank
 *                          order), Adam (lg_adam_step_split semantics; moments touched by the owner only), new
 *                          parameters stored into all ranks' parameter buffers.  Call between two lg_peer_barrier.
 *   lg_peer_allreduce      the reduce + broadcast alone (sum * grad_scale lands in every rank's gradient buffer) */
int lg_peer_alloc(size_t bytes, void** dev_ptr);
int lg_peer_free(void* dev_ptr);
int lg_peer_export(void* dev_ptr, unsigned char* handle64);
int lg_peer_open(const unsigned char* handle64, void** mapped);
int lg_peer_close(void* mapped);
int lg_peer_barrier(int rank, int world, void* const* flag_ptrs, unsigned epoch, void* stream);
int lg_peer_check(void* stream);
int lg_peer_reduce_adam(int rank, int world, void* const* grad_ptrs, void* const* param_ptrs, float* exp_avg,
                        float* exp_avg_sq, long long n, int num_segments, const long long* segment_ends,
                        const float* lrs, const float* lrs_b, const int* row_width, const int* row_split, float beta1,
                        float beta2, float eps, int step, float grad_scale, void* stream);
int lg_peer_allreduce(int rank, int world, void* const* grad_ptrs, long long n, float grad_scale, void* stream);
/* NVLS variants for buffers that are also mapped through an NVSwitch multicast object (mc_* = the multicast
 * addresses of the gradient / parameter regions, e.g. torch symmetric memory's multicast_ptr + offset): the sum over
 * ranks is one multimem.ld_reduce (added inside the switch), the delivery one multimem.st.  Same contract otherwise;
 * local_param = this rank's own (unicast) parameter region. */
int lg_peer_reduce_adam_mc(int rank, int world, const void* mc_grad, void* mc_param, const float* local_param,
                           float* exp_avg, float* exp_avg_sq, long long n, int num_segments,
                           const long long* segment_ends, const float* lrs, const float* lrs_b, const int* row_width,
                           const int* row_split, float beta1, float beta2, float eps, int step, float grad_scale,
                           void* stream);
int lg_peer_allreduce_mc(int rank, int world, void* mc_grad, long long n, float grad_scale, void* stream);

/* Adaptive density control on the flat field-major parameter buffer of the view-parallel trainer
 * (F = 11 + sh_floats floats per Gaussian: xyz 3 | SH (M,3) = f_dc then f_rest | opacity 1 | scaling 3 | rotation 4,
 * every field a contiguous (P, w) slab; the same buffer lg_adam_step updates).  Replaces the boolean-index + torch.cat
 * surgery of GaussianModel.densify_and_prune / densify_and_clone / densify_and_split / prune_points and the Adam-state
 * edits (LG/scene/gaussian_model.py:331-476), add_densification_stats (:478-480) with the max_radii2D update of
 * LG/train.py:268, and reset_opacity (:258-261).
 *
 * lg_densify_plan: per-Gaussian decisions from the raw scaling (P,3) / opacity (P,1) slabs and the statistics
 *   (grad_accum, denom: P floats), thresholds as the reference passes them (max_grad = densify_grad_threshold,
 *   min_opacity = 0.005, extent = cameras_extent, percent_dense, max_screen_size < 0 = None).  Writes, for every
 *   output row d, src_index[d] = source row | kind << 30 (0 original, 1 clone, 2/3 first/second split child) and, for
 *   children, eps_row[d] = row of the unit-normal sample; both arrays hold 2*P entries.  totals (device, 4 ints) =
 *   {kept originals, kept clones, split parents S, kept split parents}; P_new = t[0] + t[1] + 2*t[3]; the caller draws
 *   eps (2*S, 3) ~ N(0,1) (torch.normal's samples, :421-423) from a generator shared by all ranks.
 * lg_densify_apply: one gather pass per field writes the P_new-row parameter buffer and both Adam moments (moments of
 *   new rows zero, cat_tensors_to_optimizer :376-378).  Output order = the reference's (see csrc/densify.cu).
 *   Slab k of a buffer starts at (floats of the fields before it) * stride; stride >= rows lets the caller pad P to a
 *   multiple of 4 so that every slab is 16-byte aligned (the rasterizer reads rotations as float4).
 * The statistics of the new set are zeros of length P_new (densification_postfix :404-409): the caller re-allocates. */
size_t lg_densify_scratch_bytes(int P);
int lg_densify_plan(int P, const float* scaling, const float* opacity, const float* grad_accum, const float* denom,
                    float max_grad, float min_opacity, float extent, float percent_dense, float max_screen_size,
                    uint32_t* src_index, uint32_t* eps_row, int* totals, void* scratch, size_t scratch_bytes,
                    void* stream);
int lg_densify_apply(int P, int P_stride, int P_new, int P_new_stride, int sh_floats, const float* data,
                     const float* exp_avg,
                     const float* exp_avg_sq, float* data_new, float* exp_avg_new, float* exp_avg_sq_new,
                     const uint32_t* src_index, const uint32_t* eps_row, const float* eps, int eps_rows, void* stream);
/* visible = radii > 0:  grad_accum += |grad2D.xy|, denom += 1, max_radii2D = max(max_radii2D, radii) */
int lg_densify_stats(int P, const float* grad2D, const int* radii, float* grad_accum, float* denom,
                     float* max_radii2D, void* stream);
/* opacity <- inverse_sigmoid(min(sigmoid(opacity), 0.01)), Adam moments of the opacity slab <- 0 */
int lg_reset_opacity(int P, float* opacity, float* exp_avg, float* exp_avg_sq, void* stream);

/* l1_loss alone (LG/utils/loss_utils.py:40-41): out_loss[0] = mean |pred - gt| over n elements, one launch;
 * dL_dpred = g * sign(pred - gt) / n (g: device scalar), one launch.  scratch16: 16 bytes of device scratch. */
int lg_l1_loss_forward(const float* pred, const float* gt, long long n, float* out_loss, char* scratch16, void* stream);
int lg_l1_loss_backward(const float* pred, const float* gt, long long n, const float* g, float* dL_dpred, void* stream);

/* Fused photometric terms of the iteration's base loss (LG/train.py:128,182-188): out_losses[0] = mean |pred - gt|
 * (l1_loss, LG/utils/loss_utils.py:40-41), out_losses[1] = mean SSIM map (ssim/_ssim, LG/utils/loss_utils.py:58-86:
 * 11x11 Gaussian window sigma 1.5, zero "same" padding, C1 = 0.01^2, C2 = 0.03^2; same semantics as the optional
 * fusedssim CUDA path, LG/utils/loss_utils.py:16-37).  pred, gt: (C,H,W) device fp32.  With want_backward != 0 the
 * forward leaves three derivative maps in the workspace, which lg_photometric_loss_backward turns into
 * dL_dpred = g_l1 * d(l1)/dpred + g_ssim * d(ssim)/dpred (g_*: device scalars or NULL = 0).  workspace: device
 * scratch of >= lg_photometric_workspace_bytes(C,H,W) bytes, kept by the caller between forward and backward. */
size_t lg_photometric_workspace_bytes(int C, int H, int W);
int lg_photometric_loss_forward(const float* pred, const float* gt, int C, int H, int W, float* out_losses,
                                char* workspace, size_t workspace_bytes, int want_backward, void* stream);
int lg_photometric_loss_backward(const float* pred, const float* gt, int C, int H, int W, const char* workspace,
                                 const float* g_l1, const float* g_ssim, float* dL_dpred, void* stream);

/* Loss assembly of one training iteration on the device (LG/train.py:188-202), between the forward kernels above and
 * their backward: photometric_out = the two floats of lg_photometric_loss_forward, dwt_out = the out_losses of
 * lg_dwt_loss_forward.  base = (1 - lambda) * L1 + lambda * (1 - SSIM); with update_running_mean != 0 the running-mean
 * ratio of LG/train.py:190-196 is advanced in place (device scalar, no host round trip: the reference calls .item());
 * loss_out[0] = base + clamp(rm, 0.1, 10) * dwt + patch_weight * patch, loss_out[1] = base; coef_out[4] =
 * d(loss)/d(l1, ssim, dwt, patch).  lg_image_loss_backward_coefs multiplies them by the upstream gradient (device
 * scalar) -> the g_* inputs of the two backward entry points; lg_image_loss_add sums their two image gradients. */
int lg_image_loss_combine(const float* photometric_out, const float* dwt_out, float* running_mean, float lambda_dssim,
                          float patch_weight, int update_running_mean, float* loss_out, float* coef_out, void* stream);
int lg_image_loss_backward_coefs(const float* coef, const float* g, float* out4, void* stream);
int lg_image_loss_add(float* a, const float* b, long long n, void* stream);

/* The same iteration loss (LG/train.py:128-202) entered ONCE per direction: lg_image_loss_forward = the photometric and
 * wavelet forward kernels + lg_image_loss_combine; lg_image_loss_backward = the two gradient kernels, the second one
 * adding in place, each scaling its coefficient by the upstream gradient g_up (device scalar) itself.
 *   terms (device, 24 floats, written by the forward, read by the backward): [0] L1, [1] SSIM, [2..13] the out_losses of
 *   lg_dwt_loss_forward, [14] loss, [15] base, [16..19] d(loss)/d(L1, SSIM, dwt, patch), [20..23] spare.
 *   patch_mask / workspaces: as for the separate entry points; the photometric workspace, terms and patch_mask must be
 *   kept until the backward.  The *_scaled forms are the two backward kernels with the upstream factor and (wavelet)
 *   the accumulate switch exposed. */
int lg_image_loss_forward(const float* pred, const float* gt, int C, int H, int W, const float* band_weights_host,
                          int patch_size, double percentile, float patch_w_lh, float patch_w_hl, float* running_mean,
                          float lambda_dssim, float patch_weight, int update_running_mean, float* terms,
                          uint8_t* patch_mask, size_t patch_mask_bytes, char* photometric_workspace,
                          size_t photometric_bytes, char* dwt_workspace, size_t dwt_bytes, int want_backward, void* stream);
int lg_image_loss_backward(const float* pred, const float* gt, int C, int H, int W, const float* band_weights_host,
                           int patch_size, float patch_w_lh, float patch_w_hl, const float* terms, const float* g_up,
                           const uint8_t* patch_mask, const char* photometric_workspace, float* dL_dpred, void* stream);
int lg_photometric_loss_backward_scaled(const float* pred, const float* gt, int C, int H, int W, const char* workspace,
                                        const float* g_l1, const float* g_ssim, const float* g_up, float* dL_dpred,
                                        void* stream);
int lg_dwt_loss_backward_scaled(const float* pred, const float* gt, int C, int H, int W, const float* band_weights_host,
                                int patch_size, float patch_w_lh, float patch_w_hl, const float* g_dwt_dev,
                                const float* g_patch_dev, const float* g_up, const uint8_t* patch_mask,
                                const float* out_losses, float* dL_dpred, int accumulate, void* stream);

/* Single-level Haar analysis (the pytorch_wavelets.DWTForward(J=1,'symmetric','db1') call sites at
 * LG/utils/loss_utils.py:140-148).  x (N*C,H,W) -> ll (N*C,H2,W2), yh (N*C,3,H2,W2) with H2=(H+1)/2.
 * and its adjoint (for autograd through the compat module).                                              */
int lg_haar_dwt2_forward(const float* x, int planes, int H, int W, float* ll, float* yh, void* stream);
int lg_haar_dwt2_backward(const float* g_ll, const float* g_yh, int planes, int H, int W, float* g_x, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LGDWT_B200_H_INCLUDED */
