// abi.cu — the extern "C" boundary declared in include/lgdwt_b200.h: argument validation, opaque-state layout,
// stage orchestration (the job of CudaRasterizer::Rasterizer::forward/backward,
// DGR/cuda_rasterizer/rasterizer_impl.cu:198-450) and error reporting.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <nvtx3/nvToolsExt.h>

namespace lg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return LG_ERR_CUDA;
}

// ---------------------------------------------------------------- opaque state layouts
GeometryState GeometryState::from_chunk(char*& chunk, size_t P, int channels) {
    GeometryState g;
    carve(chunk, g.depths, P);
    carve(chunk, g.clamped, P * 3);
    carve(chunk, g.internal_radii, P);
    carve(chunk, g.means2D, P);
    carve(chunk, g.cov3D, P * 6);
    carve(chunk, g.conic_opacity, P);
    carve(chunk, g.rgb, P * (size_t)channels);
    carve(chunk, g.tiles_touched, P);
    carve(chunk, g.rect_packed, P);
    carve(chunk, g.grad_scratch, P * 12);
    return g;
}
size_t geometry_state_bytes(size_t P, int channels) {
    char* p = nullptr;
    GeometryState::from_chunk(p, P, channels);
    return (size_t)p + 128;
}

ImageState ImageState::from_chunk(char*& chunk, size_t W, size_t H) {
    ImageState s;
    const size_t T = (size_t)num_tiles_x((int)W) * num_tiles_y((int)H);
    carve(chunk, s.accum_alpha, W * H);
    carve(chunk, s.n_contrib, W * H);
    carve(chunk, s.ranges, T);
    carve(chunk, s.tile_order, T);
    carve(chunk, s.tile_neff, T);
    carve(chunk, s.tile_order_bwd, T);
    carve(chunk, s.counters, LG_CTR_STRIDE + T * LG_CTR_STRIDE);  // one allocation: one memset clears both
    s.tile_ctr = s.counters + LG_CTR_STRIDE;
    return s;
}
size_t image_state_bytes(size_t W, size_t H) {
    char* p = nullptr;
    ImageState::from_chunk(p, W, H);
    return (size_t)p + 128;
}

BinningState BinningState::from_chunk(char*& chunk, size_t R) {
    BinningState b;
    carve(chunk, b.pairs, R);
    carve(chunk, b.pairs_alt, R);
    carve(chunk, b.point_list, R);
    return b;
}
size_t binning_state_bytes(size_t R) {
    char* p = nullptr;
    BinningState::from_chunk(p, R);
    return (size_t)p + 128;
}

static thread_local uint32_t* g_pinned_word = nullptr;  // pinned staging for the num_rendered read-back
static thread_local cudaEvent_t g_count_event = nullptr;  // recorded behind that copy

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

// Per-stage CUDA-event timing with a ring of slots, so a benchmark can time every step of a run without a host
// sync inside the loop: each forward call advances to the next slot, the matching backward records into it.
#define LG_TIMING_SLOTS 256
static int g_timing_slots = 0;
static int g_slot = -1;
static int g_timing_every = 1;    // time one forward(+backward) call in `every` (lg_stage_timing_sample)
static long long g_timing_calls = 0;
static bool g_timing_active = false;
static cudaEvent_t g_ev[LG_TIMING_SLOTS][ST_COUNT][2];
static bool g_ev_used[LG_TIMING_SLOTS][ST_COUNT];
// NVTX ranges around the launches of every stage (host-side ranges named like lgdwt_b200.STAGES), for timeline tools;
// off unless LGDWT_NVTX=1 is set when the library is first used (nvtx3 is header-only and loads its injection library on
// demand, so nothing is linked)
static int nvtx_on() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("LGDWT_NVTX");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on;
}
static const char* const g_stage_names[ST_COUNT] = {"lg:preprocess", "lg:binning", "lg:blend_forward", "lg:blend_backward",
                                                     "lg:preprocess_backward"};
void stage_begin(int stage, cudaStream_t stream) {
    if (nvtx_on()) nvtxRangePushA(g_stage_names[stage]);
    if (g_timing_slots <= 0) return;
    if (stage == ST_PREPROCESS) {
        g_timing_active = (g_timing_calls++ % g_timing_every) == 0;
        if (g_timing_active) {
            g_slot = (g_slot + 1) % g_timing_slots;
            for (int s = 0; s < ST_COUNT; s++) g_ev_used[g_slot][s] = false;
        }
    }
    if (g_timing_active && g_slot >= 0) cudaEventRecord(g_ev[g_slot][stage][0], stream);
}
void stage_end(int stage, cudaStream_t stream) {
    if (nvtx_on()) nvtxRangePop();
    if (g_timing_slots <= 0 || g_slot < 0 || !g_timing_active) return;
    cudaEventRecord(g_ev[g_slot][stage][1], stream);
    g_ev_used[g_slot][stage] = true;
}

}  // namespace lg

using namespace lg;

extern "C" {

int lg_abi_version(void) { return LG_ABI_VERSION; }
const char* lg_last_error(void) { return lg::g_err; }

size_t lg_geometry_state_bytes(int P, int channels) { return geometry_state_bytes((size_t)P, channels); }
size_t lg_image_state_bytes(int width, int height) { return image_state_bytes((size_t)width, (size_t)height); }
size_t lg_binning_state_bytes(int num_rendered, int width, int height) {
    (void)width; (void)height;
    return binning_state_bytes((size_t)num_rendered);
}

int lg_rasterize_forward_hinted(lg_alloc_fn geometry_alloc, void* geometry_ctx, lg_alloc_fn binning_alloc,
                                void* binning_ctx, lg_alloc_fn image_alloc, void* image_ctx, int P, int D, int M,
                                int channels, const float* background, int width, int height, const float* means3D,
                                const float* shs, const float* colors_precomp, const float* opacities,
                                const float* scales, float scale_modifier, const float* rotations,
                                const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                                const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
                                float* out_invdepth, int antialiasing, int* radii, int debug, void* stream_v,
                                int capacity_hint, int* num_rendered, int* binning_capacity) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (binning_capacity) *binning_capacity = 0;
    if (num_rendered) *num_rendered = 0;
    if (P < 0 || width <= 0 || height <= 0 || !geometry_alloc || !binning_alloc || !image_alloc || !num_rendered) {
        set_error("lg_rasterize_forward: invalid sizes or missing allocator");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (channels < 1 || channels > LG_MAX_CHANNELS) {
        set_error("lg_rasterize_forward: channels must be in [1,%d], got %d", LG_MAX_CHANNELS, channels);
        return LG_ERR_UNSUPPORTED;
    }
    if (channels != 3 && colors_precomp == nullptr) {
        // same message as the reference (rasterizer_impl.cu:244-247)
        set_error("For non-RGB, provide precomputed Gaussian colors!");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (!background || !out_color || !viewmatrix || !projmatrix || !cam_pos) {
        set_error("lg_rasterize_forward: null camera/background/output pointer");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return LG_OK;  // reference: outputs keep their fill values (rasterize_points.cu:87-123)
    if (!means3D || !opacities || (!shs && !colors_precomp) || (!cov3D_precomp && (!scales || !rotations))) {
        set_error("lg_rasterize_forward: missing Gaussian parameter array");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (colors_precomp == nullptr && (M <= 0 || (D + 1) * (D + 1) > M)) {
        set_error("lg_rasterize_forward: SH degree %d needs %d coefficients, tensor has %d", D, (D + 1) * (D + 1), M);
        return LG_ERR_INVALID_ARGUMENT;
    }

    char* gchunk = geometry_alloc(geometry_ctx, geometry_state_bytes((size_t)P, channels));
    char* ichunk = image_alloc(image_ctx, image_state_bytes((size_t)width, (size_t)height));
    if (!gchunk || !ichunk) {
        set_error("lg_rasterize_forward: state allocation callback returned NULL");
        return LG_ERR_ALLOC;
    }
    GeometryState g = GeometryState::from_chunk(gchunk, (size_t)P, channels);
    ImageState img = ImageState::from_chunk(ichunk, (size_t)width, (size_t)height);
    if (radii == nullptr) radii = g.internal_radii;

    ForwardArgs f;
    f.P = P; f.D = D; f.M = M; f.C = channels; f.background = background; f.W = width; f.H = height;
    f.means3D = means3D; f.shs = shs; f.colors_precomp = colors_precomp; f.opacities = opacities; f.scales = scales;
    f.scale_modifier = scale_modifier; f.rotations = rotations; f.cov3D_precomp = cov3D_precomp;
    f.viewmatrix = viewmatrix; f.projmatrix = projmatrix; f.cam_pos = cam_pos; f.tan_fovx = tan_fovx;
    f.tan_fovy = tan_fovy;
    f.focal_y = height / (2.0f * tan_fovy);  // rasterizer_impl.cu:224-225
    f.focal_x = width / (2.0f * tan_fovx);
    f.prefiltered = prefiltered != 0; f.antialiasing = antialiasing != 0; f.debug = debug != 0;

    stage_begin(ST_PREPROCESS, stream);
    int rc = launch_preprocess(f, g, img, radii, stream);
    if (rc != LG_OK) return rc;
    stage_end(ST_PREPROCESS, stream);

    stage_begin(ST_BINNING, stream);
    rc = launch_tile_scan(width, height, img, f.debug, stream);
    if (rc != LG_OK) return rc;

    // num_rendered sizes the binning buffer, so it has to reach the host (rasterizer_impl.cu:283-288).  The reference
    // blocks right here.  With a capacity hint from the caller (typically 1.25 x the previous call's num_rendered) the
    // rest of the forward is queued FIRST, for a buffer of that capacity — every kernel behind this point reads
    // num_rendered on the device and does nothing if it exceeds the capacity — and only then does the host wait, on an
    // event recorded behind the copy, so the GPU never idles for the host's wake-up and launches.  If the hint was
    // too small (rare) the tail is simply queued again for the exact size.
    if (!g_pinned_word) {
        LG_CUDA(cudaMallocHost((void**)&g_pinned_word, 64));
        LG_CUDA(cudaEventCreateWithFlags(&g_count_event, cudaEventDisableTiming));
    }
    LG_CUDA(cudaMemcpyAsync(g_pinned_word, img.counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    LG_CUDA(cudaEventRecord(g_count_event, stream));
    const float* features = colors_precomp ? colors_precomp : g.rgb;
    int R = -1;
    int capacity = capacity_hint;
    if (capacity <= 0) {  // no hint: the reference's synchronous protocol
        LG_CUDA(cudaEventSynchronize(g_count_event));
        R = capacity = (int)*g_pinned_word;
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        char* bchunk = binning_alloc(binning_ctx, binning_state_bytes((size_t)capacity));
        if (!bchunk) {
            set_error("lg_rasterize_forward: binning allocation callback returned NULL");
            return LG_ERR_ALLOC;
        }
        BinningState b = BinningState::from_chunk(bchunk, (size_t)capacity);
        rc = launch_binning(P, capacity, width, height, g, radii, b, img, f.debug, stream);
        if (rc != LG_OK) return rc;
        if (attempt == 0) stage_end(ST_BINNING, stream);
        if (attempt == 0) stage_begin(ST_BLEND_FWD, stream);
        rc = launch_blend_forward(channels, width, height, capacity, g, b, img, features, background, out_color,
                                  out_invdepth, f.debug, stream);
        if (rc != LG_OK) return rc;
        if (attempt == 0) stage_end(ST_BLEND_FWD, stream);
        if (R < 0) {
            LG_CUDA(cudaEventSynchronize(g_count_event));
            R = (int)*g_pinned_word;
        }
        if (R <= capacity) break;
        capacity = R;  // the hint was too small: nothing behind the scan has run; queue it again at the exact size
    }
    *num_rendered = R;
    if (binning_capacity) *binning_capacity = capacity;
    return LG_OK;
}

int lg_rasterize_forward(lg_alloc_fn geometry_alloc, void* geometry_ctx, lg_alloc_fn binning_alloc, void* binning_ctx,
                         lg_alloc_fn image_alloc, void* image_ctx, int P, int D, int M, int channels,
                         const float* background, int width, int height, const float* means3D, const float* shs,
                         const float* colors_precomp, const float* opacities, const float* scales,
                         float scale_modifier, const float* rotations, const float* cov3D_precomp,
                         const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                         float tan_fovy, int prefiltered, float* out_color, float* out_invdepth, int antialiasing,
                         int* radii, int debug, void* stream_v, int* num_rendered) {
    return lg_rasterize_forward_hinted(geometry_alloc, geometry_ctx, binning_alloc, binning_ctx, image_alloc, image_ctx,
                                       P, D, M, channels, background, width, height, means3D, shs, colors_precomp,
                                       opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
                                       projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, out_invdepth,
                                       antialiasing, radii, debug, stream_v, 0, num_rendered, nullptr);
}

int lg_rasterize_backward_raw(int P, int D, int M, int R, int channels, const float* background, int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                          const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geometry_state,
                          char* binning_state, char* image_state, const float* dL_dpix,
                          const float* dL_dinvdepth_pix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolor, float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                          float* dL_dscale, float* dL_drot, int antialiasing, int debug, void* stream_v,
                          int accumulate, const float* raw_rot_norm) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (P < 0 || R < 0 || width <= 0 || height <= 0 || channels < 1 || channels > LG_MAX_CHANNELS) {
        set_error("lg_rasterize_backward: invalid sizes");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return LG_OK;
    if (!geometry_state || !binning_state || !image_state || !dL_dpix || !dL_dmean2D || !dL_dopacity ||
        !dL_dmean3D || !means3D || !opacities || !background || !viewmatrix || !projmatrix || !campos ||
        (colors_precomp && !dL_dcolor) || (cov3D_precomp && !dL_dcov3D)) {
        set_error("lg_rasterize_backward: missing required pointer");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if ((!cov3D_precomp && (!scales || !rotations || !dL_dscale || !dL_drot)) || (shs && !dL_dsh)) {
        set_error("lg_rasterize_backward: missing parameter / gradient array for the active input path");
        return LG_ERR_INVALID_ARGUMENT;
    }
    char* gp = geometry_state;
    char* bp = binning_state;
    char* ip = image_state;
    GeometryState g = GeometryState::from_chunk(gp, (size_t)P, channels);
    BinningState b = BinningState::from_chunk(bp, (size_t)R);
    ImageState img = ImageState::from_chunk(ip, (size_t)width, (size_t)height);
    if (radii == nullptr) radii = g.internal_radii;
    const float* features = colors_precomp ? colors_precomp : g.rgb;

    stage_begin(ST_BLEND_BWD, stream);
    int rc = launch_blend_backward(P, channels, width, height, g, b, img, features, background, dL_dpix,
                                   dL_dinvdepth_pix, g.grad_scratch, debug != 0, stream);
    if (rc != LG_OK) return rc;
    stage_end(ST_BLEND_BWD, stream);

    BackwardArgs a;
    a.P = P; a.D = D; a.M = M; a.C = channels; a.W = width; a.H = height; a.means3D = means3D; a.shs = shs;
    a.colors_precomp = colors_precomp; a.opacities = opacities; a.scales = cov3D_precomp ? nullptr : scales;
    a.scale_modifier = scale_modifier; a.rotations = cov3D_precomp ? nullptr : rotations;
    a.cov3D_precomp = cov3D_precomp; a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.campos = campos;
    a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
    a.focal_y = height / (2.0f * tan_fovy);
    a.focal_x = width / (2.0f * tan_fovx);
    a.antialiasing = antialiasing != 0; a.has_invdepth = dL_dinvdepth_pix != nullptr; a.accumulate = accumulate != 0;
    a.raw_rot_norm = raw_rot_norm;
    a.dL_dmean2D = dL_dmean2D; a.dL_dconic = dL_dconic; a.dL_dopacity = dL_dopacity; a.dL_dcolor = dL_dcolor;
    a.dL_dinvdepth = dL_dinvdepth; a.dL_dmean3D = dL_dmean3D; a.dL_dcov3D = dL_dcov3D; a.dL_dsh = dL_dsh;
    a.dL_dscale = dL_dscale; a.dL_drot = dL_drot;
    stage_begin(ST_PREGRAD, stream);
    rc = launch_preprocess_backward(a, g, radii, debug != 0, stream);
    stage_end(ST_PREGRAD, stream);
    return rc;
}

int lg_rasterize_backward_ex(int P, int D, int M, int R, int channels, const float* background, int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                          const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geometry_state,
                          char* binning_state, char* image_state, const float* dL_dpix,
                          const float* dL_dinvdepth_pix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolor, float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                          float* dL_dscale, float* dL_drot, int antialiasing, int debug, void* stream_v,
                          int accumulate) {
    return lg_rasterize_backward_raw(P, D, M, R, channels, background, width, height, means3D, shs, colors_precomp,
                                     opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                                     campos, tan_fovx, tan_fovy, radii, geometry_state, binning_state, image_state,
                                     dL_dpix, dL_dinvdepth_pix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                     dL_dinvdepth, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, antialiasing,
                                     debug, stream_v, accumulate, nullptr);
}

int lg_rasterize_backward(int P, int D, int M, int R, int channels, const float* background, int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                          const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geometry_state,
                          char* binning_state, char* image_state, const float* dL_dpix,
                          const float* dL_dinvdepth_pix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolor, float* dL_dinvdepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                          float* dL_dscale, float* dL_drot, int antialiasing, int debug, void* stream_v) {
    return lg_rasterize_backward_ex(P, D, M, R, channels, background, width, height, means3D, shs, colors_precomp,
                                    opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                                    campos, tan_fovx, tan_fovy, radii, geometry_state, binning_state, image_state,
                                    dL_dpix, dL_dinvdepth_pix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor,
                                    dL_dinvdepth, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, antialiasing,
                                    debug, stream_v, 0);
}

int lg_blend_work_count(int P, int channels, int width, int height, int R, const char* geometry_state,
                        const char* binning_state, const char* image_state, unsigned long long* counts_dev,
                        void* stream_v) {
    if (!geometry_state || !binning_state || !image_state || !counts_dev || P <= 0) {
        set_error("lg_blend_work_count: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    char* gp = const_cast<char*>(geometry_state);
    char* bp = const_cast<char*>(binning_state);
    char* ip = const_cast<char*>(image_state);
    GeometryState g = GeometryState::from_chunk(gp, (size_t)P, channels);
    BinningState b = BinningState::from_chunk(bp, (size_t)R);
    ImageState img = ImageState::from_chunk(ip, (size_t)width, (size_t)height);
    return launch_blend_count(width, height, g, b, img, counts_dev, (cudaStream_t)stream_v);
}

unsigned long long lg_launch_count(void) { return __atomic_load_n(&lg::g_launches, __ATOMIC_RELAXED); }

int lg_stage_timing_enable(int slots) {
    if (slots > LG_TIMING_SLOTS) slots = LG_TIMING_SLOTS;
    if (slots < 0) slots = 0;
    for (int i = 0; i < lg::g_timing_slots; i++)
        for (int s = 0; s < ST_COUNT; s++)
            for (int k = 0; k < 2; k++) cudaEventDestroy(lg::g_ev[i][s][k]);
    lg::g_timing_slots = 0;
    lg::g_slot = -1;
    lg::g_timing_calls = 0;
    lg::g_timing_active = false;
    for (int i = 0; i < slots; i++)
        for (int s = 0; s < ST_COUNT; s++) {
            for (int k = 0; k < 2; k++) LG_CUDA(cudaEventCreate(&lg::g_ev[i][s][k]));
            lg::g_ev_used[i][s] = false;
        }
    lg::g_timing_slots = slots;
    return LG_OK;
}

int lg_stage_timing_sample(int every) {
    lg::g_timing_every = every < 1 ? 1 : every;
    return LG_OK;
}

int lg_stage_timing_read(int slot, float* ms_out, int n) {
    if (slot < 0 || slot >= lg::g_timing_slots || !ms_out || n < ST_COUNT) {
        set_error("lg_stage_timing_read: bad slot %d (enabled slots %d) or output too small (need %d floats)", slot,
                  lg::g_timing_slots, (int)ST_COUNT);
        return LG_ERR_INVALID_ARGUMENT;
    }
    for (int s = 0; s < ST_COUNT; s++) {
        ms_out[s] = -1.0f;
        if (!lg::g_ev_used[slot][s]) continue;
        LG_CUDA(cudaEventSynchronize(lg::g_ev[slot][s][1]));
        LG_CUDA(cudaEventElapsedTime(&ms_out[s], lg::g_ev[slot][s][0], lg::g_ev[slot][s][1]));
    }
    return LG_OK;
}

int lg_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                    void* stream_v) {
    (void)projmatrix;  // the reference computes p_proj and never uses it (auxiliary.h:163, nvcc warning #177)
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("lg_mark_visible: invalid arguments");
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return LG_OK;
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream_v);
}

int lg_state_read(const char* name, int P, int channels, int width, int height, int R, const char* geometry_state,
                  const char* binning_state, const char* image_state, void* dst, size_t dst_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    const void* src = nullptr;
    size_t bytes = 0;
    const size_t N = (size_t)width * height, T = (size_t)num_tiles_x(width) * num_tiles_y(height);
    if (geometry_state) {
        char* p = const_cast<char*>(geometry_state);
        GeometryState g = GeometryState::from_chunk(p, (size_t)P, channels);
        if (!strcmp(name, "depths")) { src = g.depths; bytes = 4 * (size_t)P; }
        else if (!strcmp(name, "means2D")) { src = g.means2D; bytes = 8 * (size_t)P; }
        else if (!strcmp(name, "cov3D")) { src = g.cov3D; bytes = 24 * (size_t)P; }
        else if (!strcmp(name, "conic_opacity")) { src = g.conic_opacity; bytes = 16 * (size_t)P; }
        else if (!strcmp(name, "rgb")) { src = g.rgb; bytes = 4 * (size_t)channels * P; }
        else if (!strcmp(name, "clamped")) { src = g.clamped; bytes = 3 * (size_t)P; }
        else if (!strcmp(name, "tiles_touched")) { src = g.tiles_touched; bytes = 4 * (size_t)P; }
        else if (!strcmp(name, "point_offsets")) {
            // the reference's inclusive scan of tiles_touched is no longer a pipeline product: rebuilt for inspection
            if (dst_bytes < 4 * (size_t)P) {
                set_error("lg_state_read: 'point_offsets' needs %zu bytes", 4 * (size_t)P);
                return LG_ERR_INVALID_ARGUMENT;
            }
            return launch_point_offsets(P, g, (uint32_t*)dst, stream);
        }
        else if (!strcmp(name, "grad_record")) { src = g.grad_scratch; bytes = 48 * (size_t)P; }
    }
    if (!src && image_state) {
        char* p = const_cast<char*>(image_state);
        ImageState s = ImageState::from_chunk(p, (size_t)width, (size_t)height);
        if (!strcmp(name, "final_T")) { src = s.accum_alpha; bytes = 4 * N; }
        else if (!strcmp(name, "n_contrib")) { src = s.n_contrib; bytes = 4 * N; }
        else if (!strcmp(name, "ranges")) { src = s.ranges; bytes = 8 * T; }
    }
    if (!src && binning_state) {
        char* p = const_cast<char*>(binning_state);
        BinningState b = BinningState::from_chunk(p, (size_t)R);
        if (!strcmp(name, "point_list")) { src = b.point_list; bytes = 4 * (size_t)R; }
        else if (!strcmp(name, "point_list_keys")) {
            // the 64-bit (tile | depth) keys of the reference are never materialised (binning.cu); rebuild them from
            // the tile ranges and the depths for inspection
            if (!geometry_state || dst_bytes < 8 * (size_t)R) {
                set_error("lg_state_read: 'point_list_keys' needs the geometry state and %zu bytes", 8 * (size_t)R);
                return LG_ERR_INVALID_ARGUMENT;
            }
            char* gp = const_cast<char*>(geometry_state);
            GeometryState g = GeometryState::from_chunk(gp, (size_t)P, channels);
            if (!image_state) {
                set_error("lg_state_read: 'point_list_keys' needs the image state (tile ranges)");
                return LG_ERR_INVALID_ARGUMENT;
            }
            char* ip2 = const_cast<char*>(image_state);
            ImageState s2 = ImageState::from_chunk(ip2, (size_t)width, (size_t)height);
            return launch_rebuild_keys(width, height, g, b, s2, (unsigned long long*)dst, stream);
        }
    }
    if (!src) {
        set_error("lg_state_read: unknown array '%s' (or its state buffer was not given)", name);
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (dst_bytes < bytes) {
        set_error("lg_state_read: '%s' needs %zu bytes, destination has %zu", name, bytes, dst_bytes);
        return LG_ERR_INVALID_ARGUMENT;
    }
    if (bytes) LG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream));
    return LG_OK;
}

}  // extern "C"
