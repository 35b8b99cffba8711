// sb_api.cu — the extern "C" layer: handles, buffers, uniforms, stage sequencing.
//
// Mirrors the reference's object model (src/lib.rs:65-276, src/multi_model.rs:291-531): a Viewer
// owns the 8 public buffers and chains preprocess -> sort -> render; nothing here synchronises
// with the device except the explicit read_* helpers.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "sb_internal.h"

namespace {

thread_local std::string g_last_error;

struct Ctx {
    int device = 0;
    int num_sms = 0;
    uint64_t model_size_limit = 0;
    std::string last_error;
};

SbStatus fail(Ctx* ctx, SbStatus s, const std::string& msg) {
    g_last_error = msg;
    if (ctx) ctx->last_error = msg;
    return s;
}
SbStatus fail_cuda(Ctx* ctx, cudaError_t e, const char* what) {
    return fail(ctx, SB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define SB_CUDA(ctx, call)                                            \
    do {                                                              \
        cudaError_t _e = (call);                                      \
        if (_e != cudaSuccess) return fail_cuda(ctx, _e, #call);      \
    } while (0)

// Every entry point that enqueues, allocates or copies runs on ITS context's device, whatever the calling thread's current
// device is (several contexts — one per GPU — may live in one process); the previous device is restored on return.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const Ctx* c) {
        if (!c) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

bool is_unorm(int fmt) { return fmt == SB_TARGET_RGBA8_UNORM || fmt == SB_TARGET_BGRA8_UNORM; }  // linear 8-bit
bool is_srgb(int fmt) { return fmt == SB_TARGET_RGBA8_UNORM_SRGB || fmt == SB_TARGET_BGRA8_UNORM_SRGB; }
uint32_t bytes_per_pixel(int fmt) { return (is_unorm(fmt) || is_srgb(fmt)) ? 4u : fmt == SB_TARGET_RGBA16_FLOAT ? 8u : 16u; }

// ---- strict host maths for the per-frame uniforms (this file is compiled with
// -Xcompiler -ffp-contract=off): the same individually rounded operations, in the same order,
// that the oracle contract fixes for the uniform part of the reference shaders.
void mat4_mul(const float* a, const float* b, float* r) {  // column-major
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++)
            r[j * 4 + i] = ((a[i] * b[j * 4] + a[4 + i] * b[j * 4 + 1]) + a[8 + i] * b[j * 4 + 2]) + a[12 + i] * b[j * 4 + 3];
}

sb::Uniforms make_uniforms(const SbCameraPod& cam, const SbModelTransformPod& mt, const SbGaussianTransformPod& gt, int target_format,
                           bool alpha_cut = false) {
    sb::Uniforms u;
    std::memset(&u, 0, sizeof u);
    // Mat3::from_quat
    const float x = mt.rot[0], y = mt.rot[1], z = mt.rot[2], w = mt.rot[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
    const float wx = w * x2, wy = w * y2, wz = w * z2;
    const float r[9] = {1.0f - (yy + zz), xy + wz, xz - wy, xy - wz, 1.0f - (xx + zz), yz + wx, xz + wy, yz - wx, 1.0f - (xx + yy)};
    for (int c = 0; c < 3; c++)
        for (int rr = 0; rr < 3; rr++) {
            u.sr[c * 3 + rr] = r[c * 3 + rr] * mt.scale[c];      // R S
            u.inv_sr[c * 3 + rr] = r[rr * 3 + c] / mt.scale[rr];  // S^-1 R^T
        }
    for (int c = 0; c < 3; c++)
        for (int rr = 0; rr < 3; rr++) u.model[c * 4 + rr] = u.sr[c * 3 + rr];
    u.model[12] = mt.pos[0];
    u.model[13] = mt.pos[1];
    u.model[14] = mt.pos[2];
    u.model[15] = 1.0f;
    mat4_mul(cam.proj, cam.view, u.pv);
    mat4_mul(cam.view, u.model, u.vm);
    for (int c = 0; c < 3; c++)
        for (int rr = 0; rr < 3; rr++) u.w[c * 3 + rr] = cam.view[c * 4 + rr];
    u.size[0] = cam.size[0];
    u.size[1] = cam.size[1];
    u.focal[0] = cam.proj[0] * cam.size[0] * 0.5f;
    u.focal[1] = cam.proj[5] * cam.size[1] * 0.5f;
    for (int i = 0; i < 3; i++) {
        const float d = (u.w[i * 3] * cam.view[12] + u.w[i * 3 + 1] * cam.view[13]) + u.w[i * 3 + 2] * cam.view[14];
        u.cam_pos[i] = -d;
    }
    u.std_dev = (float)gt.max_std_dev / 255.0f * 3.0f;
    u.gsize = gt.size;
    u.color_scale = is_unorm(target_format) ? 255.0f : 1.0f;
    u.color_max = is_unorm(target_format) ? 255.0f : is_srgb(target_format) ? 1.0f : std::numeric_limits<float>::infinity();
    // exact alpha cut-off: only where a blend below the threshold is provably the identity (sb_common.cuh)
    u.cut_k = (alpha_cut && is_unorm(target_format) && gt.display_mode == SB_MODE_SPLAT) ? 1.0f / sb::kAlphaCut : 0.0f;
    u.mode = gt.display_mode;
    u.sh_deg = gt.sh_deg;
    u.no_sh0 = gt.no_sh0;
    u.width = (uint32_t)cam.size[0];
    u.height = (uint32_t)cam.size[1];
    u.tiles_x = (u.width + sb::kTile - 1) / sb::kTile;
    u.tiles_y = (u.height + sb::kTile - 1) / sb::kTile;
    return u;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool encode_recs_map(void* d_recs, uint64_t n, CUtensorMap* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    // Rows are declared 16 floats wide over the 48-byte record stride: each fetched row is a record plus the first 16 bytes
    // of the next one, so the four rows of a gather4 land at a 64-byte pitch in shared memory and the rasterizer addresses
    // record j at 64 j (the array is allocated 64 bytes longer for the last row's overhang).
    const cuuint64_t gdim[2] = {16, n ? n : 1};
    const cuuint64_t gstride[1] = {sizeof(sb::SplatRec)};
    const cuuint32_t box[2] = {16, 1};  // gather4 fetches four 1-row boxes (probed: scripts/probes/probe_gather4.cu)
    const cuuint32_t estr[2] = {1, 1};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_recs, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct DeviceBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct SbContext : Ctx {};

// One model's buffers and stage scratch.  A Viewer is one of these plus camera/gaussian
// transform pods; a MultiModelViewer shares the world pods across many.
struct SbViewer {
    SbContext* ctx = nullptr;
    int sh_fmt = 0, cov_fmt = 0, target_format = 0;
    uint32_t stride = 0;
    uint32_t n = 0;
    uint32_t padded = 0;

    const void* d_gaussians = nullptr;  // owned (gaussians_owned) or adopted
    DeviceBuf gaussians_owned;
    DeviceBuf indices, keys, args, recs, tboxes, pre_scratch;
    DeviceBuf sort_keys_alt, sort_vals_alt, sort_internal;
    DeviceBuf depth_keys_alt, depth_vals_alt;  // the depth sort's own ping-pong buffers: its result may stay there (sorted_pending)
    DeviceBuf dup_offsets, tboxes_sorted, win_first, dup_keys, dup_vals, tile_recs, tile_ranges, tile_order, bin_state;
    DeviceBuf selection;
    DeviceBuf orig_colors;  // NonDestructiveModifier's source copy (colour words), taken at the first edit
    DeviceBuf internal_target;
    uint64_t dup_capacity = 0;
    uint32_t tile_capacity = 0;
    volatile uint32_t* h_needed = nullptr;  // mapped pinned: [0] duplicates the last binned frame needed (saturated), [1] overflow events,
    uint32_t* d_needed = nullptr;           //                [4] visible count of the last preprocess (size hint for the sort)
    uint32_t overflow_seen = 0;             // overflow events already reported to the caller

    SbCameraPod camera;
    SbModelTransformPod model_transform;
    SbGaussianTransformPod gaussian_transform;
    bool selection_enabled = false;
    uint32_t invert_selection = 1;  // src/selection/buffer.rs:157-165
    int strict_exp = 0;
    uint32_t last_tiles_x = 0, last_tiles_y = 0;  // tile grid of the last binned frame (tile_ranges is indexed by it)
    bool strip_cull = false;     // a strip render keeps only the splats whose tile box meets the strip (sb_viewer_set_strip_cull)
    bool exact_cutoff = true;    // shrink splats to the radius beyond which a unorm8 blend is exactly the identity
    bool recs_cut = false;       // recs/tboxes currently hold cut extents (a depth-tested pass needs the full ones)
    // Viewer::render leaves the depth-sorted (key, index) pairs where the last sort pass wrote them (*d_sort_parity() = 1:
    // the alt buffers) and binning reads them there; they are copied home to indices/keys only when somebody asks for
    // them (read_*, *_ptr, the stage-wise sort call) — 48 MB less traffic per 6 M-Gaussian frame.
    bool sorted_pending = false;
    bool sort_prep_fresh = false;  // the depth sort's prep words were zeroed (with K1's scratch) and not consumed yet
    cudaStream_t last_stream = nullptr;
    CUtensorMap recs_map;        // 2-D view of recs[] (12 x n floats, 48-byte rows) for TMA gather4
    bool use_gather4 = false;
    // view-batch pipelining (sb_viewer_render_batch): a twin set of frame buffers over the same pods,
    // two internal streams forked from / joined to the caller's stream
    SbViewer* twin = nullptr;
    const uint32_t* selection_override = nullptr;  // twin: the primary's selection words
    // standalone Renderer: the caller's IndirectIndicesBuffer / IndirectArgsBuffer.instance_count
    const uint32_t* ext_indices = nullptr;
    const uint32_t* ext_count = nullptr;
    cudaStream_t bstream[2] = {nullptr, nullptr};
    cudaEvent_t bevent[3] = {nullptr, nullptr, nullptr};
    bool timing = false;
    bool counting = false;
    DeviceBuf counters;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    SbDrawIndirectArgs* d_draw() const { return args.as<SbDrawIndirectArgs>(); }
    SbDispatchIndirectArgs* d_dispatch() const { return reinterpret_cast<SbDispatchIndirectArgs*>(args.as<uint8_t>() + 16); }
    uint32_t* d_visible() const { return reinterpret_cast<uint32_t*>(args.as<uint8_t>() + 32); }
    uint32_t* d_sort_parity() const { return reinterpret_cast<uint32_t*>(args.as<uint8_t>() + 36); }
    // bin_state: [dup_count 4][overflow 4][pad 8][scan_counter 16][scan_status ...]
    uint32_t* d_dup_count() const { return bin_state.as<uint32_t>(); }
    uint32_t* d_overflow() const { return bin_state.as<uint32_t>() + 1; }
    uint32_t* d_scan_counter() const { return bin_state.as<uint32_t>() + 4; }
    unsigned long long* d_scan_status() const { return reinterpret_cast<unsigned long long*>(bin_state.as<uint8_t>() + 32); }
};

namespace {

SbStatus viewer_alloc(SbViewer* v) {
    SbContext* ctx = v->ctx;
    const uint32_t n = v->n;
    v->padded = sb_padded_key_count(n);
    SB_CUDA(ctx, v->indices.alloc((size_t)(v->padded ? v->padded : 1) * 4));
    SB_CUDA(ctx, v->keys.alloc((size_t)(v->padded ? v->padded : 1) * 4));
    SB_CUDA(ctx, v->depth_keys_alt.alloc((size_t)(v->padded ? v->padded : 1) * 4));
    SB_CUDA(ctx, v->depth_vals_alt.alloc((size_t)(v->padded ? v->padded : 1) * 4));
    SB_CUDA(ctx, v->args.alloc(64));
    SB_CUDA(ctx, v->recs.alloc((size_t)(n ? n : 1) * sizeof(sb::SplatRec) + 64));
    SB_CUDA(ctx, v->tboxes.alloc((size_t)(n ? n : 1) * sizeof(sb::TileBox)));
    {
        // rasterizer record fetch: TMA gather4 from recs[] (default) or a gathered copy + 1-D bulk copies
        const char* path = std::getenv("SB_RASTER_PATH");
        const bool want_bulk = path && std::string(path) == "bulk";
        v->use_gather4 = !want_bulk && encode_recs_map(v->recs.p, n, &v->recs_map);
    }
    // [depth sort prep words | K1 ticket + look-back]: zeroed together by the one memset in front of K1
    SB_CUDA(ctx, v->pre_scratch.alloc(sb::preprocess_scratch_prefix_bytes() + sb::preprocess_scratch_bytes(n, v->sh_fmt, v->cov_fmt)));
    SB_CUDA(ctx, v->selection.alloc(((size_t)n + 31) / 32 * 4 + 4));
    SB_CUDA(ctx, cudaMemset(v->selection.p, 0, v->selection.bytes));
    const SbDrawIndirectArgs d = {6, 0, 0, 0};          // IndirectArgsBuffer::new
    const SbDispatchIndirectArgs s = {1, 1, 1};         // RadixSortIndirectArgsBuffer::new
    uint8_t init[64] = {0};
    std::memcpy(init, &d, sizeof d);
    std::memcpy(init + 16, &s, sizeof s);
    SB_CUDA(ctx, cudaMemcpy(v->args.p, init, sizeof init, cudaMemcpyHostToDevice));
    SB_CUDA(ctx, v->dup_offsets.alloc(((size_t)n + 1) * 4));
    SB_CUDA(ctx, v->tboxes_sorted.alloc((size_t)(n ? n : 1) * sizeof(sb::TileBox)));
    const size_t scan_tiles = ((size_t)n + 1023) / 1024 + 2;
    SB_CUDA(ctx, v->bin_state.alloc(32 + scan_tiles * sizeof(unsigned long long)));
    SB_CUDA(ctx, cudaMemset(v->bin_state.p, 0, v->bin_state.bytes));
    void* hp = nullptr;
    SB_CUDA(ctx, cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
    std::memset(hp, 0, 64);
    v->h_needed = static_cast<volatile uint32_t*>(hp);
    void* dp = nullptr;
    SB_CUDA(ctx, cudaHostGetDevicePointer(&dp, hp, 0));
    v->d_needed = static_cast<uint32_t*>(dp);
    return SB_OK;
}

SbStatus viewer_reserve(SbViewer* v, uint64_t cap) {
    SbContext* ctx = v->ctx;
    if (cap < 4096) cap = 4096;
    if (cap > 0x3fffffffull) cap = 0x3fffffffull;
    v->dup_capacity = cap;
    SB_CUDA(ctx, v->dup_keys.alloc(cap * 4 + 16));  // tile_ranges_kernel reads whole uint4s
    SB_CUDA(ctx, v->dup_vals.alloc(cap * 4));
    SB_CUDA(ctx, v->win_first.alloc((cap / 4096 + 2) * 4));  // one entry per 4096-duplicate emit window
    if (!v->use_gather4) SB_CUDA(ctx, v->tile_recs.alloc(cap * sizeof(sb::SplatRec)));
    const uint64_t sort_cap = cap > v->padded ? cap : v->padded;
    SB_CUDA(ctx, v->sort_keys_alt.alloc(sort_cap * 4));
    SB_CUDA(ctx, v->sort_vals_alt.alloc(sort_cap * 4));
    SB_CUDA(ctx, v->sort_internal.alloc(sb::sort_internal_bytes((uint32_t)sort_cap)));
    return SB_OK;
}

SbStatus viewer_new(SbContext* ctx, int sh_fmt, int cov_fmt, int target_format, uint64_t n, SbViewer** out) {
    if (!ctx || !out) return fail(ctx, SB_ERR_INVALID_ARG, "null context/out");
    if (sb_pod_stride(sh_fmt, cov_fmt) == 0) return fail(ctx, SB_ERR_INVALID_ARG, "unknown pod format");
    if (target_format < 0 || target_format > SB_TARGET_BGRA8_UNORM_SRGB) return fail(ctx, SB_ERR_INVALID_ARG, "unknown target format");
    if (n > 0x3fffffffull) return fail(ctx, SB_ERR_MODEL_TOO_LARGE, "more than 2^30 gaussians");
    const uint64_t model_size = n * sb_pod_stride(sh_fmt, cov_fmt);
    if (model_size > ctx->model_size_limit) {  // src/preprocessor.rs:239-246
        char msg[160];
        std::snprintf(msg, sizeof msg, "model size %llu exceeds the device limit %llu", (unsigned long long)model_size,
                      (unsigned long long)ctx->model_size_limit);
        return fail(ctx, SB_ERR_MODEL_TOO_LARGE, msg);
    }
    SbViewer* v = new SbViewer();
    v->ctx = ctx;
    v->sh_fmt = sh_fmt;
    v->cov_fmt = cov_fmt;
    v->target_format = target_format;
    v->stride = sb_pod_stride(sh_fmt, cov_fmt);
    v->n = (uint32_t)n;
    std::memset(&v->camera, 0, sizeof v->camera);
    const float zero[3] = {0, 0, 0}, one[3] = {1, 1, 1}, ident[4] = {0, 0, 0, 1};
    sb_model_transform_pod(zero, ident, one, &v->model_transform);                // identity default
    sb_gaussian_transform_pod(1.0f, SB_MODE_SPLAT, 3, 0, 3.0f, &v->gaussian_transform);  // core defaults
    SbStatus s = viewer_alloc(v);
    if (s == SB_OK) s = viewer_reserve(v, 8ull * n + 65536);
    if (s != SB_OK) {
        sb_viewer_destroy(v);
        return s;
    }
    *out = v;
    return SB_OK;
}

// A frame whose (splat, tile) duplicates did not fit dropped its nearest splats.  The device reports it through a mapped pinned
// word (no synchronisation); the NEXT enqueue that notices grows the buffers and returns SB_ERR_OVERFLOW WITHOUT enqueuing
// anything, so the caller learns that an earlier frame was incomplete and simply renders again (now with room).
SbStatus report_overflow(SbViewer* v) {
    const uint32_t events = v->h_needed[1];
    if (events == v->overflow_seen) return SB_OK;
    v->overflow_seen = events;
    const uint64_t need = v->h_needed[0];
    SB_CUDA(v->ctx, cudaDeviceSynchronize());
    if (need > v->dup_capacity) {
        SbStatus gs = viewer_reserve(v, need + need / 2);
        if (gs != SB_OK) return gs;
    }
    *v->h_needed = 0;
    char msg[200];
    std::snprintf(msg, sizeof msg, "a previous frame needed %s%llu tile duplicates and was rendered incompletely; capacity grown to %llu, render again",
                  need == 0xffffffffull ? ">= " : "", (unsigned long long)need, (unsigned long long)v->dup_capacity);
    return fail(v->ctx, SB_ERR_OVERFLOW, msg);
}

SbStatus check_target(SbViewer* v, const SbTarget* t, const sb::Uniforms& u) {
    if (!t || !t->d_pixels) return fail(v->ctx, SB_ERR_INVALID_ARG, "null target");
    if (t->format != v->target_format) return fail(v->ctx, SB_ERR_INVALID_ARG, "target format differs from the viewer's texture_format");
    if (t->width != u.width || t->height != u.height) return fail(v->ctx, SB_ERR_INVALID_ARG, "target size differs from CameraPod.size");
    if (t->pitch_bytes < t->width * bytes_per_pixel(t->format)) return fail(v->ctx, SB_ERR_BAD_BUFFER_SIZE, "target pitch too small");
    if (t->rows != 0 && (t->row0 >= t->height || t->rows > t->height - t->row0)) return fail(v->ctx, SB_ERR_INVALID_ARG, "bad strip");
    return SB_OK;
}

SbStatus do_preprocess(SbViewer* v, const SbCameraPod& cam, const SbGaussianTransformPod& gt, cudaStream_t stream,
                       const SbTarget* strip = nullptr, uint32_t slice_lo = 0, uint32_t slice_n = 0xffffffffu) {
    sb::PreParams p;
    std::memset(&p, 0, sizeof p);
    if (slice_n == 0xffffffffu) slice_n = v->n;  // the whole model (a slice: the partitioned strips, sb_strips_scatter)
    p.gaussians = static_cast<const uint8_t*>(v->d_gaussians) + (size_t)slice_lo * v->stride;
    p.n = slice_n;
    p.index_base = slice_lo;
    p.selection = v->selection_enabled ? (v->selection_override ? v->selection_override : v->selection.as<uint32_t>()) : nullptr;
    p.invert_selection = v->invert_selection;
    p.indices = v->indices.as<uint32_t>();
    p.keys = v->keys.as<float>();
    p.keys_capacity = v->padded;
    p.draw_args = v->d_draw();
    p.sort_args = v->d_dispatch();
    p.recs = v->recs.as<sb::SplatRec>() + slice_lo;
    p.tboxes = v->tboxes.as<sb::TileBox>() + slice_lo;
    p.visible_count = v->d_visible();
    p.visible_host = v->d_needed + 4;  // mapped pinned: the next frame's sort reads it (no synchronisation) as a size hint
    p.sort_prep = v->pre_scratch.as<uint32_t>();
    v->sort_prep_fresh = true;         // zeroed by K1's memset, or / nand accumulated by K1: the next depth sort needs no clearing
    p.u = make_uniforms(cam, v->model_transform, gt, v->target_format, v->exact_cutoff);
    if (strip && strip->rows != 0) {  // strip render with strip culling: this rank's visible set is its strip's subset
        p.strip_on = 1;
        p.strip_ty_lo = strip->row0 / sb::kTile;
        p.strip_ty_hi = (strip->row0 + strip->rows - 1) / sb::kTile;
    }
    v->recs_cut = p.u.cut_k > 0.0f;
    v->sorted_pending = false;  // a new visible set replaces whatever a previous frame left behind
    if (v->timing) SB_CUDA(v->ctx, cudaEventRecord(v->ev[0], stream));
    SB_CUDA(v->ctx, sb::launch_preprocess(v->sh_fmt, v->cov_fmt, p, v->pre_scratch.p, v->pre_scratch.bytes, v->ctx->num_sms, stream));
    if (v->timing) SB_CUDA(v->ctx, cudaEventRecord(v->ev[1], stream));
    return SB_OK;
}

sb::SortScratch sort_scratch(SbViewer* v) {
    sb::SortScratch s;
    s.keys_alt = v->sort_keys_alt.as<uint32_t>();
    s.payload_alt = v->sort_vals_alt.as<uint32_t>();
    s.internal = v->sort_internal.as<uint32_t>();
    s.internal_bytes = v->sort_internal.bytes;
    return s;
}

sb::SortScratch depth_sort_scratch(SbViewer* v) {
    sb::SortScratch s = sort_scratch(v);
    s.keys_alt = v->depth_keys_alt.as<uint32_t>();
    s.payload_alt = v->depth_vals_alt.as<uint32_t>();
    return s;
}

SbStatus do_sort(SbViewer* v, cudaStream_t stream, bool defer_copy_home = false) {
    // keys = f32 depth bit patterns in [0, 0x3F800000]; pads (2.0) beyond V are left in place.  The digit plan comes from the
    // bits K1 saw varying among the visible keys (SB_DEPTH_SORT=v3 selects the fixed 4 x 8-bit round-1 sort for A/B runs).
    static const bool legacy = [] { const char* c = std::getenv("SB_DEPTH_SORT"); return c && std::string(c) == "v3"; }();
    if (legacy) {
        // size hint = the visible count the previous preprocess of this viewer reported (0 = none yet): a heuristic only
        SB_CUDA(v->ctx, sb::launch_sort(v->keys.as<uint32_t>(), v->indices.as<uint32_t>(), v->d_visible(), v->n, 0, 32, depth_sort_scratch(v),
                                        v->ctx->num_sms, stream, defer_copy_home ? v->d_sort_parity() : nullptr, v->h_needed[4]));
    } else {
        SB_CUDA(v->ctx, sb::launch_sort_adaptive(v->keys.as<uint32_t>(), v->indices.as<uint32_t>(), v->d_visible(), v->n,
                                                 v->pre_scratch.as<uint32_t>(), v->sort_prep_fresh, -1, 0, depth_sort_scratch(v),
                                                 v->ctx->num_sms, stream, v->d_sort_parity()));
        v->sort_prep_fresh = false;
        if (!defer_copy_home)
            SB_CUDA(v->ctx, sb::launch_sort_finish(v->keys.as<uint32_t>(), v->indices.as<uint32_t>(), depth_sort_scratch(v), v->d_visible(),
                                                   v->n, v->d_sort_parity(), v->ctx->num_sms, stream));
    }
    v->sorted_pending = defer_copy_home;
    v->last_stream = stream;
    if (v->timing) SB_CUDA(v->ctx, cudaEventRecord(v->ev[2], stream));
    return SB_OK;
}

// brings a deferred sort result home to indices/keys (no-op when it already is)
SbStatus flush_sorted(SbViewer* v, cudaStream_t stream) {
    if (!v->sorted_pending) return SB_OK;
    SB_CUDA(v->ctx, sb::launch_sort_finish(v->keys.as<uint32_t>(), v->indices.as<uint32_t>(), depth_sort_scratch(v), v->d_visible(), v->n,
                                           v->d_sort_parity(), v->ctx->num_sms, stream));
    v->sorted_pending = false;
    return SB_OK;
}

SbStatus do_draw(SbViewer* v, const SbCameraPod& cam, const SbGaussianTransformPod& gt, const SbTarget* target, int clear,
                 cudaStream_t stream, const SbDepthAttachment* depth = nullptr, cudaStream_t raster_stream = nullptr,
                 cudaEvent_t bin_done = nullptr) {
    sb::RasterParams p;
    std::memset(&p, 0, sizeof p);
    p.split_raster = bin_done != nullptr;  // multi-model frames: binning on `stream`, the raster on the frame's draw-order stream
    p.raster_stream = raster_stream;       // (may be the legacy default stream, i.e. null)
    p.bin_done = bin_done;
    // a reallocation below must wait for everything that may still read the old buffers, on either stream
    auto quiesce = [&]() { return bin_done ? cudaDeviceSynchronize() : cudaStreamSynchronize(stream); };
    p.u = make_uniforms(cam, v->model_transform, gt, v->target_format);
    SbStatus s = check_target(v, target, p.u);
    if (s != SB_OK) return s;
    const uint32_t tiles = p.u.tiles_x * p.u.tiles_y;
    // Capacity policy: a previous frame reported (through a mapped pinned word, no
    // sync) how many (splat,tile) duplicates it needed; grow before enqueuing this frame.  The
    // frame that overflowed dropped its nearest splats and is flagged by read_frame_stats.
    const uint64_t want = std::max<uint64_t>(8ull * v->n + 128ull * tiles + (1ull << 20), (uint64_t)*v->h_needed * 3 / 2);
    if (*v->h_needed > v->dup_capacity || (v->tile_capacity == 0 && want > v->dup_capacity) || tiles > v->tile_capacity) {
        SB_CUDA(v->ctx, quiesce());
        if (want > v->dup_capacity) {
            SbStatus gs = viewer_reserve(v, want);
            if (gs != SB_OK) return gs;
        }
        *v->h_needed = 0;
    }
    if (tiles > v->tile_capacity) {
        SB_CUDA(v->ctx, quiesce());
        SB_CUDA(v->ctx, v->tile_ranges.alloc((size_t)tiles * 8));
        SB_CUDA(v->ctx, v->tile_order.alloc((size_t)tiles * 4));
        v->tile_capacity = tiles;
    }
    p.recs = v->recs.as<sb::SplatRec>();
    p.tboxes = v->tboxes.as<sb::TileBox>();
    p.sorted_indices = v->ext_indices ? v->ext_indices : v->indices.as<uint32_t>();
    if (v->sorted_pending && !v->ext_indices) {
        if (depth && depth->d_depth) {  // the depth-tested pass reads the order on the host-visible side too: bring it home
            SbStatus fs = flush_sorted(v, stream);
            if (fs != SB_OK) return fs;
        } else {
            p.sorted_indices_alt = v->depth_vals_alt.as<uint32_t>();
            p.sort_parity = v->d_sort_parity();
        }
    }
    p.visible_count = v->ext_count ? v->ext_count : v->d_visible();
    p.max_visible = v->n;
    p.buf.dup_offsets = v->dup_offsets.as<uint32_t>();
    p.buf.tboxes_sorted = v->tboxes_sorted.as<sb::TileBox>();
    p.buf.win_first = v->win_first.as<uint32_t>();
    p.buf.win_capacity = (uint32_t)(v->win_first.bytes / 4);
    p.buf.dup_keys = v->dup_keys.as<uint32_t>();
    p.buf.dup_vals = v->dup_vals.as<uint32_t>();
    p.buf.tile_ranges = v->tile_ranges.as<uint32_t>();
    {
        static const bool row_major = [] { const char* c = std::getenv("SB_RASTER_ORDER"); return c && std::string(c) == "rows"; }();
        p.buf.tile_order = row_major ? nullptr : v->tile_order.as<uint32_t>();
    }
    p.buf.dup_count = v->d_dup_count();
    p.buf.overflow = v->d_overflow();
    p.buf.needed_host = v->d_needed;
    p.buf.scan_counter = v->d_scan_counter();
    p.buf.scan_status = v->d_scan_status();
    p.buf.tile_recs = v->tile_recs.as<sb::SplatRec>();
    p.buf.dup_capacity = v->dup_capacity;
    p.sort = sort_scratch(v);
    p.target = *target;
    p.strict_exp = v->strict_exp;
    p.no_discard = v->exact_cutoff ? 1 : 0;
    p.cut_k = 0.0f;  // set below, once it is known whether the records keep their cut extents for this pass
    p.clear = clear;
    {
        // SB_RASTER_CULL=bbox keeps the warp-level cull on the alive-region bbox only (A/B measurements)
        static const bool bbox_only = [] { const char* c = std::getenv("SB_RASTER_CULL"); return c && std::string(c) == "bbox"; }();
        p.obb_cull = bbox_only ? 0 : 1;
    }
    {
        static const int bands = [] { const char* c = std::getenv("SB_RASTER_BANDS"); return c ? std::atoi(c) : 1; }();
        p.raster_bands = bands;
    }
    p.events = v->timing ? &v->ev[3] : nullptr;
    p.recs_map = v->use_gather4 ? &v->recs_map : nullptr;
    if (depth && depth->d_depth) {
        if (!v->use_gather4) return fail(v->ctx, SB_ERR_INVALID_ARG, "a depth attachment needs the default (TMA gather4) raster path");
        if (depth->compare < SB_COMPARE_NEVER || depth->compare > SB_COMPARE_ALWAYS) return fail(v->ctx, SB_ERR_INVALID_ARG, "bad compare function");
        if (depth->pitch_bytes < target->width * 4u) return fail(v->ctx, SB_ERR_BAD_BUFFER_SIZE, "depth pitch too small");
        p.depth = static_cast<float*>(depth->d_depth);
        p.depth_pitch = depth->pitch_bytes;
        p.depth_compare = depth->compare;
        p.depth_write = depth->write_enabled != 0;
        p.pods = static_cast<const uint8_t*>(v->d_gaussians);
        p.pod_stride = v->stride;
        if (v->recs_cut) {
            // a fragment the cut-off would drop still takes part in the depth test and may write depth: redo the vertex
            // stage of the visible splats with their full extents
            sb::PreParams vp;
            std::memset(&vp, 0, sizeof vp);
            vp.gaussians = p.pods;
            vp.n = v->n;
            vp.recs = v->recs.as<sb::SplatRec>();
            vp.tboxes = v->tboxes.as<sb::TileBox>();
            vp.u = p.u;  // cut_k = 0
            SB_CUDA(v->ctx, sb::launch_vertex_stage(v->sh_fmt, v->cov_fmt, vp, p.sorted_indices, p.visible_count, v->ctx->num_sms, stream));
            v->recs_cut = false;
        }
    }
    p.cut_k = v->recs_cut ? 1.0f / sb::kAlphaCut : 0.0f;
    p.counters = v->counting ? v->counters.as<unsigned long long>() : nullptr;
    if (v->counting && clear) SB_CUDA(v->ctx, cudaMemsetAsync(v->counters.p, 0, 32, stream));
    SB_CUDA(v->ctx, sb::launch_bin_and_raster(p, v->ctx->num_sms, stream));
    v->last_tiles_x = p.u.tiles_x;
    v->last_tiles_y = p.u.tiles_y;
    return SB_OK;
}

}  // namespace

extern "C" {

const char* sb_last_error_string(const SbContext* ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

SbStatus sb_ctx_create(int32_t device_ordinal, SbContext** out) {
    if (!out) return fail(nullptr, SB_ERR_INVALID_ARG, "null out");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(nullptr, SB_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device_ordinal < 0 || device_ordinal >= count) return fail(nullptr, SB_ERR_INVALID_ARG, "bad device ordinal");
    cudaDeviceProp prop;
    SB_CUDA(nullptr, cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major != 10) return fail(nullptr, SB_ERR_CUDA, "splat_b200 is built for sm_100a (Blackwell B200) only");
    SB_CUDA(nullptr, cudaSetDevice(device_ordinal));
    SbContext* ctx = new SbContext();
    ctx->device = device_ordinal;
    ctx->num_sms = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    SB_CUDA(nullptr, cudaMemGetInfo(&free_b, &total_b));
    ctx->model_size_limit = free_b;
    *out = ctx;
    return SB_OK;
}

void sb_ctx_destroy(SbContext* ctx) { delete ctx; }

SbStatus sb_shared_frame_create(SbContext* ctx, uint64_t bytes, void** d_frame, uint8_t handle[SB_SHARED_HANDLE_BYTES]) {
    if (!ctx || !d_frame || !handle || bytes == 0) return fail(ctx, SB_ERR_INVALID_ARG, "null / empty shared frame");
    static_assert(sizeof(cudaIpcMemHandle_t) == SB_SHARED_HANDLE_BYTES, "CUDA IPC handle size");
    DeviceGuard device_guard(ctx);
    void* p = nullptr;
    SB_CUDA(ctx, cudaMalloc(&p, bytes));  // its own allocation: an IPC handle names a whole cudaMalloc block
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail_cuda(ctx, e, "cudaIpcGetMemHandle");
    }
    std::memcpy(handle, &h, sizeof h);
    *d_frame = p;
    return SB_OK;
}

SbStatus sb_shared_frame_open(SbContext* ctx, const uint8_t handle[SB_SHARED_HANDLE_BYTES], void** d_frame) {
    if (!ctx || !d_frame || !handle) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(ctx);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    SB_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_frame = p;
    return SB_OK;
}

SbStatus sb_shared_frame_close(SbContext* ctx, void* d_frame) {
    if (!ctx || !d_frame) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(ctx);
    SB_CUDA(ctx, cudaIpcCloseMemHandle(d_frame));
    return SB_OK;
}

SbStatus sb_shared_frame_destroy(SbContext* ctx, void* d_frame) {
    if (!ctx || !d_frame) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(ctx);
    SB_CUDA(ctx, cudaFree(d_frame));
    return SB_OK;
}

SbStatus sb_probe_peaks(SbContext* ctx, void* stream, double* fp32_lane_ops_per_s, double* smem_bytes_per_s) {
    if (!ctx || !fp32_lane_ops_per_s || !smem_bytes_per_s) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(ctx);
    SB_CUDA(ctx, sb::probe_fp32_peak(ctx->num_sms, static_cast<cudaStream_t>(stream), fp32_lane_ops_per_s));
    SB_CUDA(ctx, sb::probe_smem_peak(ctx->num_sms, static_cast<cudaStream_t>(stream), smem_bytes_per_s));
    return SB_OK;
}

SbStatus sb_ctx_set_model_size_limit(SbContext* ctx, uint64_t bytes) {
    if (!ctx) return fail(nullptr, SB_ERR_INVALID_ARG, "null context");
    ctx->model_size_limit = bytes;
    return SB_OK;
}

SbStatus sb_viewer_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format, const void* packed_pods, uint64_t n,
                          SbViewer** out) {
    if (n && !packed_pods) return fail(ctx, SB_ERR_INVALID_ARG, "null pods");
    DeviceGuard device_guard(ctx);
    SbViewer* v = nullptr;
    SbStatus s = viewer_new(ctx, sh_fmt, cov_fmt, target_format, n, &v);
    if (s != SB_OK) return s;
    cudaError_t e = v->gaussians_owned.alloc((size_t)n * v->stride);
    if (e == cudaSuccess && n) e = cudaMemcpy(v->gaussians_owned.p, packed_pods, (size_t)n * v->stride, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        sb_viewer_destroy(v);
        return fail_cuda(ctx, e, "upload gaussians");
    }
    v->d_gaussians = v->gaussians_owned.p;
    *out = v;
    return SB_OK;
}

SbStatus sb_viewer_create_from_gaussians(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format, const SbGaussian* src,
                                         uint64_t n, SbViewer** out) {
    const uint32_t stride = sb_pod_stride(sh_fmt, cov_fmt);
    if (stride == 0) return fail(ctx, SB_ERR_INVALID_ARG, "unknown pod format");
    std::vector<uint8_t> pods((size_t)n * stride + 16);
    SbStatus s = sb_pack_gaussians(src, n, sh_fmt, cov_fmt, pods.data());
    if (s != SB_OK) return fail(ctx, s, "pack failed");
    return sb_viewer_create(ctx, sh_fmt, cov_fmt, target_format, pods.data(), n, out);
}

SbStatus sb_viewer_create_from_device(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format, const void* d_pods,
                                      uint64_t d_bytes, uint64_t n, SbViewer** out) {
    const uint32_t stride = sb_pod_stride(sh_fmt, cov_fmt);
    if (stride == 0) return fail(ctx, SB_ERR_INVALID_ARG, "unknown pod format");
    if (d_bytes != n * stride) return fail(ctx, SB_ERR_BAD_BUFFER_SIZE, "gaussians buffer size != n * size_of::<G>()");
    if (n && (!d_pods || (reinterpret_cast<uintptr_t>(d_pods) & 15u))) return fail(ctx, SB_ERR_INVALID_ARG, "device pods must be 16-byte aligned");
    DeviceGuard device_guard(ctx);
    SbViewer* v = nullptr;
    SbStatus s = viewer_new(ctx, sh_fmt, cov_fmt, target_format, n, &v);
    if (s != SB_OK) return s;
    v->d_gaussians = d_pods;
    *out = v;
    return SB_OK;
}

void sb_viewer_destroy(SbViewer* v) {
    if (!v) return;
    DeviceGuard device_guard(v->ctx);
    if (v->twin) sb_viewer_destroy(v->twin);
    for (cudaStream_t s : v->bstream)
        if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : v->bevent)
        if (e) cudaEventDestroy(e);
    for (DeviceBuf* b : {&v->gaussians_owned, &v->indices, &v->keys, &v->args, &v->recs, &v->tboxes, &v->pre_scratch, &v->sort_keys_alt,
                         &v->sort_vals_alt, &v->sort_internal, &v->depth_keys_alt, &v->depth_vals_alt, &v->dup_offsets, &v->tboxes_sorted, &v->win_first, &v->dup_keys, &v->dup_vals, &v->tile_recs,
                         &v->tile_ranges, &v->tile_order, &v->bin_state, &v->selection, &v->orig_colors, &v->internal_target, &v->counters})
        b->release();
    if (v->h_needed) cudaFreeHost(const_cast<uint32_t*>(v->h_needed));
    for (cudaEvent_t e : v->ev)
        if (e) cudaEventDestroy(e);
    delete v;
}

SbStatus sb_viewer_update_camera_with_pod(SbViewer* v, const SbCameraPod* pod) {
    if (!v || !pod) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    v->camera = *pod;
    return SB_OK;
}

SbStatus sb_viewer_update_camera(SbViewer* v, const float pos[3], float yaw, float pitch, float z_near, float z_far,
                                 float vertical_fov, uint32_t width, uint32_t height) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    return sb_camera_pod(pos, yaw, pitch, z_near, z_far, vertical_fov, width, height, &v->camera);
}

SbStatus sb_viewer_update_model_transform(SbViewer* v, const float pos[3], const float rot[4], const float scale[3]) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    return sb_model_transform_pod(pos, rot, scale, &v->model_transform);
}

SbStatus sb_viewer_update_model_transform_with_pod(SbViewer* v, const SbModelTransformPod* pod) {
    if (!v || !pod) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    v->model_transform = *pod;
    return SB_OK;
}

SbStatus sb_viewer_update_gaussian_transform(SbViewer* v, float size, int32_t display_mode, int32_t sh_deg, int32_t no_sh0,
                                             float max_std_dev) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    SbGaussianTransformPod pod;
    SbStatus s = sb_gaussian_transform_pod(size, display_mode, sh_deg, no_sh0, max_std_dev, &pod);
    if (s != SB_OK) return fail(v->ctx, s, "bad gaussian transform");
    v->gaussian_transform = pod;
    return SB_OK;
}

SbStatus sb_viewer_update_gaussian_transform_with_pod(SbViewer* v, const SbGaussianTransformPod* pod) {
    if (!v || !pod) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    if (pod->display_mode > 2 || pod->sh_deg > 3) return fail(v->ctx, SB_ERR_INVALID_ARG, "bad gaussian transform pod");
    v->gaussian_transform = *pod;
    return SB_OK;
}

SbStatus sb_viewer_enable_selection(SbViewer* v, int32_t enabled) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    v->selection_enabled = enabled != 0;
    return SB_OK;
}

SbStatus sb_viewer_selection_ptr(SbViewer* v, uint32_t** d_words, uint64_t* n_words) {
    if (!v || !d_words || !n_words) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    *d_words = v->selection.as<uint32_t>();
    *n_words = ((uint64_t)v->n + 31) / 32;
    return SB_OK;
}

SbStatus sb_viewer_set_selection(SbViewer* v, void* stream, const uint32_t* words, uint64_t n_words) {
    if (!v || (!words && n_words)) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (n_words != ((uint64_t)v->n + 31) / 32) return fail(v->ctx, SB_ERR_BAD_BUFFER_SIZE, "selection must hold ceil(n/32) words");
    SB_CUDA(v->ctx, cudaMemcpyAsync(v->selection.p, words, n_words * 4, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    return SB_OK;
}

SbStatus sb_viewer_read_selection(SbViewer* v, void* stream, uint32_t* out, uint64_t n_words) {
    if (!v || (!out && n_words)) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (n_words != ((uint64_t)v->n + 31) / 32) return fail(v->ctx, SB_ERR_BAD_BUFFER_SIZE, "selection must hold ceil(n/32) words");
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    SB_CUDA(v->ctx, cudaMemcpy(out, v->selection.p, n_words * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}

SbStatus sb_viewer_set_invert_selection(SbViewer* v, int32_t invert) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    v->invert_selection = invert ? 1u : 0u;
    return SB_OK;
}

SbStatus sb_viewer_select_rect(SbViewer* v, void* stream, float x0, float y0, float x1, float y1) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    const sb::Uniforms u = make_uniforms(v->camera, v->model_transform, v->gaussian_transform, v->target_format);
    SB_CUDA(v->ctx, sb::launch_select_rect(static_cast<const uint8_t*>(v->d_gaussians), v->n, v->stride, u, x0, y0, x1, y1,
                                           v->selection.as<uint32_t>(), static_cast<cudaStream_t>(stream)));
    return SB_OK;
}

SbStatus sb_viewer_select_brush(SbViewer* v, void* stream, const float* points_xy, uint32_t n_points, float radius, int32_t accumulate) {
    if (!v || !points_xy) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (n_points == 0 || n_points > SB_BRUSH_MAX_POINTS || !(radius >= 0.0f))
        return fail(v->ctx, SB_ERR_INVALID_ARG, "brush stroke needs 1..SB_BRUSH_MAX_POINTS points and a non-negative radius");
    const sb::Uniforms u = make_uniforms(v->camera, v->model_transform, v->gaussian_transform, v->target_format);
    SB_CUDA(v->ctx, sb::launch_select_brush(static_cast<const uint8_t*>(v->d_gaussians), v->n, v->stride, u, points_xy, n_points, radius,
                                            accumulate, v->selection.as<uint32_t>(), static_cast<cudaStream_t>(stream)));
    return SB_OK;
}

SbStatus sb_viewer_apply_rgb_override(SbViewer* v, void* stream, const float rgb[3], float alpha) {
    if (!v || !rgb) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* pods = const_cast<uint8_t*>(static_cast<const uint8_t*>(v->d_gaussians));  // the reference edits viewer.gaussians_buffer in place
    if (!v->orig_colors.p) {
        SB_CUDA(v->ctx, v->orig_colors.alloc((size_t)(v->n ? v->n : 1) * 4));
        SB_CUDA(v->ctx, sb::launch_snapshot_colors(pods, v->n, v->stride, v->orig_colors.as<uint32_t>(), st));
    }
    SB_CUDA(v->ctx, sb::launch_rgb_override(pods, v->n, v->stride, v->orig_colors.as<uint32_t>(), v->selection.as<uint32_t>(), rgb, alpha, st));
    return SB_OK;
}

SbStatus sb_viewer_apply_basic_color_modifiers(SbViewer* v, void* stream, const SbBasicColorModifiers* m) {
    if (!v || !m) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    if (!(m->gamma > 0.0f)) return fail(v->ctx, SB_ERR_INVALID_ARG, "gamma must be positive");
    DeviceGuard device_guard(v->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* pods = const_cast<uint8_t*>(static_cast<const uint8_t*>(v->d_gaussians));
    if (!v->orig_colors.p) {  // NonDestructiveModifier: every edit starts from the source colours
        SB_CUDA(v->ctx, v->orig_colors.alloc((size_t)(v->n ? v->n : 1) * 4));
        SB_CUDA(v->ctx, sb::launch_snapshot_colors(pods, v->n, v->stride, v->orig_colors.as<uint32_t>(), st));
    }
    SB_CUDA(v->ctx, sb::launch_basic_color_modifiers(pods, v->n, v->stride, v->orig_colors.as<uint32_t>(), v->selection.as<uint32_t>(),
                                                     m->rgb_override, m->rgb_or_hsv, m->alpha, m->contrast, m->exposure, m->gamma, st));
    return SB_OK;
}

SbStatus sb_viewer_restore_gaussians(SbViewer* v, void* stream) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (!v->orig_colors.p) return SB_OK;  // never edited
    uint8_t* pods = const_cast<uint8_t*>(static_cast<const uint8_t*>(v->d_gaussians));
    const float rgb[3] = {0.0f, 0.0f, 0.0f};
    SB_CUDA(v->ctx, sb::launch_rgb_override(pods, v->n, v->stride, v->orig_colors.as<uint32_t>(), nullptr, rgb, 1.0f, static_cast<cudaStream_t>(stream)));
    return SB_OK;
}

SbStatus sb_viewer_render_with_pass(SbViewer* v, void* stream, const SbTarget* target, const SbDepthAttachment* depth, int32_t load,
                                    int32_t run_stages) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        SbStatus os = report_overflow(v);
        if (os != SB_OK) return os;
    }
    if (run_stages) {
        SbStatus s = do_preprocess(v, v->camera, v->gaussian_transform, st);
        if (s == SB_OK) s = do_sort(v, st, true);
        if (s != SB_OK) return s;
    }
    return do_draw(v, v->camera, v->gaussian_transform, target, load ? 0 : 1, st, depth);
}

SbStatus sb_viewer_preprocess(SbViewer* v, void* stream) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    return do_preprocess(v, v->camera, v->gaussian_transform, static_cast<cudaStream_t>(stream));
}

SbStatus sb_viewer_sort(SbViewer* v, void* stream) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    SbStatus fs = flush_sorted(v, static_cast<cudaStream_t>(stream));  // a previous frame's result, if it was left in the alt buffers
    if (fs != SB_OK) return fs;
    return do_sort(v, static_cast<cudaStream_t>(stream));
}

SbStatus sb_viewer_draw(SbViewer* v, void* stream, const SbTarget* target) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    SbStatus os = report_overflow(v);
    if (os != SB_OK) return os;
    return do_draw(v, v->camera, v->gaussian_transform, target, 1, static_cast<cudaStream_t>(stream));
}

SbStatus sb_viewer_render(SbViewer* v, void* stream, const SbTarget* target) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // validate before enqueuing anything
    sb::Uniforms u = make_uniforms(v->camera, v->model_transform, v->gaussian_transform, v->target_format);
    SbStatus s = check_target(v, target, u);
    if (s == SB_OK) s = report_overflow(v);
    if (s != SB_OK) return s;
    s = do_preprocess(v, v->camera, v->gaussian_transform, st, v->strip_cull ? target : nullptr);
    if (s != SB_OK) return s;
    s = do_sort(v, st, true);
    if (s != SB_OK) return s;
    return do_draw(v, v->camera, v->gaussian_transform, target, 1, st);
}

SbStatus sb_viewer_render_to_host(SbViewer* v, void* stream, const SbCameraPod* cam, void* host_pixels, uint64_t host_bytes) {
    if (!v || !cam || !host_pixels) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    const uint32_t w = (uint32_t)cam->size[0], h = (uint32_t)cam->size[1];
    const uint64_t need = (uint64_t)w * h * bytes_per_pixel(v->target_format);
    if (host_bytes != need) return fail(v->ctx, SB_ERR_BAD_BUFFER_SIZE, "host frame buffer size mismatch");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (v->internal_target.bytes < need) {
        SB_CUDA(v->ctx, cudaStreamSynchronize(st));
        SB_CUDA(v->ctx, v->internal_target.alloc(need));
    }
    v->camera = *cam;
    SbTarget t;
    t.d_pixels = v->internal_target.p;
    t.pitch_bytes = w * bytes_per_pixel(v->target_format);
    t.width = w;
    t.height = h;
    t.format = v->target_format;
    t.row0 = 0;
    t.rows = 0;
    SbStatus s = sb_viewer_render(v, stream, &t);
    if (s != SB_OK) return s;
    SB_CUDA(v->ctx, cudaMemcpyAsync(host_pixels, v->internal_target.p, need, cudaMemcpyDeviceToHost, st));
    return SB_OK;
}

namespace {

SbStatus batch_setup(SbViewer* v) {
    if (v->twin) return SB_OK;
    SbViewer* t = nullptr;
    SbStatus s = sb_viewer_create_from_device(v->ctx, v->sh_fmt, v->cov_fmt, v->target_format, v->d_gaussians, (uint64_t)v->n * v->stride,
                                              v->n, &t);
    if (s != SB_OK) return s;
    for (cudaStream_t& st : v->bstream) SB_CUDA(v->ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (cudaEvent_t& e : v->bevent) SB_CUDA(v->ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    v->twin = t;
    return SB_OK;
}

// copies the per-frame state the reference keeps in uniform buffers / the selection buffer
void sync_twin(SbViewer* v) {
    SbViewer* t = v->twin;
    t->model_transform = v->model_transform;
    t->gaussian_transform = v->gaussian_transform;
    t->selection_enabled = v->selection_enabled;
    t->invert_selection = v->invert_selection;
    t->selection_override = v->selection.as<uint32_t>();
    t->strict_exp = v->strict_exp;
    t->exact_cutoff = v->exact_cutoff;
}

}  // namespace

SbStatus sb_viewer_render_batch(SbViewer* v, void* stream, const SbCameraPod* cams, const SbTarget* targets, void* const* host_pixels,
                                uint32_t count) {
    if (!v || (count && (!cams || (!targets && !host_pixels)))) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (count == 0) return SB_OK;
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    SbStatus s = batch_setup(v);
    if (s != SB_OK) return s;
    sync_twin(v);
    // validate every target before enqueuing anything
    const uint32_t bpp = bytes_per_pixel(v->target_format);
    std::vector<SbTarget> tg(count);
    // slots alternate so that the LAST view always lands in the primary viewer: after the batch its buffers (records, indices,
    // keys, indirect args) hold the view v->camera names, as the reference's would after its last update_camera + render
    auto slot_of = [count](uint32_t i) { return (count - 1u - i) & 1u; };
    for (uint32_t i = 0; i < count; i++) {
        SbViewer* slot = slot_of(i) ? v->twin : v;
        const uint32_t w = (uint32_t)cams[i].size[0], h = (uint32_t)cams[i].size[1];
        if (targets) {
            tg[i] = targets[i];
        } else {  // render into the slot's internal target, then copy to host_pixels[i]
            const uint64_t need = (uint64_t)w * h * bpp;
            if (!host_pixels[i]) return fail(v->ctx, SB_ERR_INVALID_ARG, "null host frame");
            if (slot->internal_target.bytes < need) {
                SB_CUDA(v->ctx, cudaDeviceSynchronize());
                SB_CUDA(v->ctx, slot->internal_target.alloc(need));
            }
            tg[i] = SbTarget{slot->internal_target.p, w * bpp, w, h, v->target_format, 0, 0};
        }
        const sb::Uniforms u = make_uniforms(cams[i], v->model_transform, v->gaussian_transform, v->target_format);
        s = check_target(v, &tg[i], u);
        if (s != SB_OK) return s;
    }
    s = report_overflow(v);
    if (s == SB_OK) s = report_overflow(v->twin);
    if (s != SB_OK) return s;
    // fork: both internal streams start after everything already enqueued on the caller's stream
    SB_CUDA(v->ctx, cudaEventRecord(v->bevent[2], user));
    for (int k = 0; k < 2; k++) SB_CUDA(v->ctx, cudaStreamWaitEvent(v->bstream[k], v->bevent[2], 0));
    for (uint32_t i = 0; i < count; i++) {
        SbViewer* slot = slot_of(i) ? v->twin : v;
        cudaStream_t st = v->bstream[slot_of(i)];
        s = do_preprocess(slot, cams[i], v->gaussian_transform, st);
        if (s == SB_OK) s = do_sort(slot, st, true);
        if (s == SB_OK) s = do_draw(slot, cams[i], v->gaussian_transform, &tg[i], 1, st);
        if (s != SB_OK) return s;
        if (!targets) {
            const uint64_t need = (uint64_t)tg[i].width * tg[i].height * bpp;
            SB_CUDA(v->ctx, cudaMemcpyAsync(host_pixels[i], tg[i].d_pixels, need, cudaMemcpyDeviceToHost, st));
        }
    }
    // join: the caller's stream continues once both slots are done
    for (int k = 0; k < 2; k++) {
        SB_CUDA(v->ctx, cudaEventRecord(v->bevent[k], v->bstream[k]));
        SB_CUDA(v->ctx, cudaStreamWaitEvent(user, v->bevent[k], 0));
    }
    v->camera = cams[count - 1];
    v->last_stream = user;  // ordered after both slots by the join
    v->twin->last_stream = user;
    return SB_OK;
}

SbStatus sb_viewer_gaussians_ptr(SbViewer* v, const void** d_pods, uint64_t* bytes) {
    if (!v || !d_pods || !bytes) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    *d_pods = v->d_gaussians;
    *bytes = (uint64_t)v->n * v->stride;
    return SB_OK;
}
SbStatus sb_viewer_indirect_args_ptr(SbViewer* v, const SbDrawIndirectArgs** d_args) {
    if (!v || !d_args) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    *d_args = v->d_draw();
    return SB_OK;
}
SbStatus sb_viewer_radix_sort_indirect_args_ptr(SbViewer* v, const SbDispatchIndirectArgs** d_args) {
    if (!v || !d_args) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    *d_args = v->d_dispatch();
    return SB_OK;
}
SbStatus sb_viewer_indirect_indices_ptr(SbViewer* v, const uint32_t** d_indices, uint64_t* count) {
    if (!v || !d_indices || !count) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    SbStatus fs = flush_sorted(v, v->last_stream);
    if (fs != SB_OK) return fs;
    *d_indices = v->indices.as<uint32_t>();
    *count = v->n;
    return SB_OK;
}
SbStatus sb_viewer_gaussians_depth_ptr(SbViewer* v, const float** d_keys, uint64_t* bytes) {
    if (!v || !d_keys || !bytes) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    SbStatus fs = flush_sorted(v, v->last_stream);
    if (fs != SB_OK) return fs;
    *d_keys = v->keys.as<float>();
    *bytes = (uint64_t)v->padded * 4;
    return SB_OK;
}

SbStatus sb_viewer_read_indirect_args(SbViewer* v, void* stream, SbDrawIndirectArgs* draw, SbDispatchIndirectArgs* dispatch) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    if (draw) SB_CUDA(v->ctx, cudaMemcpy(draw, v->d_draw(), sizeof *draw, cudaMemcpyDeviceToHost));
    if (dispatch) SB_CUDA(v->ctx, cudaMemcpy(dispatch, v->d_dispatch(), sizeof *dispatch, cudaMemcpyDeviceToHost));
    return SB_OK;
}
SbStatus sb_viewer_read_indices(SbViewer* v, void* stream, uint32_t* out, uint64_t count) {
    if (!v || (!out && count)) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (count > v->padded) return fail(v->ctx, SB_ERR_BAD_BUFFER_SIZE, "count exceeds buffer");
    SbStatus fs = flush_sorted(v, static_cast<cudaStream_t>(stream));
    if (fs != SB_OK) return fs;
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    SB_CUDA(v->ctx, cudaMemcpy(out, v->indices.p, count * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}
SbStatus sb_viewer_read_depth_keys(SbViewer* v, void* stream, float* out, uint64_t count) {
    if (!v || (!out && count)) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (count > v->padded) return fail(v->ctx, SB_ERR_BAD_BUFFER_SIZE, "count exceeds buffer");
    SbStatus fs = flush_sorted(v, static_cast<cudaStream_t>(stream));
    if (fs != SB_OK) return fs;
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    SB_CUDA(v->ctx, cudaMemcpy(out, v->keys.p, count * 4, cudaMemcpyDeviceToHost));
    return SB_OK;
}
SbStatus sb_viewer_read_frame_stats(SbViewer* v, void* stream, uint64_t* visible, uint64_t* duplicates, uint32_t* overflowed) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    uint32_t vis = 0, st[2] = {0, 0};
    SB_CUDA(v->ctx, cudaMemcpy(&vis, v->d_visible(), 4, cudaMemcpyDeviceToHost));
    SB_CUDA(v->ctx, cudaMemcpy(st, v->bin_state.p, 8, cudaMemcpyDeviceToHost));
    if (visible) *visible = vis;
    if (duplicates) *duplicates = st[0];
    if (overflowed) *overflowed = st[1];
    return st[1] ? fail(v->ctx, SB_ERR_OVERFLOW, "tile-duplicate capacity exceeded; call sb_viewer_reserve_duplicates") : SB_OK;
}

SbStatus sb_viewer_raster_path(SbViewer* v, int32_t* tma_gather4) {
    if (!v || !tma_gather4) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    *tma_gather4 = v->use_gather4 ? 1 : 0;
    return SB_OK;
}

SbStatus sb_viewer_set_strict_exp(SbViewer* v, int32_t strict) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    v->strict_exp = strict != 0;
    return SB_OK;
}

SbStatus sb_viewer_set_exact_cutoff(SbViewer* v, int32_t enabled) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    v->exact_cutoff = enabled != 0;
    return SB_OK;
}

SbStatus sb_viewer_read_tile_row_work(SbViewer* v, void* stream, uint64_t* out, uint32_t n_rows) {
    if (!v || !out) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    if (n_rows != v->last_tiles_y || v->last_tiles_x == 0) return fail(v->ctx, SB_ERR_INVALID_ARG, "n_rows differs from the last frame's tile rows");
    DeviceGuard device_guard(v->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t tiles = (size_t)v->last_tiles_x * v->last_tiles_y;
    std::vector<uint32_t> ranges(tiles * 2);
    SB_CUDA(v->ctx, cudaMemcpyAsync(ranges.data(), v->tile_ranges.p, tiles * 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(v->ctx, cudaStreamSynchronize(st));
    for (uint32_t y = 0; y < n_rows; y++) {
        uint64_t sum = 0;
        for (uint32_t x = 0; x < v->last_tiles_x; x++) {
            const size_t t = (size_t)y * v->last_tiles_x + x;
            sum += ranges[2 * t + 1] - ranges[2 * t];
        }
        out[y] = sum;
    }
    return SB_OK;
}

struct SbStrips {
    SbViewer* v = nullptr;
    uint32_t world = 0, rank = 0;
    uint32_t row0s[sb::kMaxStripRanks] = {}, rows[sb::kMaxStripRanks] = {};
    uint32_t slice_lo = 0, slice_n = 0, segment_capacity = 0;
    DeviceBuf inbox;    // [world][segment_capacity] 64-byte parcels, then [world] counts (64-byte aligned)
    DeviceBuf outbox;   // the same shape, local: parcels packed for each destination before they are shipped in bulk
    DeviceBuf scratch;  // scatter tickets + look-back status
    size_t counts_offset = 0;
    void* peer_inbox[sb::kMaxStripRanks] = {};   // peer-mapped (own pointer at [rank])
    bool connected = false;
};

SbStatus sb_strips_create(SbViewer* v, uint32_t world, uint32_t rank, const uint32_t* row0s, const uint32_t* rows, SbStrips** out,
                          SbStripsExport* exported) {
    if (!v || !row0s || !rows || !out || !exported) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    if (world == 0 || world > (uint32_t)sb::kMaxStripRanks || rank >= world) return fail(v->ctx, SB_ERR_INVALID_ARG, "bad world / rank (at most 16 ranks)");
    DeviceGuard device_guard(v->ctx);
    SbStrips* s = new SbStrips();
    s->v = v;
    s->world = world;
    s->rank = rank;
    for (uint32_t r = 0; r < world; r++) {
        s->row0s[r] = row0s[r];
        s->rows[r] = rows[r];
    }
    // slices: ascending index ranges, boundaries at multiples of 32 so a selection word never straddles two ranks' K1
    auto cut = [&](uint32_t r) { return r >= world ? v->n : (uint32_t)(((uint64_t)v->n * r / world) & ~31ull); };
    s->slice_lo = cut(rank);
    s->slice_n = cut(rank + 1) - s->slice_lo;
    uint32_t cap = 0;
    for (uint32_t r = 0; r < world; r++) cap = std::max(cap, cut(r + 1) - cut(r));
    s->segment_capacity = (cap + 31u) & ~31u;
    s->counts_offset = ((size_t)world * s->segment_capacity * 64 + 63) & ~(size_t)63;
    cudaError_t e = s->inbox.alloc(s->counts_offset + 64);
    if (e == cudaSuccess) e = cudaMemset(s->inbox.p, 0, s->inbox.bytes);
    if (e == cudaSuccess) e = s->outbox.alloc(s->counts_offset + 64);
    if (e == cudaSuccess) e = cudaMemset(s->outbox.p, 0, s->outbox.bytes);
    if (e == cudaSuccess) e = s->scratch.alloc(sb::strip_scatter_scratch_bytes(s->segment_capacity, world));
    cudaIpcMemHandle_t h[3];
    std::memset(h, 0, sizeof h);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h[0], s->inbox.p);  // (handles 1 and 2 are reserved)
    if (e != cudaSuccess) {
        s->inbox.release();
        s->outbox.release();
        s->scratch.release();
        delete s;
        return fail_cuda(v->ctx, e, "strip exchange buffers");
    }
    for (int i = 0; i < 3; i++) std::memcpy(exported->handles[i], &h[i], sizeof h[i]);
    *out = s;
    return SB_OK;
}

SbStatus sb_strips_connect(SbStrips* s, const SbStripsExport* all) {
    if (!s || !all) return fail(s ? s->v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    if (s->connected) return fail(s->v->ctx, SB_ERR_INVALID_ARG, "already connected");
    DeviceGuard device_guard(s->v->ctx);
    for (uint32_t r = 0; r < s->world; r++) {
        if (r == s->rank) {
            s->peer_inbox[r] = s->inbox.p;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, all[r].handles[0], sizeof h);
        SB_CUDA(s->v->ctx, cudaIpcOpenMemHandle(&s->peer_inbox[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    s->connected = true;
    return SB_OK;
}

SbStatus sb_strips_scatter(SbStrips* s, void* stream) {
    if (!s) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    SbViewer* v = s->v;
    if (!s->connected) return fail(v->ctx, SB_ERR_INVALID_ARG, "sb_strips_connect first");
    DeviceGuard device_guard(v->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const sb::Uniforms u = make_uniforms(v->camera, v->model_transform, v->gaussian_transform, v->target_format);
    // the reference's full-frame cull, on this rank's slice of the model
    SbStatus ps = do_preprocess(v, v->camera, v->gaussian_transform, st, nullptr, s->slice_lo, s->slice_n);
    if (ps != SB_OK) return ps;
    sb::StripScatterParams p;
    std::memset(&p, 0, sizeof p);
    p.indices = v->indices.as<uint32_t>();
    p.keys = v->keys.as<float>();
    p.visible_count = v->d_visible();
    p.max_visible = s->slice_n;
    p.recs = v->recs.as<sb::SplatRec>();
    p.tboxes = v->tboxes.as<sb::TileBox>();
    p.world = s->world;
    p.rank = s->rank;
    p.segment_capacity = s->segment_capacity;
    for (uint32_t r = 0; r < s->world; r++) {
        if (s->rows[r] == 0 || s->row0s[r] >= u.height) {
            p.ty_lo[r] = 1;
            p.ty_hi[r] = 0;
        } else {
            p.ty_lo[r] = s->row0s[r] / sb::kTile;
            p.ty_hi[r] = (std::min(s->row0s[r] + s->rows[r], u.height) - 1) / sb::kTile;
        }
        p.peer_inbox[r] = static_cast<uint4*>(s->peer_inbox[r]);
        p.peer_inbox_counts[r] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(s->peer_inbox[r]) + s->counts_offset);
    }
    p.inbox = s->inbox.as<uint4>();
    p.inbox_counts = reinterpret_cast<uint32_t*>(s->inbox.as<uint8_t>() + s->counts_offset);
    p.outbox = s->outbox.as<uint4>();
    p.outbox_counts = reinterpret_cast<uint32_t*>(s->outbox.as<uint8_t>() + s->counts_offset);
    SB_CUDA(v->ctx, sb::launch_strip_scatter(p, s->scratch.p, s->scratch.bytes, st));
    return SB_OK;
}

SbStatus sb_strips_render(SbStrips* s, void* stream, const SbTarget* target) {
    if (!s || !target) return fail(s ? s->v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    SbViewer* v = s->v;
    DeviceGuard device_guard(v->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const sb::Uniforms u = make_uniforms(v->camera, v->model_transform, v->gaussian_transform, v->target_format);
    SbStatus cs = check_target(v, target, u);
    if (cs == SB_OK) cs = report_overflow(v);
    if (cs != SB_OK) return cs;
    // the depth sort's prep words start from zero again: or / nand must describe the strip's list, not the slice's
    SB_CUDA(v->ctx, cudaMemsetAsync(v->pre_scratch.p, 0, sb::preprocess_scratch_prefix_bytes(), st));
    sb::StripConcatParams c;
    std::memset(&c, 0, sizeof c);
    c.inbox = s->inbox.as<uint4>();
    c.inbox_counts = reinterpret_cast<const uint32_t*>(s->inbox.as<uint8_t>() + s->counts_offset);
    c.rank = s->rank;
    c.recs = v->recs.as<sb::SplatRec>();
    c.tboxes = v->tboxes.as<sb::TileBox>();
    c.world = s->world;
    c.segment_capacity = s->segment_capacity;
    c.max_visible = v->n;
    c.indices = v->indices.as<uint32_t>();
    c.keys = v->keys.as<float>();
    c.keys_capacity = v->padded;
    c.draw_args = v->d_draw();
    c.sort_args = v->d_dispatch();
    c.visible_count = v->d_visible();
    c.visible_host = v->d_needed + 4;
    c.sort_prep = v->pre_scratch.as<uint32_t>();
    SB_CUDA(v->ctx, sb::launch_strip_concat(c, v->ctx->num_sms, st));
    v->sort_prep_fresh = true;
    v->sorted_pending = false;
    if (v->timing) SB_CUDA(v->ctx, cudaEventRecord(v->ev[1], st));  // "preprocess" of the stage times ends after the concat
    SbStatus rs = do_sort(v, st, true);
    if (rs != SB_OK) return rs;
    return do_draw(v, v->camera, v->gaussian_transform, target, 1, st);
}

void sb_strips_destroy(SbStrips* s) {
    if (!s) return;
    DeviceGuard device_guard(s->v->ctx);
    cudaDeviceSynchronize();
    if (s->connected)
        for (uint32_t r = 0; r < s->world; r++) {
            if (r == s->rank) continue;
            cudaIpcCloseMemHandle(s->peer_inbox[r]);
        }
    s->inbox.release();
    s->outbox.release();
    s->scratch.release();
    delete s;
}

SbStatus sb_viewer_set_strip_cull(SbViewer* v, int32_t enabled) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    v->strip_cull = enabled != 0;
    return SB_OK;
}

SbStatus sb_viewer_set_raster_counting(SbViewer* v, int32_t enabled) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (enabled && !v->counters.p) SB_CUDA(v->ctx, v->counters.alloc(32));
    v->counting = enabled != 0;
    return SB_OK;
}

SbStatus sb_viewer_read_raster_counters(SbViewer* v, void* stream, uint64_t* alive, uint64_t* evaluated) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (!v->counters.p) return fail(v->ctx, SB_ERR_INVALID_ARG, "raster counting was never enabled");
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    unsigned long long c[2] = {0, 0};
    SB_CUDA(v->ctx, cudaMemcpy(c, v->counters.p, 16, cudaMemcpyDeviceToHost));
    if (alive) *alive = c[0];
    if (evaluated) *evaluated = c[1];
    return SB_OK;
}

SbStatus sb_viewer_read_raster_warp_counters(SbViewer* v, void* stream, uint64_t* warp_evals, uint64_t* warp_evals_alive) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (!v->counters.p) return fail(v->ctx, SB_ERR_INVALID_ARG, "raster counting was never enabled");
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    unsigned long long c[4] = {0, 0, 0, 0};
    SB_CUDA(v->ctx, cudaMemcpy(c, v->counters.p, 32, cudaMemcpyDeviceToHost));
    if (warp_evals) *warp_evals = c[2];
    if (warp_evals_alive) *warp_evals_alive = c[3];
    return SB_OK;
}

SbStatus sb_viewer_set_stage_timing(SbViewer* v, int32_t enabled) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (enabled && !v->ev[0])
        for (cudaEvent_t& e : v->ev) SB_CUDA(v->ctx, cudaEventCreate(&e));
    v->timing = enabled != 0;
    return SB_OK;
}

SbStatus sb_viewer_read_stage_times(SbViewer* v, void* stream, float ms[6]) {
    if (!v || !ms) return fail(v ? v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    if (!v->timing) return fail(v->ctx, SB_ERR_INVALID_ARG, "stage timing is not enabled");
    SB_CUDA(v->ctx, cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    for (int i = 0; i < 6; i++) SB_CUDA(v->ctx, cudaEventElapsedTime(&ms[i], v->ev[i], v->ev[i + 1]));
    return SB_OK;
}

SbStatus sb_viewer_reserve_duplicates(SbViewer* v, uint64_t capacity) {
    if (!v) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(v->ctx);
    SB_CUDA(v->ctx, cudaDeviceSynchronize());
    SbStatus s = viewer_reserve(v, capacity);
    if (s == SB_OK) SB_CUDA(v->ctx, cudaMemset(v->bin_state.p, 0, 8));
    *v->h_needed = 0;  // the caller's capacity stands until a frame reports that it needed more
    return s;
}

// ---------------------------------------------------------------- standalone sorter

struct SbRadixSorter {
    SbContext* ctx;
    uint32_t capacity;
    DeviceBuf keys_alt, vals_alt, internal;
};

SbStatus sb_sorter_create(SbContext* ctx, uint32_t capacity, SbRadixSorter** out) {
    if (!ctx || !out) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(ctx);
    SbRadixSorter* s = new SbRadixSorter();
    s->ctx = ctx;
    s->capacity = capacity;
    cudaError_t e = s->keys_alt.alloc((size_t)capacity * 4);
    if (e == cudaSuccess) e = s->vals_alt.alloc((size_t)capacity * 4);
    if (e == cudaSuccess) e = s->internal.alloc(sb::sort_internal_bytes(capacity));
    if (e != cudaSuccess) {
        sb_sorter_destroy(s);
        return fail_cuda(ctx, e, "sorter alloc");
    }
    *out = s;
    return SB_OK;
}

void sb_sorter_destroy(SbRadixSorter* s) {
    if (!s) return;
    DeviceGuard device_guard(s->ctx);
    s->keys_alt.release();
    s->vals_alt.release();
    s->internal.release();
    delete s;
}

SbStatus sb_sorter_sort(SbRadixSorter* s, void* stream, uint32_t* d_keys, uint32_t* d_payload, const uint32_t* d_count,
                        uint32_t max_count, int32_t begin_bit, int32_t end_bit) {
    if (!s || !d_keys || !d_payload || !d_count) return fail(s ? s->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(s->ctx);
    if (max_count > s->capacity) return fail(s->ctx, SB_ERR_BAD_BUFFER_SIZE, "max_count exceeds sorter capacity");
    if (begin_bit < 0 || end_bit > 32 || begin_bit >= end_bit) return fail(s->ctx, SB_ERR_INVALID_ARG, "bad bit range");
    sb::SortScratch sc;
    sc.keys_alt = s->keys_alt.as<uint32_t>();
    sc.payload_alt = s->vals_alt.as<uint32_t>();
    sc.internal = s->internal.as<uint32_t>();
    sc.internal_bytes = s->internal.bytes;
    SB_CUDA(s->ctx, sb::launch_sort(d_keys, d_payload, d_count, max_count, begin_bit, end_bit, sc, s->ctx->num_sms,
                                    static_cast<cudaStream_t>(stream)));
    return SB_OK;
}

// ---------------------------------------------------------------- standalone Preprocessor (Preprocessor<G, ()>)

struct SbPreprocessor {
    SbContext* ctx;
    int sh_fmt, cov_fmt;
    uint32_t stride, n;
    DeviceBuf scratch;  // tile tickets + look-back table, and the V word the kernel also writes
    DeviceBuf visible;
};

SbStatus sb_preprocessor_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, uint64_t n, SbPreprocessor** out) {
    if (!ctx || !out) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    const uint32_t stride = sb_pod_stride(sh_fmt, cov_fmt);
    if (stride == 0) return fail(ctx, SB_ERR_INVALID_ARG, "unknown pod format");
    if (n > 0x3fffffffull) return fail(ctx, SB_ERR_MODEL_TOO_LARGE, "more than 2^30 gaussians");
    if (n * stride > ctx->model_size_limit) return fail(ctx, SB_ERR_MODEL_TOO_LARGE, "model size exceeds the device limit");  // src/preprocessor.rs:381-388
    DeviceGuard device_guard(ctx);
    SbPreprocessor* p = new SbPreprocessor();
    p->ctx = ctx;
    p->sh_fmt = sh_fmt;
    p->cov_fmt = cov_fmt;
    p->stride = stride;
    p->n = (uint32_t)n;
    cudaError_t e = p->scratch.alloc(sb::preprocess_scratch_bytes(p->n, sh_fmt, cov_fmt));
    if (e == cudaSuccess) e = p->visible.alloc(16);
    if (e != cudaSuccess) {
        sb_preprocessor_destroy(p);
        return fail_cuda(ctx, e, "preprocessor alloc");
    }
    *out = p;
    return SB_OK;
}

void sb_preprocessor_destroy(SbPreprocessor* p) {
    if (!p) return;
    DeviceGuard device_guard(p->ctx);
    p->scratch.release();
    p->visible.release();
    delete p;
}

SbStatus sb_preprocessor_preprocess(SbPreprocessor* pre, void* stream, const SbPreprocessorBindGroup* bg, uint32_t gaussian_count) {
    if (!pre || !bg) return fail(pre ? pre->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(pre->ctx);
    SbContext* ctx = pre->ctx;
    if (gaussian_count > pre->n) return fail(ctx, SB_ERR_INVALID_ARG, "gaussian_count exceeds the preprocessor's model size");
    if (bg->gaussians_bytes != (uint64_t)pre->n * pre->stride) return fail(ctx, SB_ERR_BAD_BUFFER_SIZE, "gaussians buffer size != n * size_of::<G>()");
    if (bg->indirect_indices_bytes < (uint64_t)pre->n * 4) return fail(ctx, SB_ERR_BAD_BUFFER_SIZE, "indirect indices buffer smaller than 4 * n");
    const uint32_t padded = sb_padded_key_count(pre->n);
    if (bg->gaussians_depth_bytes < (uint64_t)padded * 4) return fail(ctx, SB_ERR_BAD_BUFFER_SIZE, "gaussians depth buffer smaller than 4 * ceil(n/3840)*3840");
    if (!bg->d_indirect_args || !bg->d_radix_sort_indirect_args || (pre->n && (!bg->d_gaussians || !bg->d_indirect_indices || !bg->d_gaussians_depth)))
        return fail(ctx, SB_ERR_INVALID_ARG, "null buffer in bind group");
    if (reinterpret_cast<uintptr_t>(bg->d_gaussians) & 15u) return fail(ctx, SB_ERR_INVALID_ARG, "device pods must be 16-byte aligned");
    if (bg->gaussian_transform.display_mode > 2 || bg->gaussian_transform.sh_deg > 3) return fail(ctx, SB_ERR_INVALID_ARG, "bad gaussian transform pod");
    sb::PreParams p;
    std::memset(&p, 0, sizeof p);
    p.gaussians = static_cast<const uint8_t*>(bg->d_gaussians);
    p.n = gaussian_count;
    p.selection = bg->d_selection;
    p.invert_selection = bg->invert_selection;
    p.indices = bg->d_indirect_indices;
    p.keys = bg->d_gaussians_depth;
    p.keys_capacity = (uint32_t)std::min<uint64_t>(bg->gaussians_depth_bytes / 4, 0xffffffffull);  // arrayLength(&gaussians_depth), preprocess.wesl:119-122
    p.draw_args = bg->d_indirect_args;
    p.sort_args = bg->d_radix_sort_indirect_args;
    p.recs = nullptr;  // the vertex-stage work belongs to the Renderer here
    p.tboxes = nullptr;
    p.visible_count = pre->visible.as<uint32_t>();
    p.u = make_uniforms(bg->camera, bg->model_transform, bg->gaussian_transform, SB_TARGET_RGBA8_UNORM);
    SB_CUDA(ctx, sb::launch_preprocess(pre->sh_fmt, pre->cov_fmt, p, pre->scratch.p, pre->scratch.bytes, ctx->num_sms,
                                       static_cast<cudaStream_t>(stream)));
    return SB_OK;
}

// ---------------------------------------------------------------- standalone Renderer (Renderer<G, ()>)

struct SbRenderer {
    SbViewer* v;  // owns the record / binning buffers; pods, indices and the count are the caller's
};

SbStatus sb_renderer_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format, uint64_t n, SbRenderer** out) {
    if (!ctx || !out) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(ctx);
    SbViewer* v = nullptr;
    SbStatus s = viewer_new(ctx, sh_fmt, cov_fmt, target_format, n, &v);
    if (s != SB_OK) return s;
    if (!v->use_gather4) {  // the standalone renderer has no gathered-copy path
        sb_viewer_destroy(v);
        return fail(ctx, SB_ERR_INVALID_ARG, "sb_renderer needs the default (TMA gather4) raster path");
    }
    SbRenderer* r = new SbRenderer();
    r->v = v;
    *out = r;
    return SB_OK;
}

void sb_renderer_destroy(SbRenderer* r) {
    if (!r) return;
    sb_viewer_destroy(r->v);
    delete r;
}

SbStatus sb_renderer_set_strict_exp(SbRenderer* r, int32_t strict) {
    if (!r) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    r->v->strict_exp = strict != 0;
    return SB_OK;
}

SbStatus sb_renderer_render(SbRenderer* r, void* stream, const SbRendererBindGroup* bg, const SbTarget* target,
                            const SbDrawIndirectArgs* d_indirect_args, const SbDepthAttachment* depth, int32_t load) {
    if (!r || !bg || !d_indirect_args) return fail(r ? r->v->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(r->v->ctx);
    SbViewer* v = r->v;
    SbContext* ctx = v->ctx;
    if (bg->gaussians_bytes != (uint64_t)v->n * v->stride) return fail(ctx, SB_ERR_BAD_BUFFER_SIZE, "gaussians buffer size != n * size_of::<G>()");
    if (bg->indirect_indices_bytes < (uint64_t)v->n * 4) return fail(ctx, SB_ERR_BAD_BUFFER_SIZE, "indirect indices buffer smaller than 4 * n");
    if (v->n && (!bg->d_gaussians || !bg->d_indirect_indices)) return fail(ctx, SB_ERR_INVALID_ARG, "null buffer in bind group");
    if (reinterpret_cast<uintptr_t>(bg->d_gaussians) & 15u) return fail(ctx, SB_ERR_INVALID_ARG, "device pods must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(bg->d_indirect_indices) & 3u) return fail(ctx, SB_ERR_INVALID_ARG, "indirect indices must be 4-byte aligned");
    if (bg->gaussian_transform.display_mode > 2 || bg->gaussian_transform.sh_deg > 3) return fail(ctx, SB_ERR_INVALID_ARG, "bad gaussian transform pod");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    v->d_gaussians = bg->d_gaussians;
    v->camera = bg->camera;
    v->model_transform = bg->model_transform;
    v->gaussian_transform = bg->gaussian_transform;
    v->ext_indices = bg->d_indirect_indices;
    v->ext_count = &d_indirect_args->instance_count;
    const bool depth_pass = depth && depth->d_depth;
    sb::Uniforms u = make_uniforms(v->camera, v->model_transform, v->gaussian_transform, v->target_format, v->exact_cutoff && !depth_pass);
    v->recs_cut = u.cut_k > 0.0f;  // the vertex stage below already matches the pass (no cut with a depth attachment)
    SbStatus s = check_target(v, target, u);  // validate before enqueuing anything
    if (s == SB_OK) s = report_overflow(v);
    if (s != SB_OK) return s;
    // vertex stage (render.wesl:76-130) for exactly the instances the draw names
    sb::PreParams p;
    std::memset(&p, 0, sizeof p);
    p.gaussians = static_cast<const uint8_t*>(bg->d_gaussians);
    p.n = v->n;
    p.recs = v->recs.as<sb::SplatRec>();
    p.tboxes = v->tboxes.as<sb::TileBox>();
    p.u = u;
    SB_CUDA(ctx, sb::launch_vertex_stage(v->sh_fmt, v->cov_fmt, p, v->ext_indices, v->ext_count, ctx->num_sms, st));
    return do_draw(v, v->camera, v->gaussian_transform, target, load ? 0 : 1, st, depth);
}

// ---------------------------------------------------------------- MultiModelViewer

}  // extern "C"

struct SbMultiModelViewer {
    SbContext* ctx = nullptr;
    int sh_fmt = 0, cov_fmt = 0, target_format = 0;
    SbCameraPod camera;
    SbGaussianTransformPod gaussian_transform;
    std::map<uint64_t, SbViewer*> models;  // HashMap<K, MultiModelViewerModel<G>>
    // render(): the models' preprocess / sort / binning chains run on these side streams, forked from the caller's stream;
    // only the rasters stay in draw order on the caller's stream
    static constexpr int kSide = 4;
    cudaStream_t side[kSide] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr;
    std::vector<cudaEvent_t> bin_done;
};

extern "C" {

SbStatus sb_mm_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format, SbMultiModelViewer** out) {
    if (!ctx || !out) return fail(ctx, SB_ERR_INVALID_ARG, "null");
    if (sb_pod_stride(sh_fmt, cov_fmt) == 0) return fail(ctx, SB_ERR_INVALID_ARG, "unknown pod format");
    SbMultiModelViewer* mm = new SbMultiModelViewer();
    mm->ctx = ctx;
    mm->sh_fmt = sh_fmt;
    mm->cov_fmt = cov_fmt;
    mm->target_format = target_format;
    std::memset(&mm->camera, 0, sizeof mm->camera);
    sb_gaussian_transform_pod(1.0f, SB_MODE_SPLAT, 3, 0, 3.0f, &mm->gaussian_transform);
    *out = mm;
    return SB_OK;
}

void sb_mm_destroy(SbMultiModelViewer* mm) {
    if (!mm) return;
    DeviceGuard device_guard(mm->ctx);
    for (auto& kv : mm->models) sb_viewer_destroy(kv.second);
    for (cudaStream_t st : mm->side)
        if (st) cudaStreamDestroy(st);
    if (mm->fork) cudaEventDestroy(mm->fork);
    for (cudaEvent_t e : mm->bin_done) cudaEventDestroy(e);
    delete mm;
}

SbStatus sb_mm_insert_model(SbMultiModelViewer* mm, uint64_t key, const void* packed_pods, uint64_t n, int32_t* replaced) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(mm->ctx);
    SbViewer* v = nullptr;
    SbStatus s = sb_viewer_create(mm->ctx, mm->sh_fmt, mm->cov_fmt, mm->target_format, packed_pods, n, &v);
    if (s != SB_OK) return s;
    auto it = mm->models.find(key);
    if (replaced) *replaced = it != mm->models.end();
    if (it != mm->models.end()) {
        cudaDeviceSynchronize();
        sb_viewer_destroy(it->second);
        it->second = v;
    } else {
        mm->models[key] = v;
    }
    return SB_OK;
}

SbStatus sb_mm_remove_model(SbMultiModelViewer* mm, uint64_t key, int32_t* removed) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(mm->ctx);
    auto it = mm->models.find(key);
    if (removed) *removed = it != mm->models.end();
    if (it != mm->models.end()) {
        cudaDeviceSynchronize();
        sb_viewer_destroy(it->second);
        mm->models.erase(it);
    }
    return SB_OK;
}

SbStatus sb_mm_update_camera_with_pod(SbMultiModelViewer* mm, const SbCameraPod* pod) {
    if (!mm || !pod) return fail(mm ? mm->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    mm->camera = *pod;
    return SB_OK;
}

SbStatus sb_mm_update_model_transform_with_pod(SbMultiModelViewer* mm, uint64_t key, const SbModelTransformPod* pod) {
    if (!mm || !pod) return fail(mm ? mm->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    it->second->model_transform = *pod;
    return SB_OK;
}

SbStatus sb_mm_update_gaussian_transform_with_pod(SbMultiModelViewer* mm, const SbGaussianTransformPod* pod) {
    if (!mm || !pod) return fail(mm ? mm->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    if (pod->display_mode > 2 || pod->sh_deg > 3) return fail(mm->ctx, SB_ERR_INVALID_ARG, "bad gaussian transform pod");
    mm->gaussian_transform = *pod;
    return SB_OK;
}

SbStatus sb_mm_update_camera(SbMultiModelViewer* mm, const float pos[3], float yaw, float pitch, float z_near, float z_far,
                             float vertical_fov, uint32_t width, uint32_t height) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    return sb_camera_pod(pos, yaw, pitch, z_near, z_far, vertical_fov, width, height, &mm->camera);
}

SbStatus sb_mm_update_model_transform(SbMultiModelViewer* mm, uint64_t key, const float pos[3], const float rot[4], const float scale[3]) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    return sb_model_transform_pod(pos, rot, scale, &it->second->model_transform);
}

SbStatus sb_mm_update_gaussian_transform(SbMultiModelViewer* mm, float size, int32_t display_mode, int32_t sh_deg, int32_t no_sh0,
                                         float max_std_dev) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    SbGaussianTransformPod pod;
    SbStatus s = sb_gaussian_transform_pod(size, display_mode, sh_deg, no_sh0, max_std_dev, &pod);
    if (s != SB_OK) return fail(mm->ctx, s, "bad gaussian transform");
    mm->gaussian_transform = pod;
    return SB_OK;
}

namespace {
SbStatus mm_adopt(SbMultiModelViewer* mm, uint64_t key, SbViewer* v, int32_t* replaced) {
    auto it = mm->models.find(key);
    if (replaced) *replaced = it != mm->models.end();
    if (it != mm->models.end()) {
        cudaDeviceSynchronize();
        sb_viewer_destroy(it->second);
        it->second = v;
    } else {
        mm->models[key] = v;
    }
    return SB_OK;
}
}  // namespace

SbStatus sb_mm_insert_model_from_gaussians(SbMultiModelViewer* mm, uint64_t key, const SbGaussian* src, uint64_t n, int32_t* replaced) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(mm->ctx);
    SbViewer* v = nullptr;
    SbStatus s = sb_viewer_create_from_gaussians(mm->ctx, mm->sh_fmt, mm->cov_fmt, mm->target_format, src, n, &v);
    if (s != SB_OK) return s;
    return mm_adopt(mm, key, v, replaced);
}

SbStatus sb_mm_insert_model_from_device(SbMultiModelViewer* mm, uint64_t key, const void* d_pods, uint64_t d_bytes, uint64_t n,
                                        int32_t* replaced) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(mm->ctx);
    SbViewer* v = nullptr;
    SbStatus s = sb_viewer_create_from_device(mm->ctx, mm->sh_fmt, mm->cov_fmt, mm->target_format, d_pods, d_bytes, n, &v);
    if (s != SB_OK) return s;
    return mm_adopt(mm, key, v, replaced);
}

SbStatus sb_mm_select_rect(SbMultiModelViewer* mm, uint64_t key, void* stream, float x0, float y0, float x1, float y1) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    SbViewer* v = it->second;
    v->camera = mm->camera;  // the world camera buffer the selection pass binds (selection/viewport.rs)
    return sb_viewer_select_rect(v, stream, x0, y0, x1, y1);
}

SbStatus sb_mm_select_brush(SbMultiModelViewer* mm, uint64_t key, void* stream, const float* points_xy, uint32_t n_points, float radius,
                            int32_t accumulate) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    SbViewer* v = it->second;
    v->camera = mm->camera;
    return sb_viewer_select_brush(v, stream, points_xy, n_points, radius, accumulate);
}

SbStatus sb_mm_enable_selection(SbMultiModelViewer* mm, uint64_t key, int32_t enabled, int32_t invert) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    it->second->selection_enabled = enabled != 0;
    it->second->invert_selection = invert ? 1u : 0u;
    return SB_OK;
}

SbStatus sb_mm_read_selection(SbMultiModelViewer* mm, uint64_t key, void* stream, uint32_t* out, uint64_t n_words) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    return sb_viewer_read_selection(it->second, stream, out, n_words);
}

SbStatus sb_mm_set_selection(SbMultiModelViewer* mm, uint64_t key, void* stream, const uint32_t* words, uint64_t n_words, int32_t invert) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    SbViewer* v = it->second;
    if (!words) {
        v->selection_enabled = false;
        return SB_OK;
    }
    SbStatus s = sb_viewer_set_selection(v, stream, words, n_words);
    if (s != SB_OK) return s;
    v->selection_enabled = true;
    v->invert_selection = invert ? 1u : 0u;
    return SB_OK;
}

namespace {
SbStatus mm_render(SbMultiModelViewer* mm, void* stream, const SbTarget* target, const uint64_t* keys, uint32_t n_keys,
                   const SbDepthAttachment* depth, int load) {
    if (!mm || (!keys && n_keys)) return fail(mm ? mm->ctx : nullptr, SB_ERR_INVALID_ARG, "null");
    DeviceGuard device_guard(mm->ctx);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<SbViewer*> models;
    for (uint32_t i = 0; i < n_keys; i++) {  // multi_model.rs:482-489: resolve every key first
        auto it = mm->models.find(keys[i]);
        if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
        models.push_back(it->second);
    }
    for (SbViewer* v : models) {
        SbStatus os = report_overflow(v);
        if (os != SB_OK) return os;
    }
    // Several distinct models: their preprocess / sort / binning chains are independent (own buffers) and mostly latency-
    // bound, so they run concurrently on side streams; the rasters keep the key order on the caller's stream.
    bool concurrent = models.size() > 1;
    for (size_t i = 0; i < models.size() && concurrent; i++)
        for (size_t j = 0; j < i; j++)
            if (models[i] == models[j]) concurrent = false;  // the same model twice shares its buffers: keep it sequential
    if (concurrent) {
        if (!mm->fork) {
            for (cudaStream_t& sd : mm->side) SB_CUDA(mm->ctx, cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
            SB_CUDA(mm->ctx, cudaEventCreateWithFlags(&mm->fork, cudaEventDisableTiming));
        }
        while (mm->bin_done.size() < models.size()) {
            cudaEvent_t e = nullptr;
            SB_CUDA(mm->ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            mm->bin_done.push_back(e);
        }
        // validate before enqueuing anything
        if (!target || !target->d_pixels) return fail(mm->ctx, SB_ERR_INVALID_ARG, "null target");
        SB_CUDA(mm->ctx, cudaEventRecord(mm->fork, st));
        for (cudaStream_t sd : mm->side) SB_CUDA(mm->ctx, cudaStreamWaitEvent(sd, mm->fork, 0));
        for (size_t i = 0; i < models.size(); i++) {  // multi_model.rs:491-503
            cudaStream_t sd = mm->side[i % SbMultiModelViewer::kSide];
            SbStatus s = do_preprocess(models[i], mm->camera, mm->gaussian_transform, sd);
            if (s == SB_OK) s = do_sort(models[i], sd, true);
            if (s != SB_OK) return s;
        }
        int clr = load ? 0 : 1;
        for (size_t i = 0; i < models.size(); i++) {  // multi_model.rs:505-527: one pass, models in key order
            SbStatus s = do_draw(models[i], mm->camera, mm->gaussian_transform, target, clr, mm->side[i % SbMultiModelViewer::kSide], depth, st,
                                 mm->bin_done[i]);
            if (s != SB_OK) return s;
            models[i]->last_stream = st;  // ordered after this model's sort through bin_done
            clr = 0;
        }
        return SB_OK;
    }
    for (SbViewer* v : models) {  // multi_model.rs:491-503
        SbStatus s = do_preprocess(v, mm->camera, mm->gaussian_transform, st);
        if (s != SB_OK) return s;
        s = do_sort(v, st, true);
        if (s != SB_OK) return s;
    }
    if (models.empty()) {  // a render pass that only clears
        if (!target || !target->d_pixels) return fail(mm->ctx, SB_ERR_INVALID_ARG, "null target");
        if (target->format != mm->target_format) return fail(mm->ctx, SB_ERR_INVALID_ARG, "target format mismatch");
        if (!load) SB_CUDA(mm->ctx, sb::launch_clear(*target, st));
        return SB_OK;
    }
    int clear = load ? 0 : 1;
    for (SbViewer* v : models) {  // multi_model.rs:505-527: one pass, models in key order
        SbStatus s = do_draw(v, mm->camera, mm->gaussian_transform, target, clear, st, depth);
        if (s != SB_OK) return s;
        clear = 0;
    }
    return SB_OK;
}

}  // namespace

SbStatus sb_mm_render(SbMultiModelViewer* mm, void* stream, const SbTarget* target, const uint64_t* keys, uint32_t n_keys) {
    return mm_render(mm, stream, target, keys, n_keys, nullptr, 0);
}

SbStatus sb_mm_render_with_pass(SbMultiModelViewer* mm, void* stream, const SbTarget* target, const SbDepthAttachment* depth, int32_t load,
                                const uint64_t* keys, uint32_t n_keys) {
    return mm_render(mm, stream, target, keys, n_keys, depth, load);
}

SbStatus sb_mm_read_model_indices(SbMultiModelViewer* mm, uint64_t key, void* stream, uint32_t* out, uint64_t count,
                                  SbDrawIndirectArgs* draw) {
    if (!mm) return fail(nullptr, SB_ERR_INVALID_ARG, "null");
    auto it = mm->models.find(key);
    if (it == mm->models.end()) return fail(mm->ctx, SB_ERR_MODEL_NOT_FOUND, "model not found");
    SbStatus s = sb_viewer_read_indirect_args(it->second, stream, draw, nullptr);
    if (s != SB_OK) return s;
    return sb_viewer_read_indices(it->second, stream, out, count);
}

}  // extern "C"
