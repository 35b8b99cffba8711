/*
 * splat_b200.h — C ABI of the B200-native per-frame splat pipeline.
 *
 * Drop-in boundary for the hot path of LioQing/wgpu-3dgs-viewer v0.6.1:
 *   Viewer::render = Preprocessor::preprocess -> RadixSorter::sort -> Renderer::render
 *   (reference src/lib.rs:266-275).
 * The reference exposes this path as a Rust generic API over wgpu handles, not an FFI, so
 * every entry point below names the Rust item it replaces (file:line under the reference
 * root); a Rust shim crate binds these 1:1 (INTEGRATION.md shows the `extern "C"` block).
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no C++/torch types cross the boundary;
 *   - every call returns an SbStatus; sb_last_error_string() describes the last failure;
 *   - `stream` arguments are cudaStream_t handles passed as void* (NULL = default stream);
 *     render/preprocess/sort/draw only ENQUEUE work (like recording into a
 *     wgpu::CommandEncoder) and never synchronise; update_* calls are applied in program
 *     order to the next enqueue (like queue.write_buffer before submit);
 *   - device pointers are marked d_*; everything else is host memory;
 *   - there is no CPU fallback: without a CUDA device sb_ctx_create fails.
 */
#ifndef SPLAT_B200_H
#define SPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_API __attribute__((visibility("default")))

typedef int32_t SbStatus;
enum {
    SB_OK = 0,
    SB_ERR_INVALID_ARG = 1,
    SB_ERR_CUDA = 2,
    SB_ERR_MODEL_TOO_LARGE = 3,   /* PreprocessorCreateError::ModelSizeExceedsDeviceLimit, src/error.rs:7-42 */
    SB_ERR_MODEL_NOT_FOUND = 4,   /* MultiModelViewerAccessError::ModelNotFound, src/error.rs:44-50 */
    SB_ERR_BAD_BUFFER_SIZE = 5,   /* core::FixedSizeBufferWrapperError (TryFrom<wgpu::Buffer>) */
    SB_ERR_IO = 6,
    SB_ERR_OVERFLOW = 7           /* tile-duplicate capacity exceeded in a previous frame */
};

/* GaussianPod format = (SH storage, cov3d storage); core::GaussianPodWith{Sh*}{Cov3d*}Configs.
 * Default pod (src/lib.rs:44) = SB_SH_SINGLE + SB_COV_SINGLE, 224 bytes. */
enum { SB_SH_SINGLE = 0, SB_SH_HALF = 1, SB_SH_NORM8 = 2, SB_SH_NONE = 3 };
enum { SB_COV_SINGLE = 0, SB_COV_HALF = 1, SB_COV_ROT_SCALE = 2 };
/* core::GaussianDisplayMode */
enum { SB_MODE_SPLAT = 0, SB_MODE_ELLIPSE = 1, SB_MODE_POINT = 2 };
/* wgpu::TextureFormat subset accepted as render target (src/renderer.rs:120-155) */
/* The reference renders into any wgpu::TextureFormat (src/renderer.rs:120-155).  The two *_SRGB formats store sRGB codes:
 * every blend decodes the destination to linear, blends, and encodes the result (as fixed-function blending does on an sRGB
 * attachment); Rgba8UnormSrgb is what the reference's own doctests create (src/selection/mod.rs:46). */
enum { SB_TARGET_RGBA8_UNORM = 0, SB_TARGET_BGRA8_UNORM = 1, SB_TARGET_RGBA16_FLOAT = 2, SB_TARGET_RGBA32_FLOAT = 3,
       SB_TARGET_RGBA8_UNORM_SRGB = 4, SB_TARGET_BGRA8_UNORM_SRGB = 5 };

/* core::Gaussian (tests/e2e/viewer.rs:42-48) — the un-packed source splat. */
typedef struct {
    float pos[3];
    uint8_t color[4];
    float sh[45];
    float scale[3];
    float rot[4];     /* xyzw */
} SbGaussian;         /* 224 bytes */

/* CameraPod, src/buffer/camera.rs:63-80 — byte-identical (144 bytes). */
typedef struct {
    float view[16];
    float proj[16];
    float size[2];
    uint32_t _padding[2];
} SbCameraPod;

/* core::ModelTransformPod (48 bytes) */
typedef struct {
    float pos[3];   float _pad0;
    float rot[4];   /* xyzw */
    float scale[3]; float _pad1;
} SbModelTransformPod;

/* core::GaussianTransformPod (8 bytes) */
typedef struct {
    float size;
    uint8_t display_mode;
    uint8_t sh_deg;
    uint8_t no_sh0;
    uint8_t max_std_dev;   /* round(v / 3 * 255) */
} SbGaussianTransformPod;

/* wgpu::util::DrawIndirectArgs — IndirectArgsBuffer, src/buffer/indirect_args.rs:8-55 */
typedef struct { uint32_t vertex_count, instance_count, first_vertex, first_instance; } SbDrawIndirectArgs;
/* wgpu::util::DispatchIndirectArgs — RadixSortIndirectArgsBuffer, src/buffer/indirect_args.rs:59-100 */
typedef struct { uint32_t x, y, z; } SbDispatchIndirectArgs;

/* The render target (replaces &wgpu::TextureView): caller-owned DEVICE memory. */
typedef struct {
    void* d_pixels;
    uint32_t pitch_bytes;   /* row pitch */
    uint32_t width, height; /* must equal CameraPod.size for a full frame */
    int32_t format;         /* SB_TARGET_* */
    uint32_t row0, rows;    /* screen strip rendered into d_pixels (row0 of the frame lands at
                               d_pixels + 0); rows == 0 means the full frame */
} SbTarget;

/* wgpu::CompareFunction of a DepthStencilState (ViewerCreateOptions.depth_stencil, src/lib.rs:279-284) */
enum {
    SB_COMPARE_NEVER = 1, SB_COMPARE_LESS = 2, SB_COMPARE_EQUAL = 3, SB_COMPARE_LESS_EQUAL = 4,
    SB_COMPARE_GREATER = 5, SB_COMPARE_NOT_EQUAL = 6, SB_COMPARE_GREATER_EQUAL = 7, SB_COMPARE_ALWAYS = 8
};
/* Depth attachment of a caller's render pass: Depth32Float, same width/height/strip as the colour target. */
typedef struct SbDepthAttachment {
    void* d_depth;            /* device pointer, f32 per pixel */
    uint32_t pitch_bytes;
    int32_t compare;          /* SB_COMPARE_* (depth_compare) */
    int32_t write_enabled;    /* depth_write_enabled */
} SbDepthAttachment;

typedef struct SbContext SbContext;
typedef struct SbViewer SbViewer;
typedef struct SbMultiModelViewer SbMultiModelViewer;

/* ---------------------------------------------------------------- host-only helpers (no GPU needed) */

SB_API const char* sb_version(void);
SB_API const char* sb_status_string(SbStatus s);

/* size_of::<G>() for a pod format — pinned by tests/e2e/multi_model.rs:43-53 */
SB_API uint32_t sb_pod_stride(int32_t sh_fmt, int32_t cov_fmt);
/* GaussiansBuffer::new's CPU packing (core): Gaussian -> pod bytes; out = n * stride bytes */
SB_API SbStatus sb_pack_gaussians(const SbGaussian* src, uint64_t n, int32_t sh_fmt, int32_t cov_fmt, void* out);
/* wgpu_sort::keys_buffer_size_bytes, src/radix_sorter.rs:937-939 (GaussiansDepthBuffer size) */
SB_API uint64_t sb_keys_buffer_size_bytes(uint32_t n);
/* number of depth keys actually touched by one frame: ceil(n/3840)*3840 */
SB_API uint32_t sb_padded_key_count(uint32_t n);

/* Camera{pos,z,vertical_fov,pitch,yaw} -> CameraPod::new(camera, size): src/camera.rs:71-93,
 * src/buffer/camera.rs:72-80 */
SB_API SbStatus sb_camera_pod(const float pos[3], float yaw, float pitch, float z_near, float z_far,
                              float vertical_fov, uint32_t width, uint32_t height, SbCameraPod* out);
/* ModelTransformPod::new(pos, rot, scale) (core) */
SB_API SbStatus sb_model_transform_pod(const float pos[3], const float rot_xyzw[4], const float scale[3],
                                       SbModelTransformPod* out);
/* GaussianTransformPod::new(size, display_mode, sh_deg, no_sh0, max_std_dev) (core);
 * INVALID_ARG mirrors GaussianShDegree::new / GaussianMaxStdDev::new returning None */
SB_API SbStatus sb_gaussian_transform_pod(float size, int32_t display_mode, int32_t sh_deg, int32_t no_sh0,
                                          float max_std_dev, SbGaussianTransformPod* out);

/* Gaussians::read_from_file(path, GaussiansSource::Ply) (core; examples/simple.rs:157-160).
 * *out is malloc'ed; release with sb_free. */
SB_API SbStatus sb_read_ply(const char* path, SbGaussian** out, uint64_t* n);
/* Gaussians::read_from_file(path, GaussiansSource::Spz) (examples/simple.rs:157-160): Niantic .spz, versions 2 and 3 (gzip).
 * The decoder is in the external wgpu-3dgs-core; the container layout is restated from the published format (recalled). */
SB_API SbStatus sb_read_spz(const char* path, SbGaussian** out, uint64_t* n);
SB_API void sb_free(void* p);

/* ---------------------------------------------------------------- context */

/* replaces the &wgpu::Device every constructor takes */
SB_API SbStatus sb_ctx_create(int32_t device_ordinal, SbContext** out);
SB_API void sb_ctx_destroy(SbContext* ctx);
SB_API const char* sb_last_error_string(const SbContext* ctx); /* ctx may be NULL: last global error */
/* device "max_storage_buffer_binding_size" analogue used for SB_ERR_MODEL_TOO_LARGE
 * (src/preprocessor.rs:239-246); default = free device memory at creation */
SB_API SbStatus sb_ctx_set_model_size_limit(SbContext* ctx, uint64_t bytes);

/* ---------------------------------------------------------------- Viewer (src/lib.rs:65-276) */

/* Viewer::new(device, texture_format, gaussians) with already packed pods (host memory);
 * uploads n*stride bytes (GaussiansBuffer::new). */
SB_API SbStatus sb_viewer_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format,
                                 const void* packed_pods, uint64_t n, SbViewer** out);
/* Viewer::new from source Gaussians (packs on the host, then uploads) */
SB_API SbStatus sb_viewer_create_from_gaussians(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt,
                                                int32_t target_format, const SbGaussian* src, uint64_t n,
                                                SbViewer** out);
/* GaussiansBuffer: TryFrom<wgpu::Buffer> — adopt caller-owned device memory (never freed here);
 * d_bytes must equal n*stride else SB_ERR_BAD_BUFFER_SIZE */
SB_API SbStatus sb_viewer_create_from_device(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt,
                                             int32_t target_format, const void* d_pods, uint64_t d_bytes,
                                             uint64_t n, SbViewer** out);
SB_API void sb_viewer_destroy(SbViewer* v);

/* Viewer::update_camera_with_pod / update_camera: src/lib.rs:202-214 */
SB_API SbStatus sb_viewer_update_camera_with_pod(SbViewer* v, const SbCameraPod* pod);
SB_API SbStatus sb_viewer_update_camera(SbViewer* v, const float pos[3], float yaw, float pitch, float z_near,
                                        float z_far, float vertical_fov, uint32_t width, uint32_t height);
/* Viewer::update_model_transform[_with_pod]: src/lib.rs:217-234 */
SB_API SbStatus sb_viewer_update_model_transform(SbViewer* v, const float pos[3], const float rot_xyzw[4],
                                                 const float scale[3]);
SB_API SbStatus sb_viewer_update_model_transform_with_pod(SbViewer* v, const SbModelTransformPod* pod);
/* Viewer::update_gaussian_transform[_with_pod]: src/lib.rs:237-263 */
SB_API SbStatus sb_viewer_update_gaussian_transform(SbViewer* v, float size, int32_t display_mode, int32_t sh_deg,
                                                    int32_t no_sh0, float max_std_dev);
SB_API SbStatus sb_viewer_update_gaussian_transform_with_pod(SbViewer* v, const SbGaussianTransformPod* pod);

/* viewer.selection_buffer (editor SelectionBuffer: bit i%32 of word i/32; src/lib.rs:137) and
 * viewer.invert_selection_buffer.update (src/selection/buffer.rs:167-171; default invert = 1).
 * Selection is off until sb_viewer_enable_selection is called (feature `viewer-selection`). */
SB_API SbStatus sb_viewer_enable_selection(SbViewer* v, int32_t enabled);
SB_API SbStatus sb_viewer_selection_ptr(SbViewer* v, uint32_t** d_words, uint64_t* n_words);
SB_API SbStatus sb_viewer_set_selection(SbViewer* v, void* stream, const uint32_t* words, uint64_t n_words);
SB_API SbStatus sb_viewer_read_selection(SbViewer* v, void* stream, uint32_t* out, uint64_t n_words); /* synchronises */
SB_API SbStatus sb_viewer_set_invert_selection(SbViewer* v, int32_t invert);
/* selection::viewport evaluation with an analytic rectangle mask:
 * src/shader/selection/viewport.wesl:37-69 + viewport_texture_rectangle.wesl; pixels [x0,x1)x[y0,y1) */
SB_API SbStatus sb_viewer_select_rect(SbViewer* v, void* stream, float x0, float y0, float x1, float y1);
/* selection::viewport evaluation with an analytic brush mask: src/shader/selection/viewport_texture_brush.wesl draws,
 * per stroke segment, a disc of `radius` at both ends and a quad between them into the mask texture — a capsule.  Texel
 * (ix,iy) is set iff its centre lies within `radius` of a segment of the polyline points_xy[0..n_points) (host floats
 * x0,y0,x1,y1,..; pixels; 1 <= n_points <= SB_BRUSH_MAX_POINTS, one point = a dab).  accumulate != 0 ORs the result into
 * the selection words, the way successive strokes accumulate in the reference's texture; 0 replaces them. */
#define SB_BRUSH_MAX_POINTS 64
SB_API SbStatus sb_viewer_select_brush(SbViewer* v, void* stream, const float* points_xy, uint32_t n_points, float radius,
                                       int32_t accumulate);

/* Editor-style per-Gaussian edit feeding the buffer: the colour-override use of the editor crate's
 * NonDestructiveModifier<BasicSelectionModifier> exactly as tests/e2e/selection.rs:54-116 drives it
 * (BasicColorRgbOverrideOrHsvModifiersPod::new_rgb_override(rgb), alpha, contrast 0, exposure 0, gamma 1).  Every call
 * starts from the source colours (snapshot taken at the first call): Gaussians whose selection bit is set get
 * colour = pack4x8unorm(rgb, source alpha * alpha), all others their source colour again — in the viewer's Gaussian
 * buffer, in place, like modifier.apply(..., &viewer.gaussians_buffer, ...).  SH coefficients are left untouched
 * (the HSV / contrast / exposure / gamma edits of the external crate are not built).  restore drops the edit. */
SB_API SbStatus sb_viewer_apply_rgb_override(SbViewer* v, void* stream, const float rgb[3], float alpha);
SB_API SbStatus sb_viewer_restore_gaussians(SbViewer* v, void* stream);
/* The editor's full BasicColorModifiers (`basic_color_modifiers_buffer.update(queue, rgb_or_hsv, alpha, contrast, exposure, gamma)`,
 * tests/e2e/selection.rs:80-92; examples/selection.rs:218-229) on the selected Gaussians, non-destructively like the override:
 * rgb override or HSV (hue shift in turns, saturation scale, value scale), then contrast, exposure (stops), gamma, alpha scale.
 * Neutral: {0, {0,1,1}, 1, 0, 0, 1}.  The arithmetic is the external wgpu-3dgs-editor's, restated from its documented meaning. */
typedef struct SbBasicColorModifiers {
    int32_t rgb_override;   /* 1: rgb_or_hsv is the override colour; 0: it is (hue shift, saturation scale, value scale) */
    float rgb_or_hsv[3];
    float alpha, contrast, exposure, gamma;
} SbBasicColorModifiers;
SB_API SbStatus sb_viewer_apply_basic_color_modifiers(SbViewer* v, void* stream, const SbBasicColorModifiers* m);

/* Viewer::render(encoder, texture_view): src/lib.rs:266-275 — enqueue the whole frame */
SB_API SbStatus sb_viewer_render(SbViewer* v, void* stream, const SbTarget* target);
/* The three public stages, individually (viewer.preprocessor / radix_sorter / renderer):
 *   Preprocessor::preprocess  src/preprocessor.rs:283-290
 *   RadixSorter::sort         src/radix_sorter.rs:56-66
 *   Renderer::render          src/renderer.rs:163-184 */
SB_API SbStatus sb_viewer_preprocess(SbViewer* v, void* stream);
SB_API SbStatus sb_viewer_sort(SbViewer* v, void* stream);
SB_API SbStatus sb_viewer_draw(SbViewer* v, void* stream, const SbTarget* target);
/* Renderer::render_with_pass (src/renderer.rs:187-195) inside a caller's render pass, with the pipeline's optional
 * depth_stencil state (src/renderer.rs:123, 304): the splats composite over what the target already holds when
 * load != 0 (LoadOp::Load; 0 = clear to BLACK first), and every fragment that survives the shader's discard is depth
 * tested against `depth` (nullable) with its splat's depth = the centre's ndc z (render.wesl:123), writing it back when
 * write_enabled.  Runs the preprocess and sort stages too when run_stages != 0 (Viewer::render), else only the draw. */
SB_API SbStatus sb_viewer_render_with_pass(SbViewer* v, void* stream, const SbTarget* target, const SbDepthAttachment* depth,
                                           int32_t load, int32_t run_stages);

/* Convenience for host-resident callers: update camera, render into an internal device
 * target, copy the frame to host_pixels (pinned recommended) — all enqueued on `stream`. */
SB_API SbStatus sb_viewer_render_to_host(SbViewer* v, void* stream, const SbCameraPod* cam,
                                         void* host_pixels, uint64_t host_bytes);

/* A batch of camera views of the same scene (BASELINE config 5a; the reference would call
 * update_camera + render once per view, tests/e2e/viewer.rs:23-34).  View i is rendered with cams[i]
 * into targets[i]; if `targets` is NULL each frame is rendered into an internal target and copied to
 * host_pixels[i] (pinned recommended).  Views are pipelined two deep on internal streams over a twin
 * set of frame buffers — the latency-bound stages of one view overlap the rasterizer of the other —
 * forked from and joined back to `stream`, so the call is ordered on `stream` like any other. */
SB_API SbStatus sb_viewer_render_batch(SbViewer* v, void* stream, const SbCameraPod* cams, const SbTarget* targets,
                                       void* const* host_pixels, uint32_t count);

/* Public buffers of the Viewer (src/lib.rs:65-82) as device pointers */
SB_API SbStatus sb_viewer_gaussians_ptr(SbViewer* v, const void** d_pods, uint64_t* bytes);
SB_API SbStatus sb_viewer_indirect_args_ptr(SbViewer* v, const SbDrawIndirectArgs** d_args);
SB_API SbStatus sb_viewer_radix_sort_indirect_args_ptr(SbViewer* v, const SbDispatchIndirectArgs** d_args);
SB_API SbStatus sb_viewer_indirect_indices_ptr(SbViewer* v, const uint32_t** d_indices, uint64_t* count);
SB_API SbStatus sb_viewer_gaussians_depth_ptr(SbViewer* v, const float** d_keys, uint64_t* bytes);
/* synchronising read-backs for tests/tools (cudaMemcpy D2H after syncing `stream`) */
SB_API SbStatus sb_viewer_read_indirect_args(SbViewer* v, void* stream, SbDrawIndirectArgs* draw,
                                             SbDispatchIndirectArgs* dispatch);
SB_API SbStatus sb_viewer_read_indices(SbViewer* v, void* stream, uint32_t* out, uint64_t count);
SB_API SbStatus sb_viewer_read_depth_keys(SbViewer* v, void* stream, float* out, uint64_t count);
/* frame statistics of the last rendered frame (synchronises): visible splats, (splat,tile)
 * duplicates, overflow flag */
SB_API SbStatus sb_viewer_read_frame_stats(SbViewer* v, void* stream, uint64_t* visible, uint64_t* duplicates,
                                           uint32_t* overflowed);
/* how the rasterizer fetches splat records: 1 = TMA tile::gather4 straight from the per-Gaussian array
 * (default), 0 = gathered copy + 1-D TMA bulk copies (environment SB_RASTER_PATH=bulk at viewer creation) */
SB_API SbStatus sb_viewer_raster_path(SbViewer* v, int32_t* tma_gather4);
/* testing knob: 1 = bit-reproducible polynomial exp in the fragment stage (matches the
 * oracle's strict_exp); 0 (default) = MUFU ex2 fast path */
SB_API SbStatus sb_viewer_set_strict_exp(SbViewer* v, int32_t strict);
/* Exact alpha cut-off (default on): in splat mode on a unorm8 target a blend with alpha < 0.5/255 returns the destination
 * unchanged after re-quantisation, so each splat is shrunk to the radius beyond which alpha = a*exp(-r^2) stays below that:
 * fewer (splat, tile) duplicates and fragments, bit-identical frames.  Never applied to float targets, ellipse/point modes
 * or depth-tested passes.  0 switches it off (the instrumented fragment counts then equal the reference's). */
SB_API SbStatus sb_viewer_set_exact_cutoff(SbViewer* v, int32_t enabled);
/* One frame split into screen strips over several GPUs (BASELINE config 5b; the reference has no multi-GPU code —
 * the unit being sharded is Viewer::render, src/lib.rs:266-275).  With strip culling on, a sb_viewer_render whose target is a
 * strip (SbTarget.rows != 0) runs the reference's full-frame cull as always, then keeps only the visible splats whose tile box
 * meets the strip: the indirect args, indices and keys of that frame describe the strip's own visible set (a subset of the
 * full frame's, in the same order), and the depth sort, binning and colour evaluation run on that subset only.  The strips of
 * all ranks reassemble the single-GPU frame bit for bit.  Default off (a strip render then keeps the full-frame artefacts). */
SB_API SbStatus sb_viewer_set_strip_cull(SbViewer* v, int32_t enabled);
/* Measured denominators for the rasterizer's roofline (no reference counterpart; SURVEY 8d asks for "achieved FP32 and
 * shared-memory throughput vs peak"): two microbenchmarks run on the context's device — FP32 issue in lane-ops/s with an FMA
 * counted once, and conflict-free shared-memory read bandwidth in bytes/s.  Synchronises. */
SB_API SbStatus sb_probe_peaks(SbContext* ctx, void* stream, double* fp32_lane_ops_per_s, double* smem_bytes_per_s);
/* Work per 16-pixel tile row of the last binned frame: out[y] = number of (splat, tile) duplicates in tile row y (synchronises).
 * After a full-frame render it is what a strip partition balances on (rasterizer and binning time follow the duplicates,
 * not the rows): splat_b200/sharding.py: balanced_strips. */
SB_API SbStatus sb_viewer_read_tile_row_work(SbViewer* v, void* stream, uint64_t* out, uint32_t n_rows);
/* Strip gather without a gather: the rank that owns the final frame allocates it with sb_shared_frame_create and publishes the
 * 64-byte handle (any transport: the tests and bench.py use torch.distributed); every other process of the node opens it and
 * gets a device pointer, valid on ITS GPU, that aliases the owner's memory over NVLink (CUDA IPC, peer access enabled on
 * open).  Each rank then renders its strip straight into `frame + row0 * pitch` — the rasterizer's final pixel stores are the
 * transfer — and one barrier (NCCL) after the raster kernels orders them before the owner reads the frame. */
#define SB_SHARED_HANDLE_BYTES 64
SB_API SbStatus sb_shared_frame_create(SbContext* ctx, uint64_t bytes, void** d_frame, uint8_t handle[SB_SHARED_HANDLE_BYTES]);
SB_API SbStatus sb_shared_frame_open(SbContext* ctx, const uint8_t handle[SB_SHARED_HANDLE_BYTES], void** d_frame);
SB_API SbStatus sb_shared_frame_close(SbContext* ctx, void* d_frame);    /* a pointer obtained from _open */
SB_API SbStatus sb_shared_frame_destroy(SbContext* ctx, void* d_frame);  /* a pointer obtained from _create */
/* Strips with the Preprocessor's work partitioned as well (the scalable form of the strip case; same frames bit for bit): rank r
 * preprocesses Gaussians [n r / G, n (r + 1) / G) only and stores, over NVLink, the (index, key) pairs, splat records and tile
 * boxes each strip needs straight into the owning rank's buffers; after ONE barrier across the ranks (the caller's: any
 * stream-ordered collective) each rank strings its inbox together — ascending source rank = ascending Gaussian index, i.e. the
 * strip's visible list exactly as a single-GPU cull compacts it — and runs the unchanged sort / binning / rasterizer on it.
 *   create   (every rank) viewer = the replicated scene; row0s/rows = every rank's strip (whole tile rows; rows 0 = no strip);
 *            `exported` = this rank's three IPC handles, to be given to every other rank
 *   connect  (every rank) all ranks' exports, in rank order
 *   scatter  enqueue: K1 on the slice + the scatter into the peers      -> barrier across ranks ->
 *   render   enqueue: concat + depth sort + binning + rasterizer of this rank's strip into `target` (SbTarget.row0/rows = its strip)
 * After `render` the viewer's artefacts (args, indices, keys) describe the strip's visible set, as with sb_viewer_set_strip_cull.
 * A second barrier must separate a frame's `render` from the next frame's `scatter` (the strip fence of the shared frame does). */
typedef struct SbStrips SbStrips;
typedef struct SbStripsExport { uint8_t handles[3][SB_SHARED_HANDLE_BYTES]; } SbStripsExport;
SB_API SbStatus sb_strips_create(SbViewer* v, uint32_t world, uint32_t rank, const uint32_t* row0s, const uint32_t* rows, SbStrips** out,
                                 SbStripsExport* exported);
SB_API SbStatus sb_strips_connect(SbStrips* s, const SbStripsExport* all_ranks);
SB_API SbStatus sb_strips_scatter(SbStrips* s, void* stream);
SB_API SbStatus sb_strips_render(SbStrips* s, void* stream, const SbTarget* target);
SB_API void sb_strips_destroy(SbStrips* s);
/* Tracing (the reference has none: every pass has timestamp_writes: None, src/radix_sorter.rs:504-507).
 * When enabled, cudaEvents are recorded on the launching stream between the stages of a frame;
 * ms[0..5] = preprocess, depth sort, tile count+emit, tile sort, gather, raster. */
SB_API SbStatus sb_viewer_set_stage_timing(SbViewer* v, int32_t enabled);
SB_API SbStatus sb_viewer_read_stage_times(SbViewer* v, void* stream, float ms[6]);
/* Instrumented rasterizer build (separate kernel instantiation, off by default): counts the
 * fragments that were blended and the (pixel, splat) lane pairs that were evaluated in the last
 * frame — the pair counts the FP32 roofline of the rasterizer is computed from. */
SB_API SbStatus sb_viewer_set_raster_counting(SbViewer* v, int32_t enabled);
SB_API SbStatus sb_viewer_read_raster_counters(SbViewer* v, void* stream, uint64_t* alive, uint64_t* evaluated);
/* same instrumented build: (warp, splat) evaluations issued, and how many of them had at least one alive lane */
SB_API SbStatus sb_viewer_read_raster_warp_counters(SbViewer* v, void* stream, uint64_t* warp_evals, uint64_t* warp_evals_alive);
/* capacity (in duplicates) of the tile-binning buffers; default 8*n + tiles */
SB_API SbStatus sb_viewer_reserve_duplicates(SbViewer* v, uint64_t capacity);

/* ---------------------------------------------------------------- standalone sort (RadixSorter<()>) */

/* RadixSorter::new_without_bind_groups + sort(encoder, bind_groups, args): src/radix_sorter.rs:71-96.
 * Sorts the first min(d_args->x*3840, capacity) ... see DESIGN.md: `d_count` holds the key count.
 * Stable ascending sort of (u32 key, u32 payload); result in d_keys/d_payload. */
typedef struct SbRadixSorter SbRadixSorter;
SB_API SbStatus sb_sorter_create(SbContext* ctx, uint32_t capacity, SbRadixSorter** out);
SB_API void sb_sorter_destroy(SbRadixSorter* s);
SB_API SbStatus sb_sorter_sort(SbRadixSorter* s, void* stream, uint32_t* d_keys, uint32_t* d_payload,
                               const uint32_t* d_count, uint32_t max_count, int32_t begin_bit, int32_t end_bit);

/* ---------------------------------------------------------------- standalone Preprocessor / Renderer (the B = () variants) */

/* Preprocessor::new_without_bind_group + create_bind_group + preprocess(encoder, bind_group, gaussian_count):
 * src/preprocessor.rs:375-435, 37-68, 438-450.  The bind group (layout src/preprocessor.rs:104-221) is this struct of
 * caller-owned device buffers plus the three uniform pods (bindings 0-2), passed per call; sizes are validated
 * the way TryFrom<wgpu::Buffer> validates them (SB_ERR_BAD_BUFFER_SIZE). */
typedef struct SbPreprocessorBindGroup {
    SbCameraPod camera;                                   /* binding 0 */
    SbModelTransformPod model_transform;                  /* binding 1 */
    SbGaussianTransformPod gaussian_transform;            /* binding 2 */
    const void* d_gaussians;                              /* binding 3: GaussiansBuffer<G> */
    uint64_t gaussians_bytes;                             /*   must equal n * size_of::<G>() */
    SbDrawIndirectArgs* d_indirect_args;                  /* binding 4: 16 bytes */
    SbDispatchIndirectArgs* d_radix_sort_indirect_args;   /* binding 5: 12 bytes */
    uint32_t* d_indirect_indices;                         /* binding 6: IndirectIndicesBuffer */
    uint64_t indirect_indices_bytes;                      /*   >= 4 * n */
    float* d_gaussians_depth;                             /* binding 7: GaussiansDepthBuffer */
    uint64_t gaussians_depth_bytes;                       /*   >= 4 * sb_padded_key_count(n) (the reference allocates 4x that) */
    const uint32_t* d_selection;                          /* binding 8: ceil(n/32) words; NULL = feature off */
    uint32_t invert_selection;                            /* binding 9 */
} SbPreprocessorBindGroup;
typedef struct SbPreprocessor SbPreprocessor;
SB_API SbStatus sb_preprocessor_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, uint64_t n, SbPreprocessor** out);
SB_API void sb_preprocessor_destroy(SbPreprocessor* p);
/* enqueue pre+main+post over the first gaussian_count (<= n) Gaussians: visible indices (ascending), depth keys,
 * pad keys, {6,V,0,0} and {ceil(V/3840),1,1}.  The visible count for a following sb_sorter_sort is
 * &d_indirect_args->instance_count. */
SB_API SbStatus sb_preprocessor_preprocess(SbPreprocessor* p, void* stream, const SbPreprocessorBindGroup* bind_group,
                                           uint32_t gaussian_count);

/* Renderer::new_without_bind_group + create_bind_group + render(encoder, view, bind_group, indirect_args) /
 * render_with_pass: src/renderer.rs:247-318, 23-41, 321-356.  Bind group layout src/renderer.rs:56-116. */
typedef struct SbRendererBindGroup {
    SbCameraPod camera;                                   /* binding 0 */
    SbModelTransformPod model_transform;                  /* binding 1 */
    SbGaussianTransformPod gaussian_transform;            /* binding 2 */
    const void* d_gaussians;                              /* binding 3 */
    uint64_t gaussians_bytes;
    const uint32_t* d_indirect_indices;                   /* binding 4: instance i draws Gaussian indices[i] */
    uint64_t indirect_indices_bytes;
} SbRendererBindGroup;
typedef struct SbRenderer SbRenderer;
SB_API SbStatus sb_renderer_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format, uint64_t n,
                                   SbRenderer** out);
SB_API void sb_renderer_destroy(SbRenderer* r);
/* One indirect instanced draw: instances 0..d_indirect_args->instance_count-1 in order, whatever produced the indices
 * (the vertex stage runs here for exactly those Gaussians; a quad whose centre has w <= 0 or ndc z outside [0,1] is
 * clipped as a whole, as the fixed-function clipper does with render.wesl:123's flat z/w).  load == 0 clears to BLACK
 * first (Renderer::render), load != 0 composites over the target (render_with_pass inside a caller's pass); depth is
 * the optional depth_stencil attachment as in sb_viewer_render_with_pass. */
SB_API SbStatus sb_renderer_render(SbRenderer* r, void* stream, const SbRendererBindGroup* bind_group, const SbTarget* target,
                                   const SbDrawIndirectArgs* d_indirect_args, const SbDepthAttachment* depth, int32_t load);
SB_API SbStatus sb_renderer_set_strict_exp(SbRenderer* r, int32_t strict);

/* ---------------------------------------------------------------- MultiModelViewer (src/multi_model.rs:291-531) */

SB_API SbStatus sb_mm_create(SbContext* ctx, int32_t sh_fmt, int32_t cov_fmt, int32_t target_format,
                             SbMultiModelViewer** out);
SB_API void sb_mm_destroy(SbMultiModelViewer* mm);
/* insert_model(device, key, gaussians) -> Option<old>: *replaced = 1 when a model was replaced */
SB_API SbStatus sb_mm_insert_model(SbMultiModelViewer* mm, uint64_t key, const void* packed_pods, uint64_t n,
                                   int32_t* replaced);
SB_API SbStatus sb_mm_remove_model(SbMultiModelViewer* mm, uint64_t key, int32_t* removed);
SB_API SbStatus sb_mm_update_camera_with_pod(SbMultiModelViewer* mm, const SbCameraPod* pod);
SB_API SbStatus sb_mm_update_model_transform_with_pod(SbMultiModelViewer* mm, uint64_t key,
                                                      const SbModelTransformPod* pod);
SB_API SbStatus sb_mm_update_gaussian_transform_with_pod(SbMultiModelViewer* mm, const SbGaussianTransformPod* pod);
SB_API SbStatus sb_mm_set_selection(SbMultiModelViewer* mm, uint64_t key, void* stream, const uint32_t* words,
                                    uint64_t n_words, int32_t invert);
/* update_camera / update_model_transform / update_gaussian_transform (non-pod variants): src/multi_model.rs:398-464 */
SB_API SbStatus sb_mm_update_camera(SbMultiModelViewer* mm, const float pos[3], float yaw, float pitch, float z_near, float z_far,
                                    float vertical_fov, uint32_t width, uint32_t height);
SB_API SbStatus sb_mm_update_model_transform(SbMultiModelViewer* mm, uint64_t key, const float pos[3], const float rot_xyzw[4],
                                             const float scale[3]);
SB_API SbStatus sb_mm_update_gaussian_transform(SbMultiModelViewer* mm, float size, int32_t display_mode, int32_t sh_deg,
                                                int32_t no_sh0, float max_std_dev);
/* insert_model(device, key, &impl IterGaussian) packing source Gaussians on the host (src/multi_model.rs:353-362), and
 * insert_model_with on a caller-owned device buffer (MultiModelViewerGaussianBuffers over an adopted GaussiansBuffer,
 * src/multi_model.rs:366-391; never freed here; d_bytes must equal n * stride) */
SB_API SbStatus sb_mm_insert_model_from_gaussians(SbMultiModelViewer* mm, uint64_t key, const SbGaussian* src, uint64_t n,
                                                  int32_t* replaced);
SB_API SbStatus sb_mm_insert_model_from_device(SbMultiModelViewer* mm, uint64_t key, const void* d_pods, uint64_t d_bytes, uint64_t n,
                                               int32_t* replaced);
/* models[key].gaussian_buffers.selection_buffer / invert_selection_buffer (src/multi_model.rs:81-84): viewport selection
 * evaluated against the shared camera (selection::viewport, as sb_viewer_select_rect / _brush), the per-model switch of the
 * preprocessor's mask test, and a synchronising read-back */
SB_API SbStatus sb_mm_select_rect(SbMultiModelViewer* mm, uint64_t key, void* stream, float x0, float y0, float x1, float y1);
SB_API SbStatus sb_mm_select_brush(SbMultiModelViewer* mm, uint64_t key, void* stream, const float* points_xy, uint32_t n_points,
                                   float radius, int32_t accumulate);
SB_API SbStatus sb_mm_enable_selection(SbMultiModelViewer* mm, uint64_t key, int32_t enabled, int32_t invert);
SB_API SbStatus sb_mm_read_selection(SbMultiModelViewer* mm, uint64_t key, void* stream, uint32_t* out, uint64_t n_words);
/* render(encoder, view, keys): models drawn in key order, model k+1 over model k */
SB_API SbStatus sb_mm_render(SbMultiModelViewer* mm, void* stream, const SbTarget* target, const uint64_t* keys,
                             uint32_t n_keys);
/* MultiModelViewer::new_with_options(depth_stencil) (src/multi_model.rs:319-337): the shared Renderer then carries a
 * DepthStencilState, and the models are drawn with renderer.render_with_pass inside a pass of the caller's that owns the depth
 * attachment (MultiModelViewer::render itself attaches none, src/multi_model.rs:505-527).  Models composite in key order over the
 * target (load != 0: LoadOp::Load, else cleared to BLACK) and are depth tested / written as in sb_viewer_render_with_pass. */
SB_API SbStatus sb_mm_render_with_pass(SbMultiModelViewer* mm, void* stream, const SbTarget* target, const SbDepthAttachment* depth,
                                       int32_t load, const uint64_t* keys, uint32_t n_keys);
SB_API SbStatus sb_mm_read_model_indices(SbMultiModelViewer* mm, uint64_t key, void* stream, uint32_t* out,
                                         uint64_t count, SbDrawIndirectArgs* draw);

#ifdef __cplusplus
}
#endif
#endif
