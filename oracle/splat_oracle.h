/*
 * splat_oracle.h — CPU ORACLE for the wgpu-3dgs-viewer per-frame splat pipeline.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (wgpu-3dgs-viewer_b200/csrc) never links, loads or calls anything in oracle/.
 *
 * It restates, in plain C (strict IEEE f32, -ffp-contract=off), the algorithm of
 *   /root/reference/src/shader/preprocess.wesl      (pre / main / post)
 *   /root/reference/src/shader/utils.wesl           (cull, cov2d, cov2d_axes, view_color)
 *   /root/reference/src/shader/camera.wesl          (Camera, world_to_camera)
 *   /root/reference/src/shader/render.wesl          (vert_main, splat/ellipse/point, frag_main)
 *   /root/reference/src/shader/radix_sort.wgsl      (semantics: stable ascending LSD sort)
 *   /root/reference/src/renderer.rs:163-195,284-308 (clear BLACK, ALPHA_BLENDING, draw order)
 *   /root/reference/src/multi_model.rs:476-530      (per-model sort, models drawn in key order)
 *   /root/reference/src/camera.rs:71-93             (look_to_rh / perspective_rh camera)
 *
 * PARITY STATUS: "parity unpinned" for every numeric quantity.  The reference cannot be
 * compiled or run in this environment (no Rust toolchain, no Vulkan), it
 * ships no golden vectors for this path, and half of the arithmetic
 * (pod layouts, unpackers, model/gaussian-transform helpers) lives in the un-vendored
 * crates.io dependency wgpu-3dgs-core 0.6.0 (Cargo.lock:2746-2749).  What IS pinned:
 * the derived known-answer values of SURVEY.md §8(c) (tests/golden/), buffer sizes and
 * the presence/absence behaviour of the reference's e2e tests (tests/test_reference_e2e.py).
 *
 * Evaluation-order contract (WGSL leaves contraction/association to the driver, so the
 * reference itself is not bit-reproducible across backends; the oracle fixes one order):
 * every operation is an individually rounded IEEE-754 binary32 op in the order written
 * in the WGSL source, matrix*vector = ((c0*x + c1*y) + c2*z) + c3*w, no FMA — except
 * in the fragment stage, where the formulas marked "fma" below use fused multiply-add.
 */
#ifndef SPLAT_ORACLE_H
#define SPLAT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- source Gaussian (wgpu_3dgs_core::Gaussian, /root/reference/tests/e2e/viewer.rs:42-48) */
typedef struct {
    float pos[3];
    uint8_t color[4];   /* rgba, rgb already = 0.5 + SH_C0*dc  (utils.wesl:94) */
    float sh[45];       /* 15 coefficients x rgb, coefficient-major: sh[i*3+c] */
    float scale[3];
    float rot[4];       /* quaternion xyzw */
} SoGaussian;           /* 224 bytes */

enum { SO_SH_SINGLE = 0, SO_SH_HALF = 1, SO_SH_NORM8 = 2, SO_SH_NONE = 3 };
enum { SO_COV_SINGLE = 0, SO_COV_HALF = 1, SO_COV_ROT_SCALE = 2 };
enum { SO_MODE_SPLAT = 0, SO_MODE_ELLIPSE = 1, SO_MODE_POINT = 2 };
enum { SO_TARGET_RGBA8 = 0, SO_TARGET_BGRA8 = 1, SO_TARGET_RGBA16F = 2, SO_TARGET_RGBA32F = 3, SO_TARGET_RGBA8_SRGB = 4,
       SO_TARGET_BGRA8_SRGB = 5 };

/* CameraPod: /root/reference/src/buffer/camera.rs:63-80 — 144 bytes */
typedef struct {
    float view[16];   /* column-major */
    float proj[16];
    float size[2];
    uint32_t pad[2];
} SoCameraPod;

/* ModelTransformPod (core, recalled): pos+pad, rot xyzw, scale+pad — 48 bytes */
typedef struct {
    float pos[3];  float pad0;
    float rot[4];
    float scale[3]; float pad1;
} SoModelTransformPod;

/* GaussianTransformPod (core, recalled): size + 4 flag bytes — 8 bytes */
typedef struct {
    float size;
    uint8_t display_mode;   /* 0 splat, 1 ellipse, 2 point */
    uint8_t sh_deg;         /* 0..3 */
    uint8_t no_sh0;         /* 0/1 */
    uint8_t max_std_dev;    /* round(v/3*255) */
} SoGaussianTransformPod;

/* One model of a (multi-model) frame. */
typedef struct {
    const void* pods;            /* packed GaussianPod array */
    uint32_t n;
    int32_t sh_fmt, cov_fmt;
    SoModelTransformPod model_transform;
    const uint32_t* selection;   /* NULL = feature off; else ceil(n/32) words */
    uint32_t invert_selection;   /* reference default 1 */
} SoModel;

/* Statistics the rasterizer roofline needs (pairs counted, not estimated). */
typedef struct {
    uint64_t visible;        /* sum of V over models */
    uint64_t bbox_pixels;    /* (pixel,splat) evaluations inside clipped quad bboxes */
    uint64_t alive_pixels;   /* of those, fragments that were blended */
} SoStats;

uint32_t so_pod_stride(int sh_fmt, int cov_fmt);
void so_pack_gaussians(const SoGaussian* src, uint32_t n, int sh_fmt, int cov_fmt, void* out);

/* PLY (INRIA layout) raw properties -> Gaussian; props = n x 62 f32 in file order. */
void so_gaussians_from_ply_props(const float* props, uint32_t n, SoGaussian* out);

void so_camera_pod(const float pos[3], float yaw, float pitch, float z_near, float z_far,
                   float fov_y, uint32_t width, uint32_t height, SoCameraPod* out);
void so_model_transform_pod(const float pos[3], const float rot_xyzw[4], const float scale[3],
                            SoModelTransformPod* out);
void so_gaussian_transform_pod(float size, int display_mode, int sh_deg, int no_sh0,
                               float max_std_dev, SoGaussianTransformPod* out);

/* preprocess pre+main+post.  keys must hold ceil(n/3840)*3840 floats.  Returns V.
 * indices/keys are written in ascending Gaussian-index order (the canonical order of the
 * reference's racy atomicAdd compaction); visible_mask gets ceil(n/32) words. */
uint32_t so_preprocess(const SoModel* model, const SoCameraPod* cam,
                       const SoGaussianTransformPod* gt,
                       uint32_t* indices, float* keys, uint32_t* visible_mask,
                       uint32_t draw_args[4], uint32_t sort_args[3]);

/* Stable ascending LSD radix sort (8-bit digits x 4) of count (key,payload) pairs. */
void so_radix_sort(uint32_t* keys, uint32_t* payload, uint32_t count);

/* Per-splat projected record (what vert_main computes once per instance). */
typedef struct {
    float cx, cy;           /* pixel-space centre */
    float ax, ay;           /* qx = dx*ax + dy*ay   (major axis row, y-flip folded in) */
    float bx, by;           /* qy = dx*bx + dy*by */
    float r, g, b, a;       /* colour (rgb >= 0, un-clamped above) and opacity */
    float ext_x, ext_y;     /* half extent of the quad's alive region in pixels */
    int32_t valid;          /* 0 = degenerate (axes == 0 or NaN): nothing drawn */
    float z;                /* ndc z of the centre = the depth of every fragment of the quad (render.wesl:123) */
} SoSplat;

void so_project(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
                const uint32_t* indices, uint32_t count, SoSplat* out);

/* Full frame: clear BLACK, then for each model in order: preprocess, sort, draw.
 * target: RGBA8/BGRA8 -> uint8[h*w*4]; RGBA16F -> uint16[h*w*4]; RGBA32F -> float[h*w*4].
 * strict_exp != 0 uses so_exp_neg_poly() instead of libm expf in splat mode.
 * row0/rows restrict the rendered rows (screen strips); pass 0,height for a full frame. */
void so_render(const SoModel* models, uint32_t n_models, const SoCameraPod* cam,
               const SoGaussianTransformPod* gt, int target_format, int strict_exp,
               uint32_t row0, uint32_t rows, void* target, SoStats* stats, int n_threads);

/* Renderer::render_with_pass with an optional depth_stencil state (src/renderer.rs:123, 187-195, 304): one model drawn
 * inside a caller's pass.  load != 0: `target` already holds the pass's colour (LoadOp::Load), else it is cleared to
 * BLACK.  depth (nullable): f32[h*w] attachment; a fragment that survives the shader's discard passes when
 * compare(splat z, depth) holds (1 never, 2 less, 3 equal, 4 less-equal, 5 greater, 6 not-equal, 7 greater-equal,
 * 8 always) and writes its z back when depth_write != 0.  Full frames only. */
void so_render_pass(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
                    int target_format, int strict_exp, int load, void* target, float* depth, int depth_compare,
                    int depth_write, int n_threads);

/* Renderer<G, ()>::render / render_with_pass on a caller's bind group (src/renderer.rs:321-356): one indirect instanced
 * draw of instances 0..count-1, instance i = Gaussian indices[i], composited in that order — no preprocess, no sort.
 * A quad whose centre has clip w <= 0 or z outside [0, w] is clipped as a whole (flat z/w, render.wesl:123). */
void so_draw(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
             const uint32_t* indices, uint32_t count, int target_format, int strict_exp, int load, void* target,
             float* depth, int depth_compare, int depth_write, int n_threads);

/* exp(-x) for x in [0, ~88] from exactly-rounded fma steps (bit-reproducible on GPU). */
float so_exp_neg_poly(float x);

/* viewport selection (next-row f1): /root/reference/src/shader/selection/viewport.wesl:37-69
 * with an analytic rectangle mask [x0,x1) x [y0,y1) in pixels.  Writes ceil(n/32) words. */
void so_select_rect(const SoModel* model, const SoCameraPod* cam,
                    float x0, float y0, float x1, float y1, uint32_t* dest);
/* same with the brush mask (viewport_texture_brush.wesl): texel centres within `radius` of the stroke polyline */
void so_select_brush(const SoModel* model, const SoCameraPod* cam, const float* points_xy, uint32_t n_points,
                     float radius, int accumulate, uint32_t* dest);

int so_max_threads(void);
void so_set_threads(int n);   /* override OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1) */
int so_unorm8_newton_mismatches(void);

#ifdef __cplusplus
}
#endif
#endif
