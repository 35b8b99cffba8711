// splat_b200.hpp — header-only C++17 mirror of the reference's Rust API over the C ABI.
// (The reference's host language is Rust; no Rust toolchain exists in this image, so the typed host
// side is C++: same names, argument meaning and error behaviour — DESIGN.md §1.)
#pragma once

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/splat_b200.h"

namespace splat_b200 {

// ViewerCreateError / MultiModelViewerAccessError (reference src/error.rs:7-50)
struct Error : std::runtime_error {
    SbStatus status;
    Error(SbStatus s, const std::string& m) : std::runtime_error(m), status(s) {}
};
struct ModelSizeExceedsDeviceLimit : Error { using Error::Error; };
struct ModelNotFound : Error { using Error::Error; };

inline void check(SbStatus s, const SbContext* ctx = nullptr) {
    if (s == SB_OK) return;
    const std::string msg = sb_last_error_string(ctx);
    if (s == SB_ERR_MODEL_TOO_LARGE) throw ModelSizeExceedsDeviceLimit(s, msg);
    if (s == SB_ERR_MODEL_NOT_FOUND) throw ModelNotFound(s, msg);
    throw Error(s, msg);
}

// Camera (reference src/camera.rs:18-93): yaw/pitch FPS camera, look_to_rh + perspective_rh.
struct Camera {
    float pos[3] = {0, 0, 0};
    float z_near = 0.1f, z_far = 1e4f;
    float vertical_fov = 1.04719755f;
    float pitch = 0, yaw = 0;
    Camera() = default;
    Camera(float z_near_, float z_far_, float fov) : z_near(z_near_), z_far(z_far_), vertical_fov(fov) {}
    void get_forward(float f[3]) const { f[0] = std::cos(pitch) * std::sin(yaw); f[1] = std::sin(pitch); f[2] = std::cos(pitch) * std::cos(yaw); }
    void move_by(float forward, float right) {
        float f[3]; get_forward(f);
        float r[3] = {-f[2], 0.0f, f[0]};  // cross(forward, UP) normalised below
        const float rl = std::sqrt(r[0] * r[0] + r[2] * r[2]);
        for (int i = 0; i < 3; i++) pos[i] += f[i] * forward + (rl > 0 ? r[i] / rl : 0.0f) * right;
    }
    void move_up(float up) { pos[1] += up; }
    void pitch_by(float d) { const float lim = 1.57079633f - 1e-6f; pitch = std::fmin(std::fmax(pitch + d, -lim), lim); }
    void yaw_by(float d) { const float two_pi = 6.28318531f; yaw = std::fmod(std::fmod(yaw + d, two_pi) + two_pi, two_pi); }
    SbCameraPod pod(uint32_t width, uint32_t height) const {  // CameraPod::new(camera, size)
        SbCameraPod p;
        check(sb_camera_pod(pos, yaw, pitch, z_near, z_far, vertical_fov, width, height, &p));
        return p;
    }
};

class Context {  // replaces &wgpu::Device
public:
    explicit Context(int device = 0) { check(sb_ctx_create(device, &h_)); }
    ~Context() { sb_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    SbContext* raw() const { return h_; }
private:
    SbContext* h_ = nullptr;
};

// Viewer<G> (reference src/lib.rs:65-276); G = (sh_fmt, cov_fmt)
class Viewer {
public:
    Viewer(Context& ctx, int32_t texture_format, const std::vector<SbGaussian>& gaussians, int32_t sh_fmt = SB_SH_SINGLE, int32_t cov_fmt = SB_COV_SINGLE)
        : ctx_(ctx) { check(sb_viewer_create_from_gaussians(ctx.raw(), sh_fmt, cov_fmt, texture_format, gaussians.data(), gaussians.size(), &h_), ctx.raw()); }
    ~Viewer() { sb_viewer_destroy(h_); }
    Viewer(const Viewer&) = delete;
    Viewer& operator=(const Viewer&) = delete;
    void update_camera(const Camera& c, uint32_t w, uint32_t h) { update_camera_with_pod(c.pod(w, h)); }
    void update_camera_with_pod(const SbCameraPod& p) { check(sb_viewer_update_camera_with_pod(h_, &p), ctx_.raw()); }
    void update_model_transform(const float pos[3], const float rot_xyzw[4], const float scale[3]) { check(sb_viewer_update_model_transform(h_, pos, rot_xyzw, scale), ctx_.raw()); }
    void update_gaussian_transform(float size, int mode, int sh_deg, bool no_sh0, float max_std_dev) { check(sb_viewer_update_gaussian_transform(h_, size, mode, sh_deg, no_sh0, max_std_dev), ctx_.raw()); }
    void render(void* stream, const SbTarget& t) { check(sb_viewer_render(h_, stream, &t), ctx_.raw()); }        // Viewer::render
    void preprocess(void* stream) { check(sb_viewer_preprocess(h_, stream), ctx_.raw()); }                      // viewer.preprocessor.preprocess
    void sort(void* stream) { check(sb_viewer_sort(h_, stream), ctx_.raw()); }                                  // viewer.radix_sorter.sort
    void draw(void* stream, const SbTarget& t) { check(sb_viewer_draw(h_, stream, &t), ctx_.raw()); }           // viewer.renderer.render
    // viewer.renderer.render_with_pass inside the caller's pass (+ ViewerCreateOptions.depth_stencil): src/renderer.rs:187-195
    void render_with_pass(void* stream, const SbTarget& t, const SbDepthAttachment* depth = nullptr, bool load = true, bool run_stages = false) {
        check(sb_viewer_render_with_pass(h_, stream, &t, depth, load, run_stages), ctx_.raw());
    }
    // selection::ViewportSelector evaluation with an analytic rectangle / brush mask
    void select_rect(void* stream, float x0, float y0, float x1, float y1) { check(sb_viewer_select_rect(h_, stream, x0, y0, x1, y1), ctx_.raw()); }
    void select_brush(void* stream, const std::vector<float>& points_xy, float radius, bool accumulate = false) {
        check(sb_viewer_select_brush(h_, stream, points_xy.data(), (uint32_t)(points_xy.size() / 2), radius, accumulate), ctx_.raw());
    }
    // editor NonDestructiveModifier + rgb override on the selected Gaussians (tests/e2e/selection.rs:54-116)
    void apply_rgb_override(void* stream, const float rgb[3], float alpha = 1.0f) { check(sb_viewer_apply_rgb_override(h_, stream, rgb, alpha), ctx_.raw()); }
    void restore_gaussians(void* stream) { check(sb_viewer_restore_gaussians(h_, stream), ctx_.raw()); }
    SbViewer* raw() const { return h_; }
private:
    Context& ctx_;
    SbViewer* h_ = nullptr;
};

// MultiModelViewer<G, K = u64> (reference src/multi_model.rs:291-531)
class MultiModelViewer {
public:
    MultiModelViewer(Context& ctx, int32_t texture_format, int32_t sh_fmt = SB_SH_SINGLE, int32_t cov_fmt = SB_COV_SINGLE)
        : ctx_(ctx), sh_(sh_fmt), cov_(cov_fmt) { check(sb_mm_create(ctx.raw(), sh_fmt, cov_fmt, texture_format, &h_), ctx.raw()); }
    ~MultiModelViewer() { sb_mm_destroy(h_); }
    bool insert_model(uint64_t key, const std::vector<SbGaussian>& g) {
        std::vector<uint8_t> pods((size_t)sb_pod_stride(sh_, cov_) * g.size() + 16);
        check(sb_pack_gaussians(g.data(), g.size(), sh_, cov_, pods.data()), ctx_.raw());
        int32_t replaced = 0;
        check(sb_mm_insert_model(h_, key, pods.data(), g.size(), &replaced), ctx_.raw());
        return replaced != 0;
    }
    bool remove_model(uint64_t key) { int32_t r = 0; check(sb_mm_remove_model(h_, key, &r), ctx_.raw()); return r != 0; }
    void update_camera_with_pod(const SbCameraPod& p) { check(sb_mm_update_camera_with_pod(h_, &p), ctx_.raw()); }
    void update_model_transform_with_pod(uint64_t key, const SbModelTransformPod& p) { check(sb_mm_update_model_transform_with_pod(h_, key, &p), ctx_.raw()); }
    void render(void* stream, const SbTarget& t, const std::vector<uint64_t>& keys) { check(sb_mm_render(h_, stream, &t, keys.data(), (uint32_t)keys.size()), ctx_.raw()); }
private:
    Context& ctx_;
    int32_t sh_, cov_;
    SbMultiModelViewer* h_ = nullptr;
};

// Preprocessor<G, ()> / RadixSorter<()> / Renderer<G, ()> on caller-owned device buffers
// (reference src/preprocessor.rs:370-450, src/radix_sorter.rs:71-96, src/renderer.rs:242-356)
class Preprocessor {
public:
    Preprocessor(Context& ctx, uint64_t n, int32_t sh_fmt = SB_SH_SINGLE, int32_t cov_fmt = SB_COV_SINGLE) : ctx_(ctx) { check(sb_preprocessor_create(ctx.raw(), sh_fmt, cov_fmt, n, &h_), ctx.raw()); }
    ~Preprocessor() { sb_preprocessor_destroy(h_); }
    void preprocess(void* stream, const SbPreprocessorBindGroup& bg, uint32_t gaussian_count) { check(sb_preprocessor_preprocess(h_, stream, &bg, gaussian_count), ctx_.raw()); }
private:
    Context& ctx_;
    SbPreprocessor* h_ = nullptr;
};

class RadixSorter {
public:
    RadixSorter(Context& ctx, uint32_t capacity) : ctx_(ctx), cap_(capacity) { check(sb_sorter_create(ctx.raw(), capacity, &h_), ctx.raw()); }
    ~RadixSorter() { sb_sorter_destroy(h_); }
    // sort(encoder, bind_groups, args): the count is IndirectArgsBuffer.instance_count, read on the device
    void sort(void* stream, float* d_depth_keys, uint32_t* d_indices, const SbDrawIndirectArgs* d_indirect_args) {
        check(sb_sorter_sort(h_, stream, reinterpret_cast<uint32_t*>(d_depth_keys), d_indices, &d_indirect_args->instance_count, cap_, 0, 32), ctx_.raw());
    }
private:
    Context& ctx_;
    uint32_t cap_;
    SbRadixSorter* h_ = nullptr;
};

class Renderer {
public:
    Renderer(Context& ctx, int32_t texture_format, uint64_t n, int32_t sh_fmt = SB_SH_SINGLE, int32_t cov_fmt = SB_COV_SINGLE) : ctx_(ctx) { check(sb_renderer_create(ctx.raw(), sh_fmt, cov_fmt, texture_format, n, &h_), ctx.raw()); }
    ~Renderer() { sb_renderer_destroy(h_); }
    void render(void* stream, const SbTarget& t, const SbRendererBindGroup& bg, const SbDrawIndirectArgs* d_indirect_args) { check(sb_renderer_render(h_, stream, &bg, &t, d_indirect_args, nullptr, 0), ctx_.raw()); }
    void render_with_pass(void* stream, const SbTarget& t, const SbDepthAttachment* depth, const SbRendererBindGroup& bg, const SbDrawIndirectArgs* d_indirect_args) { check(sb_renderer_render(h_, stream, &bg, &t, d_indirect_args, depth, 1), ctx_.raw()); }
private:
    Context& ctx_;
    SbRenderer* h_ = nullptr;
};

}  // namespace splat_b200
