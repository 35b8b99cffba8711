// sb_select.cu — K7: viewport selection evaluation.
//
// Replaces selection::viewport::main (src/shader/selection/viewport.wesl:37-69) fed by the
// rectangle mask texture (src/shader/selection/viewport_texture_rectangle.wesl): a Gaussian is
// selected iff its centre passes cull() and the mask texel under it is set.  The rectangle is
// evaluated analytically: texel (ix,iy) is set iff its centre lies in [x0,x1) x [y0,y1).
// The brush mask (src/shader/selection/viewport_texture_brush.wesl: two discs + a quad per stroke segment,
// i.e. a capsule) is evaluated analytically too: a texel is set iff its centre is within `radius` of a segment
// of the stroke polyline; strokes accumulate like they do in the reference's mask texture (op = union).
// One warp owns one 32-bit selection word, so no atomics are needed.
#include "sb_internal.h"

namespace sb {

namespace {

// texel under a Gaussian's centre (viewport.wesl:44-60); false when the centre is culled or off the texture
__device__ __forceinline__ bool centre_texel(const uint8_t* __restrict__ gaussians, uint32_t g, uint32_t stride, const Uniforms& u,
                                             float& px, float& py) {
    const float4 head = __ldg(reinterpret_cast<const float4*>(gaussians + (size_t)g * stride));
    float world[3], clip[4];
#pragma unroll
    for (int i = 0; i < 3; i++)
        world[i] = sadd(sadd(sadd(smul(u.model[i], head.x), smul(u.model[4 + i], head.y)), smul(u.model[8 + i], head.z)), u.model[12 + i]);
#pragma unroll
    for (int i = 0; i < 4; i++)
        clip[i] = sadd(sadd(sadd(smul(u.pv[i], world[0]), smul(u.pv[4 + i], world[1])), smul(u.pv[8 + i], world[2])), u.pv[12 + i]);
    const float nx = sdiv(clip[0], clip[3]), ny = sdiv(clip[1], clip[3]), nz = sdiv(clip[2], clip[3]);
    if (!((nx >= -1.0f && ny >= -1.0f && nz >= 0.0f) && (nx <= 1.0f && ny <= 1.0f && nz <= 1.0f))) return false;
    // ndc_to_camera_texture (camera.wesl:18-20) then vec2<i32>() truncation (viewport.wesl:60)
    const float tx = smul(smul(sadd(smul(nx, 1.0f), 1.0f), u.size[0]), 0.5f);
    const float ty = smul(smul(sadd(smul(ny, -1.0f), 1.0f), u.size[1]), 0.5f);
    const int ix = (int)tx, iy = (int)ty;
    if (!(ix >= 0 && iy >= 0 && ix < (int)u.size[0] && iy < (int)u.size[1])) return false;
    px = (float)ix + 0.5f;
    py = (float)iy + 0.5f;
    return true;
}

// squared distance from p to the segment a-b, every step individually rounded (mirrors so_seg_dist2 in the oracle)
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float ex = ssub(bx, ax), ey = ssub(by, ay), wx = ssub(px, ax), wy = ssub(py, ay);
    const float len2 = sadd(smul(ex, ex), smul(ey, ey));
    float t = 0.0f;
    if (len2 > 0.0f) t = fminf(fmaxf(sdiv(sadd(smul(wx, ex), smul(wy, ey)), len2), 0.0f), 1.0f);
    const float dx = ssub(wx, smul(t, ex)), dy = ssub(wy, smul(t, ey));
    return sadd(smul(dx, dx), smul(dy, dy));
}

struct BrushStroke {
    float xy[2 * SB_BRUSH_MAX_POINTS];
    uint32_t n_points;
    float radius;
};

__global__ void __launch_bounds__(256) select_brush_kernel(const uint8_t* __restrict__ gaussians, uint32_t n, uint32_t stride,
                                                           const __grid_constant__ Uniforms u, const __grid_constant__ BrushStroke b,
                                                           int accumulate, uint32_t* __restrict__ words) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    float px, py;
    if (g < n && centre_texel(gaussians, g, stride, u, px, py)) {
        const float r2 = smul(b.radius, b.radius);
        if (b.n_points == 1) sel = seg_dist2(px, py, b.xy[0], b.xy[1], b.xy[0], b.xy[1]) <= r2;
        for (uint32_t i = 0; i + 1 < b.n_points; i++)
            sel = sel || seg_dist2(px, py, b.xy[2 * i], b.xy[2 * i + 1], b.xy[2 * i + 2], b.xy[2 * i + 3]) <= r2;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, sel);
    if ((threadIdx.x & 31u) == 0 && g < n) words[g >> 5] = accumulate ? (words[g >> 5] | bits) : bits;
}

__global__ void __launch_bounds__(256) select_rect_kernel(const uint8_t* __restrict__ gaussians, uint32_t n, uint32_t stride,
                                                          const __grid_constant__ Uniforms u, float x0, float y0, float x1, float y1,
                                                          uint32_t* __restrict__ words) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    bool sel = false;
    if (g < n) {
        const float4 head = __ldg(reinterpret_cast<const float4*>(gaussians + (size_t)g * stride));
        float world[3], clip[4];
#pragma unroll
        for (int i = 0; i < 3; i++)
            world[i] = sadd(sadd(sadd(smul(u.model[i], head.x), smul(u.model[4 + i], head.y)), smul(u.model[8 + i], head.z)),
                            u.model[12 + i]);
#pragma unroll
        for (int i = 0; i < 4; i++)
            clip[i] = sadd(sadd(sadd(smul(u.pv[i], world[0]), smul(u.pv[4 + i], world[1])), smul(u.pv[8 + i], world[2])),
                           u.pv[12 + i]);
        const float nx = sdiv(clip[0], clip[3]), ny = sdiv(clip[1], clip[3]), nz = sdiv(clip[2], clip[3]);
        const bool culled = !((nx >= -1.0f && ny >= -1.0f && nz >= 0.0f) && (nx <= 1.0f && ny <= 1.0f && nz <= 1.0f));
        if (!culled) {
            // ndc_to_camera_texture (camera.wesl:18-20) then vec2<i32>() truncation (viewport.wesl:60)
            const float tx = smul(smul(sadd(smul(nx, 1.0f), 1.0f), u.size[0]), 0.5f);
            const float ty = smul(smul(sadd(smul(ny, -1.0f), 1.0f), u.size[1]), 0.5f);
            const int ix = (int)tx, iy = (int)ty;
            if (ix >= 0 && iy >= 0 && ix < (int)u.size[0] && iy < (int)u.size[1]) {
                const float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
                sel = px >= x0 && px < x1 && py >= y0 && py < y1;
            }
        }
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, sel);
    if ((threadIdx.x & 31u) == 0 && g < n) words[g >> 5] = bits;
}

// K8: the colour-override edit of the editor crate's BasicSelectionModifier in the shape the
// reference's own test uses it (tests/e2e/selection.rs:54-116): NonDestructiveModifier keeps the source colours and, on every
// apply, rewrites the colour word of each Gaussian — selected: the override rgb (pack4x8unorm) with the source alpha
// scaled by `alpha`; not selected: the source colour again.  `orig` is the snapshot of the colour words (byte 12 of a pod).
__global__ void __launch_bounds__(256) snapshot_colors_kernel(const uint8_t* __restrict__ gaussians, uint32_t n, uint32_t stride,
                                                              uint32_t* __restrict__ orig) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) orig[g] = *reinterpret_cast<const uint32_t*>(gaussians + (size_t)g * stride + 12);
}

__device__ __forceinline__ uint32_t unorm8_pack(float x) {  // pack4x8unorm: round(clamp(x, 0, 1) * 255)
    return (uint32_t)__float2int_rn(__fmul_rn(fminf(fmaxf(x, 0.0f), 1.0f), 255.0f));
}

__global__ void __launch_bounds__(256) rgb_override_kernel(uint8_t* __restrict__ gaussians, uint32_t n, uint32_t stride,
                                                           const uint32_t* __restrict__ orig, const uint32_t* __restrict__ selection,
                                                           float r, float g_, float b, float alpha) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint32_t c = orig[g];
    if (selection && ((selection[g >> 5] >> (g & 31u)) & 1u)) {
        const float a = __fmul_rn(__fdiv_rn((float)(c >> 24), 255.0f), alpha);
        c = unorm8_pack(r) | (unorm8_pack(g_) << 8) | (unorm8_pack(b) << 16) | (unorm8_pack(a) << 24);
    }
    *reinterpret_cast<uint32_t*>(gaussians + (size_t)g * stride + 12) = c;
}

// The editor's BasicColorModifiers on the selected Gaussians (`basic_color_modifiers_buffer.update(queue, rgb_or_hsv, alpha, contrast,
// exposure, gamma)`, tests/e2e/selection.rs:80-92; the arithmetic itself lives in the un-vendored wgpu-3dgs-editor and is RESTATED
// here from its documented meaning, not from its source): on the source colour of a selected Gaussian, in this order —
//   rgb override, or HSV: hue += dh (turns, wraps), saturation *= ks, value *= kv (each clamped to [0,1]);
//   contrast:  c = (c - 0.5) * (1 + contrast) + 0.5;   exposure:  c *= 2^exposure;   gamma:  c = max(c, 0)^gamma;
//   alpha *= alpha scale;  then pack4x8unorm.  Neutral values (0, 1, 1 / 1 / 0 / 0 / 1) leave the colour word unchanged.
// Every step but 2^exposure (done once on the host) and the power is an individually rounded f32 operation.
struct ColorMods {
    int rgb_override;
    float p0, p1, p2;  // override rgb, or (dh, ks, kv)
    float alpha, contrast1, gain, gamma;  // contrast1 = 1 + contrast, gain = 2^exposure
};

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

__global__ void __launch_bounds__(256) basic_color_kernel(uint8_t* __restrict__ gaussians, uint32_t n, uint32_t stride,
                                                          const uint32_t* __restrict__ orig, const uint32_t* __restrict__ selection,
                                                          const ColorMods m) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint32_t c = orig[g];
    if (selection && ((selection[g >> 5] >> (g & 31u)) & 1u)) {
        float r = __fdiv_rn((float)(c & 255u), 255.0f), gg = __fdiv_rn((float)((c >> 8) & 255u), 255.0f),
              b = __fdiv_rn((float)((c >> 16) & 255u), 255.0f);
        const float a = __fmul_rn(__fdiv_rn((float)(c >> 24), 255.0f), m.alpha);
        if (m.rgb_override) {
            r = m.p0; gg = m.p1; b = m.p2;
        } else {
            const float mx = fmaxf(r, fmaxf(gg, b)), mn = fminf(r, fminf(gg, b)), d = __fsub_rn(mx, mn);
            float h6 = 0.0f;
            if (d > 0.0f) {
                if (mx == r) h6 = __fdiv_rn(__fsub_rn(gg, b), d);
                else if (mx == gg) h6 = __fadd_rn(__fdiv_rn(__fsub_rn(b, r), d), 2.0f);
                else h6 = __fadd_rn(__fdiv_rn(__fsub_rn(r, gg), d), 4.0f);
            }
            float h = __fadd_rn(__fdiv_rn(h6, 6.0f), m.p0);
            h = __fsub_rn(h, floorf(h));
            const float sat = clamp01(__fmul_rn(mx > 0.0f ? __fdiv_rn(d, mx) : 0.0f, m.p1));
            const float val = clamp01(__fmul_rn(mx, m.p2));
            const float hh = __fmul_rn(h, 6.0f), vs = __fmul_rn(val, sat);
            float out[3];
            const float ns[3] = {5.0f, 3.0f, 1.0f};
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float k = __fadd_rn(ns[i], hh);
                if (k >= 6.0f) k = __fsub_rn(k, 6.0f);
                const float t = fmaxf(0.0f, fminf(fminf(k, __fsub_rn(4.0f, k)), 1.0f));
                out[i] = __fsub_rn(val, __fmul_rn(vs, t));
            }
            r = out[0]; gg = out[1]; b = out[2];
        }
        float ch[3] = {r, gg, b};
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float x = __fadd_rn(__fmul_rn(__fsub_rn(ch[i], 0.5f), m.contrast1), 0.5f);
            x = __fmul_rn(x, m.gain);
            if (m.gamma != 1.0f) x = powf(fmaxf(x, 0.0f), m.gamma);
            ch[i] = x;
        }
        c = unorm8_pack(ch[0]) | (unorm8_pack(ch[1]) << 8) | (unorm8_pack(ch[2]) << 16) | (unorm8_pack(a) << 24);
    }
    *reinterpret_cast<uint32_t*>(gaussians + (size_t)g * stride + 12) = c;
}

}  // namespace

cudaError_t launch_basic_color_modifiers(uint8_t* gaussians, uint32_t n, uint32_t stride, const uint32_t* orig, const uint32_t* selection,
                                         int rgb_override, const float rgb_or_hsv[3], float alpha, float contrast, float exposure,
                                         float gamma, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    ColorMods m;
    m.rgb_override = rgb_override;
    m.p0 = rgb_or_hsv[0]; m.p1 = rgb_or_hsv[1]; m.p2 = rgb_or_hsv[2];
    m.alpha = alpha;
    m.contrast1 = 1.0f + contrast;
    m.gain = exp2f(exposure);
    m.gamma = gamma;
    basic_color_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gaussians, n, stride, orig, selection, m);
    return cudaGetLastError();
}

cudaError_t launch_select_rect(const uint8_t* gaussians, uint32_t n, uint32_t stride, const Uniforms& u, float x0, float y0, float x1,
                               float y1, uint32_t* words, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    select_rect_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gaussians, n, stride, u, x0, y0, x1, y1, words);
    return cudaGetLastError();
}

cudaError_t launch_select_brush(const uint8_t* gaussians, uint32_t n, uint32_t stride, const Uniforms& u, const float* points_xy,
                                uint32_t n_points, float radius, int accumulate, uint32_t* words, cudaStream_t stream) {
    if (n == 0 || n_points == 0) return cudaSuccess;
    if (n_points > SB_BRUSH_MAX_POINTS) return cudaErrorInvalidValue;
    BrushStroke b;
    for (uint32_t i = 0; i < 2 * n_points; i++) b.xy[i] = points_xy[i];
    for (uint32_t i = 2 * n_points; i < 2 * SB_BRUSH_MAX_POINTS; i++) b.xy[i] = 0.0f;
    b.n_points = n_points;
    b.radius = radius;
    select_brush_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gaussians, n, stride, u, b, accumulate, words);
    return cudaGetLastError();
}

cudaError_t launch_snapshot_colors(const uint8_t* gaussians, uint32_t n, uint32_t stride, uint32_t* orig, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    snapshot_colors_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gaussians, n, stride, orig);
    return cudaGetLastError();
}

cudaError_t launch_rgb_override(uint8_t* gaussians, uint32_t n, uint32_t stride, const uint32_t* orig, const uint32_t* selection,
                                const float rgb[3], float alpha, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    rgb_override_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gaussians, n, stride, orig, selection, rgb[0], rgb[1], rgb[2], alpha);
    return cudaGetLastError();
}

}  // namespace sb
