// sb_preprocess.cu — K1: the fused Preprocessor.
//
// Replaces, in ONE launch, the reference's three compute dispatches
//   pre / main / post                       (src/shader/preprocess.wesl:52-126)
// and the per-vertex work that vert_main repeats 6x per splat
//   color() + cov2d_axes() + clip position  (src/shader/render.wesl:58-130, utils.wesl:25-135).
//
// Shape: persistent CTAs (one per SM).  A producer warp streams tiles of T consecutive pods
// from HBM into a shared-memory ring with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP)
// signalled on mbarriers; T consumer threads each own one Gaussian of the tile.  Tiles are
// handed out by a global ticket so the visible-splat compaction can be ORDER PRESERVING
// (decoupled look-back over per-tile visible counts): slot order == ascending Gaussian index,
// the canonical order of the reference's racy atomicAdd compaction.
// The CTA-wide scan uses SPLIT-PHASE mbarriers (arrive right after the cull decision, wait
// only after the per-splat vertex work), so no warp ever blocks on bar.sync: round-1 ncu showed
// 53% of issue stalls on the two blocking barriers this replaces.
//
// Bit-exact artefacts (visible mask, count, keys, indirect args) use strict f32 intrinsics
// (__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn) in the order fixed by the oracle contract.
#include <cuda_fp16.h>

#include "sb_internal.h"

namespace sb {

namespace {

__host__ __device__ constexpr int sh_bytes(int sh) { return sh == SB_SH_SINGLE ? 180 : sh == SB_SH_HALF ? 92 : sh == SB_SH_NORM8 ? 52 : 0; }
__host__ __device__ constexpr int cov_bytes(int cov) { return cov == SB_COV_SINGLE ? 24 : cov == SB_COV_HALF ? 12 : 28; }
__host__ __device__ constexpr int pod_stride(int sh, int cov) { return (16 + sh_bytes(sh) + cov_bytes(cov) + 15) & ~15; }

// Consumer threads per CTA (= pods per tile) and ring depth, chosen so the ring fits 227 KB.
__host__ __device__ constexpr int tile_records(int) { return 512; }  // 768 for the small pods was measured: no gain
__host__ __device__ constexpr int ring_stages(int stride) {
    int s = (226 * 1024) / (tile_records(stride) * stride);
    return s < 2 ? 2 : (s > 4 ? 4 : s);
}

__device__ __forceinline__ float half_bits_to_float(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)h)); }

// gaussian_unpack_cov3d (core WESL, external): [xx, xy, xz, yy, yz, zz]
template <int SH, int COV>
__device__ __forceinline__ void unpack_cov3d(const uint8_t* rec, float cov[6]) {
    const uint8_t* c = rec + 16 + sh_bytes(SH);
    if constexpr (COV == SB_COV_SINGLE) {
        const float* f = reinterpret_cast<const float*>(c);  // 4-byte aligned
#pragma unroll
        for (int k = 0; k < 6; k++) cov[k] = f[k];
    } else if constexpr (COV == SB_COV_HALF) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(c);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t v = w[k];
            cov[2 * k] = half_bits_to_float(v & 0xffffu);
            cov[2 * k + 1] = half_bits_to_float(v >> 16);
        }
    } else {  // rot (xyzw) + scale: cov3d = (R S)(R S)^T, strict order of the oracle
        const float* f = reinterpret_cast<const float*>(c);
        float x = f[0], y = f[1], z = f[2], w = f[3];
        float s0 = f[4], s1 = f[5], s2 = f[6];
        float x2 = sadd(x, x), y2 = sadd(y, y), z2 = sadd(z, z);
        float xx = smul(x, x2), xy = smul(x, y2), xz = smul(x, z2);
        float yy = smul(y, y2), yz = smul(y, z2), zz = smul(z, z2);
        float wx = smul(w, x2), wy = smul(w, y2), wz = smul(w, z2);
        // R columns
        float r00 = ssub(1.0f, sadd(yy, zz)), r01 = sadd(xy, wz), r02 = ssub(xz, wy);
        float r10 = ssub(xy, wz), r11 = ssub(1.0f, sadd(xx, zz)), r12 = sadd(yz, wx);
        float r20 = sadd(xz, wy), r21 = ssub(yz, wx), r22 = ssub(1.0f, sadd(xx, yy));
        // m[row][col] = R[col][row] * s[col]
        float m00 = smul(r00, s0), m01 = smul(r10, s1), m02 = smul(r20, s2);
        float m10 = smul(r01, s0), m11 = smul(r11, s1), m12 = smul(r21, s2);
        float m20 = smul(r02, s0), m21 = smul(r12, s1), m22 = smul(r22, s2);
        cov[0] = sadd(sadd(smul(m00, m00), smul(m01, m01)), smul(m02, m02));
        cov[1] = sadd(sadd(smul(m00, m10), smul(m01, m11)), smul(m02, m12));
        cov[2] = sadd(sadd(smul(m00, m20), smul(m01, m21)), smul(m02, m22));
        cov[3] = sadd(sadd(smul(m10, m10), smul(m11, m11)), smul(m12, m12));
        cov[4] = sadd(sadd(smul(m10, m20), smul(m11, m21)), smul(m12, m22));
        cov[5] = sadd(sadd(smul(m20, m20), smul(m21, m21)), smul(m22, m22));
    }
}

// cov2d (utils.wesl:25-48) + cov2d_axes (utils.wesl:54-79), strict f32.
__device__ __forceinline__ void cov2d_axes(const Uniforms& u, float px, float py, float pz, const float cov[6],
                                           float std_dev, float axes[4]) {
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        t[i] = sadd(sadd(sadd(smul(u.vm[i], px), smul(u.vm[4 + i], py)), smul(u.vm[8 + i], pz)), u.vm[12 + i]);
    float tz2 = smul(t[2], t[2]);
    float j00 = sdiv(u.focal[0], t[2]);
    float j02 = sdiv(-smul(u.focal[0], t[0]), tz2);
    float j11 = sdiv(u.focal[1], t[2]);
    float j12 = sdiv(-smul(u.focal[1], t[1]), tz2);
    float jw0[3], jw1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        jw0[c] = sadd(smul(j00, u.w[c * 3 + 0]), smul(j02, u.w[c * 3 + 2]));
        jw1[c] = sadd(smul(j11, u.w[c * 3 + 1]), smul(j12, u.w[c * 3 + 2]));
    }
    float t0[3], t1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        t0[c] = sadd(sadd(smul(jw0[0], u.sr[c * 3 + 0]), smul(jw0[1], u.sr[c * 3 + 1])), smul(jw0[2], u.sr[c * 3 + 2]));
        t1[c] = sadd(sadd(smul(jw1[0], u.sr[c * 3 + 0]), smul(jw1[1], u.sr[c * 3 + 1])), smul(jw1[2], u.sr[c * 3 + 2]));
    }
    const float v[3][3] = {{cov[0], cov[1], cov[2]}, {cov[1], cov[3], cov[4]}, {cov[2], cov[4], cov[5]}};
    float tv0[3], tv1[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        tv0[c] = sadd(sadd(smul(t0[0], v[0][c]), smul(t0[1], v[1][c])), smul(t0[2], v[2][c]));
        tv1[c] = sadd(sadd(smul(t1[0], v[0][c]), smul(t1[1], v[1][c])), smul(t1[2], v[2][c]));
    }
    float ca = sadd(sadd(smul(tv0[0], t0[0]), smul(tv0[1], t0[1])), smul(tv0[2], t0[2]));
    float cb = sadd(sadd(smul(tv1[0], t0[0]), smul(tv1[1], t0[1])), smul(tv1[2], t0[2]));
    float cc = sadd(sadd(smul(tv1[0], t1[0]), smul(tv1[1], t1[1])), smul(tv1[2], t1[2]));

    float mid = smul(0.5f, sadd(ca, cc));
    float hx = smul(0.5f, ssub(ca, cc));
    float radius = ssqrt(sadd(smul(hx, hx), smul(cb, cb)));
    float major_lambda = sadd(mid, radius);
    float minor_lambda = ssub(mid, radius);
    if (minor_lambda < 0.0f) {
        axes[0] = axes[1] = axes[2] = axes[3] = 0.0f;
        return;
    }
    float dx = cb, dy = ssub(major_lambda, ca);
    float ddx, ddy;
    if (dx == 0.0f && dy == 0.0f) {
        ddx = 0.0f;
        ddy = 1.0f;
    } else {
        float l = ssqrt(sadd(smul(dx, dx), smul(dy, dy)));
        ddx = sdiv(dx, l);
        ddy = sdiv(dy, l);
    }
    float major_len = fminf(smul(std_dev, ssqrt(major_lambda)), 1024.0f);
    float minor_len = fminf(smul(std_dev, ssqrt(minor_lambda)), 1024.0f);
    axes[0] = smul(major_len, ddx);
    axes[1] = smul(major_len, ddy);
    axes[2] = smul(minor_len, ddy);
    axes[3] = smul(minor_len, -ddx);
}

// utils.wesl:18-20
__device__ __forceinline__ bool cull(float x, float y, float z) {
    return !((x >= -1.0f && y >= -1.0f && z >= 0.0f) && (x <= 1.0f && y <= 1.0f && z <= 1.0f));
}

// u8 / 255 correctly rounded without the IEEE-division sequence: one Newton step on x * (1/255)
// (the identity is verified exhaustively for x = 0..255 by so_unorm8_newton_mismatches(), tests/test_oracle_golden.py).
__device__ __forceinline__ float unorm8(uint32_t x) {
    const float fx = (float)x;
    const float q = __fmul_rn(fx, 0.00392156886f);
    const float r = __fmaf_rn(-q, 255.0f, fx);
    return __fmaf_rn(r, 0.00392156886f, q);
}

// gaussian_unpack_sh(g, i) for i in 0..14 -> rgb
template <int SH>
__device__ __forceinline__ void unpack_sh(const uint8_t* rec, int i, float& r, float& g, float& b, float mn, float mx) {
    const uint8_t* q = rec + 16;
    if constexpr (SH == SB_SH_SINGLE) {
        const float* f = reinterpret_cast<const float*>(q);
        r = f[i * 3 + 0]; g = f[i * 3 + 1]; b = f[i * 3 + 2];
    } else if constexpr (SH == SB_SH_HALF) {
        const unsigned short* h = reinterpret_cast<const unsigned short*>(q);
        r = half_bits_to_float(h[i * 3 + 0]); g = half_bits_to_float(h[i * 3 + 1]); b = half_bits_to_float(h[i * 3 + 2]);
    } else if constexpr (SH == SB_SH_NORM8) {
        const uint8_t* c = q + 4;
        const float tr = unorm8(c[i * 3 + 0]), tg = unorm8(c[i * 3 + 1]), tb = unorm8(c[i * 3 + 2]);
        r = sadd(smul(mn, ssub(1.0f, tr)), smul(mx, tr));
        g = sadd(smul(mn, ssub(1.0f, tg)), smul(mx, tg));
        b = sadd(smul(mn, ssub(1.0f, tb)), smul(mx, tb));
    } else {
        r = g = b = 0.0f;
    }
}

// view_color (utils.wesl:82-135).  Contract (matches the oracle bit for bit): basis factors are
// individually rounded products, each degree's sum is one fma chain.
template <int SH>
__device__ __forceinline__ void view_color(const Uniforms& u, const uint8_t* rec, uint32_t packed_color, float x, float y,
                                           float z, float rgb[3]) {
    const float sh_c1 = 0.4886025f;
    const float c2_0 = 1.0925484f, c2_1 = -1.0925484f, c2_2 = 0.3153916f, c2_3 = -1.0925484f, c2_4 = 0.5462742f;
    const float c3_0 = -0.5900436f, c3_1 = 2.8906114f, c3_2 = -0.4570458f, c3_3 = 0.3731763f, c3_4 = -0.4570458f,
                c3_5 = 1.4453057f, c3_6 = -0.5900436f;
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; c++) res[c] = u.no_sh0 ? 0.5f : unorm8((packed_color >> (8 * c)) & 255u);
    if (SH != SB_SH_NONE && u.sh_deg >= 1) {
        float mn = 0.0f, mx = 0.0f;
        if constexpr (SH == SB_SH_NORM8) {
            uint32_t mm = *reinterpret_cast<const uint32_t*>(rec + 16);
            mn = half_bits_to_float(mm & 0xffffu);
            mx = half_bits_to_float(mm >> 16);
        }
        float s[15][3];
#pragma unroll
        for (int i = 0; i < 15; i++) unpack_sh<SH>(rec, i, s[i][0], s[i][1], s[i][2], mn, mx);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float t = __fmaf_rn(s[1][c], z, -smul(s[0][c], y));
            t = __fmaf_rn(-s[2][c], x, t);
            res[c] = __fmaf_rn(sh_c1, t, res[c]);
        }
        if (u.sh_deg >= 2) {
            const float xx = smul(x, x), yy = smul(y, y), zz = smul(z, z), xy = smul(x, y), yz = smul(y, z), xz = smul(x, z);
            const float b0 = smul(c2_0, xy), b1 = smul(c2_1, yz), b2 = smul(c2_2, ssub(ssub(smul(2.0f, zz), xx), yy)),
                        b3 = smul(c2_3, xz), b4 = smul(c2_4, ssub(xx, yy));
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float acc = smul(b0, s[3][c]);
                acc = __fmaf_rn(b1, s[4][c], acc);
                acc = __fmaf_rn(b2, s[5][c], acc);
                acc = __fmaf_rn(b3, s[6][c], acc);
                acc = __fmaf_rn(b4, s[7][c], acc);
                res[c] = sadd(res[c], acc);
            }
            if (u.sh_deg >= 3) {
                const float zz4 = ssub(ssub(smul(4.0f, zz), xx), yy);
                const float d0 = smul(smul(c3_0, y), ssub(smul(3.0f, xx), yy));
                const float d1 = smul(smul(c3_1, xy), z);
                const float d2 = smul(smul(c3_2, y), zz4);
                const float d3 = smul(smul(c3_3, z), ssub(ssub(smul(2.0f, zz), smul(3.0f, xx)), smul(3.0f, yy)));
                const float d4 = smul(smul(c3_4, x), zz4);
                const float d5 = smul(smul(c3_5, z), ssub(xx, yy));
                const float d6 = smul(smul(c3_6, x), ssub(xx, smul(3.0f, yy)));
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float acc = smul(d0, s[8][c]);
                    acc = __fmaf_rn(d1, s[9][c], acc);
                    acc = __fmaf_rn(d2, s[10][c], acc);
                    acc = __fmaf_rn(d3, s[11][c], acc);
                    acc = __fmaf_rn(d4, s[12][c], acc);
                    acc = __fmaf_rn(d5, s[13][c], acc);
                    acc = __fmaf_rn(d6, s[14][c], acc);
                    res[c] = sadd(res[c], acc);
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) rgb[c] = fmaxf(res[c], 0.0f);
}

// Tile bbox of the alive region, with a 0.01 px safety margin against f32 rounding.
__device__ __forceinline__ void tile_bbox(const Uniforms& u, float cx, float cy, float ex, float ey, uint32_t& tmin,
                                          uint32_t& tmax) {
    const float W = (float)u.width, H = (float)u.height;
    float x0f = ceilf(cx - ex - 0.51f), x1f = floorf(cx + ex - 0.49f);
    float y0f = ceilf(cy - ey - 0.51f), y1f = floorf(cy + ey - 0.49f);
    bool ok = (x1f >= 0.0f) && (y1f >= 0.0f) && (x0f <= W - 1.0f) && (y0f <= H - 1.0f) && (x0f <= x1f) && (y0f <= y1f);
    if (!ok) {  // also catches NaN
        tmin = 1u | (1u << 16);
        tmax = 0u;
        return;
    }
    uint32_t x0 = (uint32_t)fmaxf(x0f, 0.0f), x1 = (uint32_t)fminf(x1f, W - 1.0f);
    uint32_t y0 = (uint32_t)fmaxf(y0f, 0.0f), y1 = (uint32_t)fminf(y1f, H - 1.0f);
    tmin = (x0 / kTile) | ((y0 / kTile) << 16);
    tmax = (x1 / kTile) | ((y1 / kTile) << 16);
}

__device__ __forceinline__ bool finite4(float a, float b, float c, float d) {
    return isfinite(a) && isfinite(b) && isfinite(c) && isfinite(d);
}

// ---- per-Gaussian work shared by the fused preprocess kernel and the standalone Renderer's vertex stage ----------
// Phase 1 (preprocess.wesl:60-106): selection, projection, frustum cull (+ the border cull for centres just
// outside).  `rec` points at the pod — in shared memory (K1) or global memory (vertex_kernel).
struct CullOut {
    float4 head;
    float world[3];
    float nx, ny, nz, cw;
    float axes[4];
};
template <int SH, int COV>
__device__ __forceinline__ bool cull_gaussian(const PreParams& p, const Uniforms& u, uint32_t g, const uint8_t* rec, float sd_size,
                                              bool vis, CullOut& o) {
    float4 head = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vis) head = *reinterpret_cast<const float4*>(rec);
    // selection: preprocess.wesl:68-78
    if (p.selection != nullptr && vis) {
        const uint32_t gi = g + p.index_base;  // (a slice of the model: the mask is indexed by the Gaussian's index in the whole model)
        const uint32_t word = __ldg(&p.selection[gi >> 5]);
        const bool bit = (word >> (gi & 31u)) & 1u;
        const bool inverted = p.invert_selection != 0u;
        if (inverted == bit) vis = false;
    }

    // model_to_world + world_to_camera: preprocess.wesl:82-84 (strict)
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

    // cov2d_axes is needed by the border cull (centre outside) and by every survivor in
    // splat/ellipse mode: evaluate it once, here, for whoever needs it.
    const bool centre_out = cull(nx, ny, nz);
    float axes[4] = {0.f, 0.f, 0.f, 0.f};
    if (vis && (centre_out || u.mode != SB_MODE_POINT)) {
        float cov[6];
        unpack_cov3d<SH, COV>(rec, cov);
        cov2d_axes(u, head.x, head.y, head.z, cov, sd_size, axes);
    }
    if (vis && centre_out) {  // preprocess.wesl:87-99
        const float mx = sdiv(smul(axes[0], u.std_dev), u.size[0]);
        const float my = sdiv(smul(axes[1], u.std_dev), u.size[1]);
        const float ndc_major_len = ssqrt(sadd(smul(mx, mx), smul(my, my)));
        const float l = ssqrt(sadd(smul(nx, nx), smul(ny, ny)));
        const float dirx = sdiv(-nx, l), diry = sdiv(-ny, l);
        const float m = fminf(ndc_major_len, l);
        const float bx = sadd(nx, smul(m, dirx)), by = sadd(ny, smul(m, diry));
        if (cull(bx, by, nz)) vis = false;
    }

    o.head = head;
#pragma unroll
    for (int i = 0; i < 3; i++) o.world[i] = world[i];
    o.nx = nx;
    o.ny = ny;
    o.nz = nz;
    o.cw = clip[3];
#pragma unroll
    for (int i = 0; i < 4; i++) o.axes[i] = axes[i];
    return vis;
}

// Phase 2 (render.wesl:76-130): the vertex-stage work of a splat, written once per splat.  Split in two so that a strip
// render ("drops splats whose quad bbox misses its strip") can decide on the tile box BEFORE the compaction and
// evaluate the colour only for the strip's own splats.
// 2a: geometry — pixel centre, inverse axes, cull extents (with the exact alpha cut-off), tile box.
struct SplatGeom {
    float cx, cy, ax, ay, bx, by, ex, ey, a;
    uint32_t tmin, tmax;
};
__device__ __forceinline__ void splat_geometry(const Uniforms& u, const CullOut& o, SplatGeom& out) {
    const float4 head = o.head;
    const float* axes = o.axes;
    out.cx = smul(smul(sadd(o.nx, 1.0f), 0.5f), u.size[0]);
    out.cy = smul(smul(ssub(1.0f, o.ny), 0.5f), u.size[1]);
    out.a = unorm8(__float_as_uint(head.w) >> 24);
    bool valid;
    if (u.mode == SB_MODE_POINT) {  // render.wesl:92-104
        float vp[3];
#pragma unroll
        for (int i = 0; i < 3; i++)
            vp[i] = sadd(sadd(sadd(smul(u.vm[i], head.x), smul(u.vm[4 + i], head.y)), smul(u.vm[8 + i], head.z)),
                         u.vm[12 + i]);
        const float len = ssqrt(sadd(sadd(smul(vp[0], vp[0]), smul(vp[1], vp[1])), smul(vp[2], vp[2])));
        const float half = sdiv(smul(smul(smul(0.01f, u.gsize), 0.5f), u.size[1]), len);
        const float inv = sdiv(1.0f, half);
        out.ax = inv; out.ay = 0.0f; out.bx = 0.0f; out.by = inv;
        out.ex = half; out.ey = half;
        valid = (half > 0.0f) && isfinite(inv) && isfinite(out.cx) && isfinite(out.cy);
    } else {
        const float mm = sadd(smul(axes[0], axes[0]), smul(axes[1], axes[1]));
        const float nn = sadd(smul(axes[2], axes[2]), smul(axes[3], axes[3]));
        out.ax = sdiv(smul(2.0f, axes[0]), mm);
        out.ay = -sdiv(smul(2.0f, axes[1]), mm);
        out.bx = sdiv(smul(2.0f, axes[2]), nn);
        out.by = -sdiv(smul(2.0f, axes[3]), nn);
        const float hs = smul(0.5f, u.std_dev);
        out.ex = smul(hs, ssqrt(sadd(smul(axes[0], axes[0]), smul(axes[2], axes[2]))));
        out.ey = smul(hs, ssqrt(sadd(smul(axes[1], axes[1]), smul(axes[3], axes[3]))));
        valid = finite4(out.ax, out.ay, out.bx, out.by) && finite4(out.cx, out.cy, out.ex, out.ey);
        if (u.cut_k > 0.0f) {  // exact alpha cut-off (sb_common.cuh): splat mode on a unorm8 target
            const float rc2 = logf(out.a * u.cut_k) + kAlphaCutMargin;  // a = 0 -> -inf
            if (!(rc2 > 0.0f)) {
                valid = false;  // a < kAlphaCut: every blend of this splat is the identity
            } else if (rc2 < u.std_dev * u.std_dev) {
                const float s = sqrtf(rc2) / u.std_dev * 1.0001f;
                out.ex *= s;
                out.ey *= s;
            }
        }
    }
    if (valid) {
        tile_bbox(u, out.cx, out.cy, out.ex, out.ey, out.tmin, out.tmax);
    } else {
        out.tmin = 1u | (1u << 16);
        out.tmax = 0u;
        out.ex = out.ey = 0.0f;
    }
}

// 2b: colour (render.wesl:58-73) + the record / tile-box stores.
template <int SH, int COV>
__device__ __forceinline__ void emit_splat(const PreParams& p, const Uniforms& u, uint32_t g, const uint8_t* rec, const CullOut& o,
                                           const SplatGeom& geo) {
    const float* world = o.world;
    // color(): -normalize(v) = -(v * (1/|v|))
    const float vdx = ssub(u.cam_pos[0], world[0]), vdy = ssub(u.cam_pos[1], world[1]), vdz = ssub(u.cam_pos[2], world[2]);
    float md[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        md[i] = sadd(sadd(smul(u.inv_sr[i], vdx), smul(u.inv_sr[3 + i], vdy)), smul(u.inv_sr[6 + i], vdz));
    const float inv_ml = sdiv(1.0f, ssqrt(sadd(sadd(smul(md[0], md[0]), smul(md[1], md[1])), smul(md[2], md[2]))));
    float rgb[3];
    const uint32_t packed = __float_as_uint(o.head.w);
    view_color<SH>(u, rec, packed, -smul(md[0], inv_ml), -smul(md[1], inv_ml), -smul(md[2], inv_ml), rgb);
    // Fixed-point colour attachments clamp the SOURCE colour to [0,1] before the blend equation
    // (Vulkan 1.3 spec 29.1 "Blending"; the reference renders to Rgba8Unorm, src/renderer.rs:296-300):
    // done here once per splat, which also makes the post-blend clamp redundant (d, c <= 255, alpha <= 1).
    const float cr = fminf(smul(rgb[0], u.color_scale), u.color_max);
    const float cg = fminf(smul(rgb[1], u.color_scale), u.color_max);
    const float cb = fminf(smul(rgb[2], u.color_scale), u.color_max);
    float4* dst = reinterpret_cast<float4*>(&p.recs[g]);
    dst[0] = make_float4(geo.cx, geo.cy, geo.ax, geo.bx);
    dst[1] = make_float4(geo.ay, geo.by, geo.ex, geo.ey);
    dst[2] = make_float4(cr, cg, cb, geo.a);
    *reinterpret_cast<uint2*>(&p.tboxes[g]) = make_uint2(geo.tmin, geo.tmax);
}

template <int SH, int COV, bool STRIP>
__global__ void __launch_bounds__(tile_records(pod_stride(SH, COV)) + 64, 1)
    preprocess_kernel(const __grid_constant__ PreParams p) {
    constexpr int STRIDE = pod_stride(SH, COV);
    constexpr int T = tile_records(STRIDE);
    constexpr int S = ring_stages(STRIDE);
    constexpr int NW = T / 32;
    constexpr uint32_t STAGE_BYTES = T * STRIDE;

    constexpr int R = 8;  // ring depth of the scan state (counts / base), > maximum warp drift

    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * STAGE_BYTES);  // [S][NW] chunk landed (TMA)
    uint64_t* empty_bar = full_bar + S * NW;                                     // [S][NW] chunk consumed
    uint64_t* counted_bar = empty_bar + S * NW;                                  // [R] all warp counts written
    uint64_t* based_bar = counted_bar + R;                                       // [R] tile base known
    uint32_t* tile_id = reinterpret_cast<uint32_t*>(based_bar + R);              // [S][NW] (per warp: warps drift)
    uint32_t* tile_base = tile_id + S * NW;                                         // [R]
    uint32_t* warp_counts = tile_base + R;                                       // [R][NW]
    uint32_t* ring_tile = warp_counts + R * NW;                                  // [R] tile of a ring slot (0xffffffff = end)

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const Uniforms& u = p.u;

    if (tid == 0) {
        for (int i = 0; i < S * NW; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < R; i++) {
            mbar_init(&counted_bar[i], NW);
            mbar_init(&based_bar[i], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ---------------- producer warp: ticket -> one TMA bulk copy per warp chunk (32 pods).
        // The ticket is taken only once the first chunk slot of the stage is free, so a ticketed
        // tile starts loading at once (successor tiles look back on its aggregate).  Chunks are
        // refilled as soon as their own warp has released them: ~NW copies in flight per SM.
        if (lane == 0) {
            for (uint32_t it = 0;; ++it) {
                const uint32_t s = it % S, ph = (it / S) & 1u;
                mbar_wait(&empty_bar[s * NW], ph ^ 1u);
                const uint32_t tile = atomicAdd(p.tile_counter, 1u);
                if (tile >= p.num_tiles) {
                    for (int w = 0; w < NW; w++) {
                        if (w > 0) mbar_wait(&empty_bar[s * NW + w], ph ^ 1u);
                        tile_id[s * NW + w] = tile;
                        mbar_arrive(&full_bar[s * NW + w]);
                    }
                    break;
                }
                const uint32_t first = tile * T;
                const uint32_t cnt = min((uint32_t)T, p.n - first);
                const uint8_t* src = p.gaussians + (size_t)first * STRIDE;
                uint8_t* dst = smem + s * STAGE_BYTES;
                for (int w = 0; w < NW; w++) {
                    // a warp's slot (pods AND tile id) may be rewritten only after that warp released it
                    if (w > 0) mbar_wait(&empty_bar[s * NW + w], ph ^ 1u);
                    tile_id[s * NW + w] = tile;
                    const uint32_t lo = min(cnt, (uint32_t)w * 32u), hi = min(cnt, (uint32_t)w * 32u + 32u);
                    const uint32_t bytes = (hi - lo) * STRIDE;
                    if (bytes == 0) {
                        mbar_arrive(&full_bar[s * NW + w]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s * NW + w], bytes);
                        bulk_g2s(dst + (size_t)lo * STRIDE, src + (size_t)lo * STRIDE, bytes, &full_bar[s * NW + w]);
                    }
                }
            }
        }
        return;
    }

    if (warp == NW + 1) {
        // ---------------- scan warp: publishes tile aggregates the moment a tile is counted and
        // resolves the decoupled look-backs opportunistically.  Neither ever blocks the other (or the
        // consumers): a blocking look-back in a consumer warp convoys every SM, because a tile's
        // aggregate would be published only after its predecessor's look-back had finished.
        uint32_t pub_it = 0, res_it = 0, acc = 0, my_agg = 0;
        int win = -2;  // -2: look-back of tile res_it not started
        bool pub_done = false;
        for (;;) {
            bool progress = false;
            // warp-uniform probe: lanes may observe the phase flip at different instants
            const bool counted = !pub_done && __all_sync(0xffffffffu, mbar_test_wait(&counted_bar[pub_it % R], (pub_it / R) & 1u));
            if (counted) {
                progress = true;
                const uint32_t ring = pub_it % R;
                const uint32_t tile = ring_tile[ring];
                if (tile == 0xffffffffu) {
                    pub_done = true;
                } else {
                    uint32_t c = lane < NW ? warp_counts[ring * NW + lane] : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    if (lane == 0) st_relaxed_u64(&p.tile_status[tile], (tile == 0 ? kFlagPrefix : kFlagAggregate) | c);
                    ++pub_it;
                }
            }
            if (res_it < pub_it) {
                const uint32_t ring = res_it % R;
                const uint32_t tile = ring_tile[ring];
                bool resolved = false;
                if (win == -2) {
                    acc = 0;
                    win = (int)tile - 1;
                    uint32_t c = lane < NW ? warp_counts[ring * NW + lane] : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                    my_agg = c;
                    resolved = tile == 0;
                }
                if (!resolved) {
                    const int idx = win - (int)lane;
                    const unsigned long long v = idx >= 0 ? ld_relaxed_u64(&p.tile_status[idx]) : kFlagPrefix;
                    if (!__any_sync(0xffffffffu, (v >> 32) == 0)) {
                        const uint32_t pm = __ballot_sync(0xffffffffu, (v >> 32) == 2);
                        const int first = pm ? (__ffs(pm) - 1) : 31;
                        uint32_t contrib = ((int)lane <= first) ? (uint32_t)v : 0u;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                        acc += contrib;
                        progress = true;
                        if (pm) resolved = true;
                        else win -= 32;
                    }
                }
                if (resolved) {
                    if (lane == 0) {
                        if (tile != 0) st_relaxed_u64(&p.tile_status[tile], kFlagPrefix | (unsigned long long)(acc + my_agg));
                        tile_base[ring] = acc;
                        mbar_arrive(&based_bar[ring]);
                    }
                    ++res_it;
                    win = -2;
                }
            }
            if (pub_done && res_it == pub_it) break;
            if (!progress) __nanosleep(64);  // do not steal issue slots from the consumer warps while idle
        }
        return;
    }

    // ---------------- consumers: one Gaussian per thread per tile; warps never block on each other.
    // The (index, key) writes of a tile need its base from the decoupled look-back; they are
    // deferred by three iterations so that latency hides behind the next tiles' work.
    const float sd_size = smul(u.std_dev, u.gsize);
    struct Deferred {
        bool vis = false;
        uint32_t g = 0, lane_rank = 0, tile = 0xffffffffu, ring = 0, par = 0;
        float key = 0.0f;
    };
    Deferred d1, d2, d3;  // outputs of the previous three tiles, oldest = d3
    uint32_t key_or = 0u, key_nand = 0u;  // which key bits are 1 / 0 in at least one visible key: the depth sort's digit plan

    // Writes a tile's (index, key) pairs.  Everything that depends on other warps (the warp prefix inside
    // the tile) or other CTAs (the tile base from the look-back) is read here, three tiles late.
    auto flush = [&](const Deferred& d) {
        if (d.tile == 0xffffffffu) return;
        mbar_wait(&based_bar[d.ring], d.par);  // implies counted
        const uint32_t base = tile_base[d.ring];
        const uint32_t c = lane < NW ? warp_counts[d.ring * NW + lane] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        const uint32_t warp_excl = __shfl_sync(0xffffffffu, inc - c, warp);
        if (d.vis) {
            const uint32_t slot = base + warp_excl + d.lane_rank;
            p.indices[slot] = d.g + p.index_base;
            p.keys[slot] = d.key;
        }
        if (d.tile == p.num_tiles - 1) {
            // post: preprocess.wesl:108-126 — indirect args + pad keys with 2.0
            const uint32_t v = base + total;
            const uint32_t blocks = (v + kHistoBlockKvs - 1) / kHistoBlockKvs;
            if (tid == 0) {
                p.draw_args->vertex_count = 6;
                p.draw_args->instance_count = v;
                p.draw_args->first_vertex = 0;
                p.draw_args->first_instance = 0;
                p.sort_args->x = blocks;
                p.sort_args->y = 1;
                p.sort_args->z = 1;
                *p.visible_count = v;
                if (p.visible_host) *p.visible_host = v;
            }
            const uint32_t padded = min(blocks * kHistoBlockKvs, p.keys_capacity);
            for (uint32_t i = v + tid; i < padded; i += T) p.keys[i] = 2.0f;
        }
    };

    for (uint32_t it = 0;; ++it) {
        const uint32_t s = it % S, ph = (it / S) & 1u;
        const uint32_t ring = it % R, rpar = (it / R) & 1u;
        mbar_wait(&full_bar[s * NW + warp], ph);
        const uint32_t tile = tile_id[s * NW + warp];
        if (tile >= p.num_tiles) {
            if (tid == 0) {  // tell the scan warp that this CTA has no more tiles
                ring_tile[ring] = 0xffffffffu;
                mbar_arrive_n(&counted_bar[ring], NW);
            }
            break;
        }
        if (tid == 0) ring_tile[ring] = tile;

        const uint32_t g = tile * T + tid;
        const uint8_t* rec = smem + s * STAGE_BYTES + tid * STRIDE;
        bool vis = g < p.n;
        CullOut co;
        vis = cull_gaussian<SH, COV>(p, u, g, rec, sd_size, vis, co);
        const float nz = co.nz;
        // Strip render (one frame split into screen strips over several GPUs): the full-frame cull above is what every rank
        // agrees on; a rank then keeps only the splats whose tile box meets its own tile rows, so its visible list — and the
        // sort, binning and colour work behind it — is its strip's subset.  A subset of an order-preserving compaction keeps
        // the order, so the strips reassemble the single-GPU frame bit for bit.
        SplatGeom geo;
        if constexpr (STRIP) {  // its own instantiation: the full-frame kernel keeps its register allocation and schedule
            if (vis) {
                splat_geometry(u, co, geo);
                const uint32_t y0 = geo.tmin >> 16, y1 = geo.tmax >> 16;
                vis = (geo.tmin & 0xffffu) <= (geo.tmax & 0xffffu) && y0 <= y1 && y1 >= p.strip_ty_lo && y0 <= p.strip_ty_hi;
            }
        }

        // ---- order-preserving compaction, phase 1: publish this warp's count (non-blocking)
        const uint32_t bal = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) {
            warp_counts[ring * NW + warp] = __popc(bal);
            mbar_arrive(&counted_bar[ring]);
        }
        // ---- (index, key) pairs of the tile three iterations back: its base has had time to arrive
        flush(d3);
        d3 = d2;
        d2 = d1;

        // ---- vertex-stage work for survivors (render.wesl:76-130), written once per splat
        if (vis && p.recs != nullptr) {  // a standalone Preprocessor has no record buffer
            if constexpr (!STRIP) splat_geometry(u, co, geo);
            emit_splat<SH, COV>(p, u, g, rec, co, geo);
        }

        // the pods of this chunk are no longer needed: let the producer refill the slot
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s * NW + warp]);

        d1.vis = vis;
        d1.g = g;
        d1.lane_rank = __popc(bal & lanemask_lt());
        d1.key = ssub(1.0f, nz);  // preprocess.wesl:105
        key_or |= vis ? __float_as_uint(d1.key) : 0u;
        key_nand |= vis ? ~__float_as_uint(d1.key) : 0u;
        d1.tile = tile;
        d1.ring = ring;
        d1.par = rpar;
    }
    flush(d3);
    flush(d2);
    flush(d1);
    if (p.sort_prep != nullptr) {
        key_or = __reduce_or_sync(0xffffffffu, key_or);
        key_nand = __reduce_or_sync(0xffffffffu, key_nand);
        if (lane == 0) {
            if (key_or) atomicOr(&p.sort_prep[kSortPrepOr], key_or);
            if (key_nand) atomicOr(&p.sort_prep[kSortPrepNand], key_nand);
        }
    }
}

// ================================================================ vertex stage of a standalone Renderer
// Renderer::render(encoder, view, indirect_args) on caller-owned buffers (src/renderer.rs:163-195): the splat records
// K1 normally produces are computed here for the `*count` Gaussians the indirect indices name, whatever produced them.
// A quad whose w <= 0 or whose (flat) depth lies outside [0,1] is clipped as a whole, as the fixed-function rasterizer
// clips it (render.wesl:123: clip_pos.zw = proj_pos.zw for all four vertices); x/y are not culled.
template <int SH, int COV>
__global__ void __launch_bounds__(256) vertex_kernel(const __grid_constant__ PreParams p, const uint32_t* __restrict__ indices,
                                                     const uint32_t* __restrict__ count) {
    constexpr int STRIDE = pod_stride(SH, COV);
    const Uniforms& u = p.u;
    const float sd_size = smul(u.std_dev, u.gsize);
    const uint32_t v = min(*count, p.n);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < v; i += gridDim.x * blockDim.x) {
        const uint32_t g = indices[i];
        if (g >= p.n) continue;
        const uint8_t* rec = p.gaussians + (size_t)g * STRIDE;
        CullOut co;
        cull_gaussian<SH, COV>(p, u, g, rec, sd_size, true, co);
        if (co.cw > 0.0f && co.nz >= 0.0f && co.nz <= 1.0f) {
            SplatGeom geo;
            splat_geometry(u, co, geo);
            emit_splat<SH, COV>(p, u, g, rec, co, geo);
        } else {
            *reinterpret_cast<uint2*>(&p.tboxes[g]) = make_uint2(1u | (1u << 16), 0u);  // no tiles
        }
    }
}

template <int SH, int COV>
cudaError_t launch_one(PreParams& p, int num_sms, cudaStream_t stream) {
    constexpr int STRIDE = pod_stride(SH, COV);
    constexpr int T = tile_records(STRIDE);
    constexpr int S = ring_stages(STRIDE);
    constexpr size_t smem = (size_t)S * T * STRIDE + 8 * (2 * S * (T / 32) + 16) + 4 * (S * (T / 32) + 8 + 8 * (T / 32) + 8);
    {   // per-device function attribute: set once per device, not once per process
        static bool configured[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            e = cudaFuncSetAttribute(preprocess_kernel<SH, COV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(preprocess_kernel<SH, COV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    p.num_tiles = (p.n + T - 1) / T;
    const int grid = (int)min((uint32_t)num_sms, p.num_tiles);
    if (p.strip_on && p.recs != nullptr)
        preprocess_kernel<SH, COV, true><<<grid, T + 64, smem, stream>>>(p);
    else
        preprocess_kernel<SH, COV, false><<<grid, T + 64, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int SH, int COV>
cudaError_t launch_vertex_one(const PreParams& p, const uint32_t* indices, const uint32_t* count, int num_sms, cudaStream_t stream) {
    vertex_kernel<SH, COV><<<num_sms * 8, 256, 0, stream>>>(p, indices, count);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_vertex_stage(int sh_fmt, int cov_fmt, PreParams& p, const uint32_t* indices, const uint32_t* count, int num_sms,
                                cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    p.selection = nullptr;  // the selection mask belongs to the Preprocessor
#define SB_CASE(SHV, COVV) \
    if (sh_fmt == SHV && cov_fmt == COVV) return launch_vertex_one<SHV, COVV>(p, indices, count, num_sms, stream);
    SB_CASE(SB_SH_SINGLE, SB_COV_SINGLE)
    SB_CASE(SB_SH_SINGLE, SB_COV_HALF)
    SB_CASE(SB_SH_SINGLE, SB_COV_ROT_SCALE)
    SB_CASE(SB_SH_HALF, SB_COV_SINGLE)
    SB_CASE(SB_SH_HALF, SB_COV_HALF)
    SB_CASE(SB_SH_HALF, SB_COV_ROT_SCALE)
    SB_CASE(SB_SH_NORM8, SB_COV_SINGLE)
    SB_CASE(SB_SH_NORM8, SB_COV_HALF)
    SB_CASE(SB_SH_NORM8, SB_COV_ROT_SCALE)
    SB_CASE(SB_SH_NONE, SB_COV_SINGLE)
    SB_CASE(SB_SH_NONE, SB_COV_HALF)
    SB_CASE(SB_SH_NONE, SB_COV_ROT_SCALE)
#undef SB_CASE
    return cudaErrorInvalidValue;
}

int preprocess_records_per_tile(int sh_fmt, int cov_fmt) { return tile_records(pod_stride(sh_fmt, cov_fmt)); }

size_t preprocess_scratch_prefix_bytes() { return (sort_prep_bytes() + 255) & ~(size_t)255; }

size_t preprocess_scratch_bytes(uint32_t n, int sh_fmt, int cov_fmt) {
    const int t = preprocess_records_per_tile(sh_fmt, cov_fmt);
    const size_t tiles = (n + t - 1) / t;
    return 16 + tiles * sizeof(unsigned long long);
}

cudaError_t launch_preprocess(int sh_fmt, int cov_fmt, PreParams& p, void* scratch, size_t scratch_bytes, int num_sms,
                              cudaStream_t stream) {
    if (p.n == 0) {  // nothing to do: args = {6,0,0,0} / {0,1,1}
        const SbDrawIndirectArgs d = {6, 0, 0, 0};
        const SbDispatchIndirectArgs s = {0, 1, 1};
        cudaError_t e = cudaMemcpyAsync(p.draw_args, &d, sizeof d, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(p.sort_args, &s, sizeof s, cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        return cudaMemsetAsync(p.visible_count, 0, 4, stream);
    }
    // one memset zeroes K1's ticket + look-back words AND, in front of them, the depth sort's prep words (p.sort_prep, when the
    // caller placed them there): the sort then needs no clearing launch of its own
    cudaError_t e = cudaMemsetAsync(scratch, 0, scratch_bytes, stream);
    if (e != cudaSuccess) return e;
    uint8_t* k1 = reinterpret_cast<uint8_t*>(scratch) + (p.sort_prep ? preprocess_scratch_prefix_bytes() : 0);
    p.tile_counter = reinterpret_cast<uint32_t*>(k1);
    p.tile_status = reinterpret_cast<unsigned long long*>(k1 + 16);
#define SB_CASE(SHV, COVV) \
    if (sh_fmt == SHV && cov_fmt == COVV) return launch_one<SHV, COVV>(p, num_sms, stream);
    SB_CASE(SB_SH_SINGLE, SB_COV_SINGLE)
    SB_CASE(SB_SH_SINGLE, SB_COV_HALF)
    SB_CASE(SB_SH_SINGLE, SB_COV_ROT_SCALE)
    SB_CASE(SB_SH_HALF, SB_COV_SINGLE)
    SB_CASE(SB_SH_HALF, SB_COV_HALF)
    SB_CASE(SB_SH_HALF, SB_COV_ROT_SCALE)
    SB_CASE(SB_SH_NORM8, SB_COV_SINGLE)
    SB_CASE(SB_SH_NORM8, SB_COV_HALF)
    SB_CASE(SB_SH_NORM8, SB_COV_ROT_SCALE)
    SB_CASE(SB_SH_NONE, SB_COV_SINGLE)
    SB_CASE(SB_SH_NONE, SB_COV_HALF)
    SB_CASE(SB_SH_NONE, SB_COV_ROT_SCALE)
#undef SB_CASE
    return cudaErrorInvalidValue;
}

}  // namespace sb
