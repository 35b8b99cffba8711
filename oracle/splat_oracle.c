/*
 * splat_oracle.c — CPU ORACLE (test infrastructure; see splat_oracle.h for the rules).
 *
 * Compile: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp -fPIC -shared
 * Every function cites the reference lines it restates.  Arithmetic contract: each f32
 * operation is rounded individually in the order written (no contraction), except the
 * explicit fmaf() calls in the fragment stage.
 */
#include "splat_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ small helpers */

static float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f; memcpy(&f, &bits, 4); return f;
}

static uint16_t f32_to_f16(float f) { /* round to nearest even */
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (absx > 0x7f800000u ? 0x200u : 0));
    if (absx >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u); /* overflow -> inf */
    if (absx < 0x33000001u) return (uint16_t)sign;              /* underflow -> 0 */
    int32_t e = (int32_t)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7fffffu) | 0x800000u;
    uint32_t shift, half;
    uint32_t out;
    if (e < -14) { /* subnormal result */
        shift = (uint32_t)(13 + (-14 - e));
        out = m >> shift;
        half = 1u << (shift - 1);
        uint32_t rem = m & ((1u << shift) - 1);
        if (rem > half || (rem == half && (out & 1u))) out++;
        return (uint16_t)(sign | out);
    }
    out = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ffu);
    uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (out & 1u))) out++;
    return (uint16_t)(sign | out);
}

/* next f16 toward -inf (dir<0) or +inf (dir>0) */
static uint16_t f16_step(uint16_t h, int dir) {
    int neg = (h & 0x8000u) != 0;
    if ((h & 0x7fffu) == 0) return dir > 0 ? 0x0001u : 0x8001u;
    if ((dir > 0) != neg) return (uint16_t)(h + 1);
    return (uint16_t)(h - 1);
}

typedef struct { float c[3][3]; } Mat3; /* c[col][row], column-major like WGSL/glam */
typedef struct { float c[4][4]; } Mat4;

/* glam Mat3::from_quat (glam 0.30.9); quaternion xyzw */
static Mat3 mat3_from_quat(const float q[4]) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, xy = x * y2, xz = x * z2;
    float yy = y * y2, yz = y * z2, zz = z * z2;
    float wx = w * x2, wy = w * y2, wz = w * z2;
    Mat3 m;
    m.c[0][0] = 1.0f - (yy + zz); m.c[0][1] = xy + wz;          m.c[0][2] = xz - wy;
    m.c[1][0] = xy - wz;          m.c[1][1] = 1.0f - (xx + zz); m.c[1][2] = yz + wx;
    m.c[2][0] = xz + wy;          m.c[2][1] = yz - wx;          m.c[2][2] = 1.0f - (xx + yy);
    return m;
}

/* mat4 * mat4, column j = ((A.c0*b0 + A.c1*b1) + A.c2*b2) + A.c3*b3 */
static Mat4 mat4_mul(const Mat4* a, const Mat4* b) {
    Mat4 r;
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++)
            r.c[j][i] = ((a->c[0][i] * b->c[j][0] + a->c[1][i] * b->c[j][1]) + a->c[2][i] * b->c[j][2]) +
                        a->c[3][i] * b->c[j][3];
    return r;
}

/* ------------------------------------------------------------------ layouts */

static uint32_t sh_bytes(int sh_fmt) {
    switch (sh_fmt) {
        case SO_SH_SINGLE: return 180; /* 45 x f32 */
        case SO_SH_HALF: return 92;    /* 46 x f16 (last = padding) */
        case SO_SH_NORM8: return 52;   /* f16 min, f16 max, 45 x unorm8, 3 pad */
        default: return 0;
    }
}
static uint32_t cov_bytes(int cov_fmt) {
    switch (cov_fmt) {
        case SO_COV_SINGLE: return 24;    /* 6 x f32 */
        case SO_COV_HALF: return 12;      /* 6 x f16 */
        default: return 28;               /* quat xyzw + scale xyz, f32 */
    }
}

uint32_t so_pod_stride(int sh_fmt, int cov_fmt) {
    uint32_t s = 16 + sh_bytes(sh_fmt) + cov_bytes(cov_fmt);
    return (s + 15u) & ~15u; /* WGSL struct alignment 16 (vec3<f32> member) */
}

/* cov3d = (R S)(R S)^T, upper triangle [xx, xy, xz, yy, yz, zz] */
static void cov3d_from_rot_scale(const float rot[4], const float scale[3], float out[6]) {
    Mat3 r = mat3_from_quat(rot);
    float m[3][3]; /* m[row][col] = R[col][row] * s[col] */
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) m[row][col] = r.c[col][row] * scale[col];
    int k = 0;
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++)
            out[k++] = (m[i][0] * m[j][0] + m[i][1] * m[j][1]) + m[i][2] * m[j][2];
}

void so_pack_gaussians(const SoGaussian* src, uint32_t n, int sh_fmt, int cov_fmt, void* out) {
    uint32_t stride = so_pod_stride(sh_fmt, cov_fmt);
    uint8_t* base = (uint8_t*)out;
    memset(base, 0, (size_t)stride * n);
    for (uint32_t i = 0; i < n; i++) {
        const SoGaussian* g = &src[i];
        uint8_t* p = base + (size_t)stride * i;
        memcpy(p, g->pos, 12);
        memcpy(p + 12, g->color, 4);
        uint8_t* q = p + 16;
        if (sh_fmt == SO_SH_SINGLE) {
            memcpy(q, g->sh, 180);
        } else if (sh_fmt == SO_SH_HALF) {
            for (int k = 0; k < 45; k++) { uint16_t h = f32_to_f16(g->sh[k]); memcpy(q + 2 * k, &h, 2); }
        } else if (sh_fmt == SO_SH_NORM8) {
            float mn = g->sh[0], mx = g->sh[0];
            for (int k = 1; k < 45; k++) { if (g->sh[k] < mn) mn = g->sh[k]; if (g->sh[k] > mx) mx = g->sh[k]; }
            uint16_t hmn = f32_to_f16(mn), hmx = f32_to_f16(mx);
            /* make the decoded range enclose the data */
            if (f16_to_f32(hmn) > mn) hmn = f16_step(hmn, -1);
            if (f16_to_f32(hmx) < mx) hmx = f16_step(hmx, +1);
            float fmn = f16_to_f32(hmn), fmx = f16_to_f32(hmx);
            memcpy(q, &hmn, 2); memcpy(q + 2, &hmx, 2);
            float range = fmx - fmn;
            for (int k = 0; k < 45; k++) {
                float t = range > 0.0f ? (g->sh[k] - fmn) / range : 0.0f;
                float v = rintf(t * 255.0f);
                if (v < 0.0f) v = 0.0f;
                if (v > 255.0f) v = 255.0f;
                q[4 + k] = (uint8_t)v;
            }
        }
        uint8_t* c = q + sh_bytes(sh_fmt);
        if (cov_fmt == SO_COV_ROT_SCALE) {
            memcpy(c, g->rot, 16);
            memcpy(c + 16, g->scale, 12);
        } else {
            float cov[6];
            cov3d_from_rot_scale(g->rot, g->scale, cov);
            if (cov_fmt == SO_COV_SINGLE) memcpy(c, cov, 24);
            else for (int k = 0; k < 6; k++) { uint16_t h = f32_to_f16(cov[k]); memcpy(c + 2 * k, &h, 2); }
        }
    }
}

/* gaussian_unpack_* (core WESL, external) */
static void unpack_cov3d(const uint8_t* pod, int sh_fmt, int cov_fmt, float cov[6]) {
    const uint8_t* c = pod + 16 + sh_bytes(sh_fmt);
    if (cov_fmt == SO_COV_SINGLE) {
        memcpy(cov, c, 24);
    } else if (cov_fmt == SO_COV_HALF) {
        for (int k = 0; k < 6; k++) { uint16_t h; memcpy(&h, c + 2 * k, 2); cov[k] = f16_to_f32(h); }
    } else {
        float rot[4], scale[3];
        memcpy(rot, c, 16); memcpy(scale, c + 16, 12);
        cov3d_from_rot_scale(rot, scale, cov);
    }
}
static void unpack_sh(const uint8_t* pod, int sh_fmt, float sh[45]) {
    const uint8_t* q = pod + 16;
    if (sh_fmt == SO_SH_SINGLE) {
        memcpy(sh, q, 180);
    } else if (sh_fmt == SO_SH_HALF) {
        for (int k = 0; k < 45; k++) { uint16_t h; memcpy(&h, q + 2 * k, 2); sh[k] = f16_to_f32(h); }
    } else if (sh_fmt == SO_SH_NORM8) {
        uint16_t hmn, hmx; memcpy(&hmn, q, 2); memcpy(&hmx, q + 2, 2);
        float mn = f16_to_f32(hmn), mx = f16_to_f32(hmx);
        for (int k = 0; k < 45; k++) {
            float t = (float)q[4 + k] / 255.0f;
            sh[k] = mn * (1.0f - t) + mx * t; /* WGSL mix */
        }
    } else {
        for (int k = 0; k < 45; k++) sh[k] = 0.0f;
    }
}

/* PLY -> Gaussian (core, recalled).  props: x y z nx ny nz f_dc[3]
 * f_rest[45] opacity scale[3] rot[4] (rot_0 = w). */
void so_gaussians_from_ply_props(const float* props, uint32_t n, SoGaussian* out) {
    const float SH_C0 = 0.2820948f;
    for (uint32_t i = 0; i < n; i++) {
        const float* p = props + (size_t)62 * i;
        SoGaussian* g = &out[i];
        g->pos[0] = p[0]; g->pos[1] = p[1]; g->pos[2] = p[2];
        for (int c = 0; c < 3; c++) {
            float v = (0.5f + SH_C0 * p[6 + c]) * 255.0f;
            if (!(v > 0.0f)) v = 0.0f;
            if (v > 255.0f) v = 255.0f;
            g->color[c] = (uint8_t)v; /* Rust `as u8`: saturating truncation */
        }
        float op = 1.0f / (1.0f + expf(-p[54]));
        float a = op * 255.0f;
        if (!(a > 0.0f)) a = 0.0f;
        if (a > 255.0f) a = 255.0f;
        g->color[3] = (uint8_t)a;
        for (int k = 0; k < 15; k++)
            for (int c = 0; c < 3; c++) g->sh[k * 3 + c] = p[9 + c * 15 + k];
        g->scale[0] = expf(p[55]); g->scale[1] = expf(p[56]); g->scale[2] = expf(p[57]);
        float qx = p[59], qy = p[60], qz = p[61], qw = p[58];
        float len = sqrtf(((qx * qx + qy * qy) + qz * qz) + qw * qw);
        float inv = 1.0f / len;
        g->rot[0] = qx * inv; g->rot[1] = qy * inv; g->rot[2] = qz * inv; g->rot[3] = qw * inv;
    }
}

/* ------------------------------------------------------------------ camera (src/camera.rs:71-93, glam) */

void so_camera_pod(const float pos[3], float yaw, float pitch, float z_near, float z_far,
                   float fov_y, uint32_t width, uint32_t height, SoCameraPod* out) {
    /* get_forward: src/camera.rs:71-77 */
    float f[3] = { cosf(pitch) * sinf(yaw), sinf(pitch), cosf(pitch) * cosf(yaw) };
    /* Mat4::look_to_rh(eye, dir, up=Y): f = normalize(dir); s = normalize(cross(f, up)); u = cross(s, f) */
    float fl = 1.0f / sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
    f[0] *= fl; f[1] *= fl; f[2] *= fl;
    const float up[3] = { 0.0f, 1.0f, 0.0f };
    float s[3] = { f[1] * up[2] - up[1] * f[2], f[2] * up[0] - up[2] * f[0], f[0] * up[1] - up[0] * f[1] };
    float sl = 1.0f / sqrtf((s[0] * s[0] + s[1] * s[1]) + s[2] * s[2]);
    s[0] *= sl; s[1] *= sl; s[2] *= sl;
    float u[3] = { s[1] * f[2] - f[1] * s[2], s[2] * f[0] - f[2] * s[0], s[0] * f[1] - f[0] * s[1] };
    float ds = (pos[0] * s[0] + pos[1] * s[1]) + pos[2] * s[2];
    float du = (pos[0] * u[0] + pos[1] * u[1]) + pos[2] * u[2];
    float df = (pos[0] * f[0] + pos[1] * f[1]) + pos[2] * f[2];
    float view[16] = { s[0], u[0], -f[0], 0.0f, s[1], u[1], -f[1], 0.0f,
                       s[2], u[2], -f[2], 0.0f, -ds, -du, df, 1.0f };
    memcpy(out->view, view, sizeof view);
    /* Mat4::perspective_rh (depth 0..1); aspect = w/h as in CameraPod::new (buffer/camera.rs:72-80) */
    float aspect = (float)width / (float)height;
    float sn = sinf(0.5f * fov_y), cs = cosf(0.5f * fov_y);
    float h = cs / sn;
    float w = h / aspect;
    float r = z_far / (z_near - z_far);
    float proj[16] = { w, 0, 0, 0, 0, h, 0, 0, 0, 0, r, -1.0f, 0, 0, r * z_near, 0 };
    memcpy(out->proj, proj, sizeof proj);
    out->size[0] = (float)width; out->size[1] = (float)height;
    out->pad[0] = out->pad[1] = 0;
}

void so_model_transform_pod(const float pos[3], const float rot[4], const float scale[3],
                            SoModelTransformPod* out) {
    memset(out, 0, sizeof *out);
    memcpy(out->pos, pos, 12); memcpy(out->rot, rot, 16); memcpy(out->scale, scale, 12);
}

void so_gaussian_transform_pod(float size, int display_mode, int sh_deg, int no_sh0,
                               float max_std_dev, SoGaussianTransformPod* out) {
    out->size = size;
    out->display_mode = (uint8_t)display_mode;
    out->sh_deg = (uint8_t)sh_deg;
    out->no_sh0 = (uint8_t)(no_sh0 != 0);
    float v = rintf(max_std_dev / 3.0f * 255.0f);
    if (v < 0.0f) v = 0.0f;
    if (v > 255.0f) v = 255.0f;
    out->max_std_dev = (uint8_t)v;
}

/* ------------------------------------------------------------------ per-frame uniforms */

typedef struct {
    Mat4 view, proj, pv, model, vm;
    Mat3 sr, inv_sr, w;
    float size[2], focal[2], cam_pos[3];
    float std_dev;      /* gaussian_transform_max_std_dev */
    float gsize;        /* gaussian_transform.size */
    int mode, sh_deg, no_sh0;
} Uniforms;

static Uniforms make_uniforms(const SoCameraPod* cam, const SoModelTransformPod* mt,
                              const SoGaussianTransformPod* gt) {
    Uniforms u;
    memcpy(u.view.c, cam->view, 64);
    memcpy(u.proj.c, cam->proj, 64);
    u.size[0] = cam->size[0]; u.size[1] = cam->size[1];
    /* model_transform_mat = T(pos) R(rot) S(scale); model_scale_rot_mat = R S;
       model_transform_inv_sr_mat = S^-1 R^T   (core WESL, recalled) */
    Mat3 r = mat3_from_quat(mt->rot);
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) u.sr.c[col][row] = r.c[col][row] * mt->scale[col];
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) u.inv_sr.c[col][row] = r.c[row][col] / mt->scale[row];
    memset(u.model.c, 0, 64);
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) u.model.c[col][row] = u.sr.c[col][row];
    u.model.c[3][0] = mt->pos[0]; u.model.c[3][1] = mt->pos[1]; u.model.c[3][2] = mt->pos[2];
    u.model.c[3][3] = 1.0f;
    u.pv = mat4_mul(&u.proj, &u.view);  /* camera.wesl:13-15: (proj * view) * world */
    u.vm = mat4_mul(&u.view, &u.model); /* utils.wesl:37: (view * model_mat) * pos */
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) u.w.c[col][row] = u.view.c[col][row];
    /* focal: utils.wesl:35 */
    u.focal[0] = u.proj.c[0][0] * u.size[0] * 0.5f;
    u.focal[1] = u.proj.c[1][1] * u.size[1] * 0.5f;
    /* world_camera_pos: render.wesl:59-63  = -(transpose(W) * view[3].xyz) */
    for (int i = 0; i < 3; i++) {
        float d = (u.w.c[i][0] * u.view.c[3][0] + u.w.c[i][1] * u.view.c[3][1]) + u.w.c[i][2] * u.view.c[3][2];
        u.cam_pos[i] = -d;
    }
    u.std_dev = (float)gt->max_std_dev / 255.0f * 3.0f;
    u.gsize = gt->size;
    u.mode = gt->display_mode; u.sh_deg = gt->sh_deg; u.no_sh0 = gt->no_sh0;
    return u;
}

/* utils.wesl:18-20 */
static int cull(float x, float y, float z) {
    return !((x >= -1.0f && y >= -1.0f && z >= 0.0f) && (x <= 1.0f && y <= 1.0f && z <= 1.0f));
}

/* utils.wesl:25-48.  Zero terms of J are skipped (adds of exact zeros). */
static void cov2d(const Uniforms* u, const float pos[3], const float cov3d[6], float out[3]) {
    const Mat4* vm = &u->vm;
    float t[3];
    for (int i = 0; i < 3; i++)
        t[i] = ((vm->c[0][i] * pos[0] + vm->c[1][i] * pos[1]) + vm->c[2][i] * pos[2]) + vm->c[3][i];
    float j00 = u->focal[0] / t[2];
    float j02 = -(u->focal[0] * t[0]) / (t[2] * t[2]);
    float j11 = u->focal[1] / t[2];
    float j12 = -(u->focal[1] * t[1]) / (t[2] * t[2]);
    /* JW = J * W  (rows 0,1) */
    float jw[2][3];
    for (int c = 0; c < 3; c++) {
        jw[0][c] = j00 * u->w.c[c][0] + j02 * u->w.c[c][2];
        jw[1][c] = j11 * u->w.c[c][1] + j12 * u->w.c[c][2];
    }
    /* T = JW * SR */
    float tm[2][3];
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++)
            tm[r][c] = (jw[r][0] * u->sr.c[c][0] + jw[r][1] * u->sr.c[c][1]) + jw[r][2] * u->sr.c[c][2];
    float v[3][3] = { { cov3d[0], cov3d[1], cov3d[2] }, { cov3d[1], cov3d[3], cov3d[4] }, { cov3d[2], cov3d[4], cov3d[5] } };
    /* TV = T * Vrk ; cov = TV * T^T */
    float tv[2][3];
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++)
            tv[r][c] = (tm[r][0] * v[0][c] + tm[r][1] * v[1][c]) + tm[r][2] * v[2][c];
    out[0] = (tv[0][0] * tm[0][0] + tv[0][1] * tm[0][1]) + tv[0][2] * tm[0][2]; /* cov2d[0][0] */
    out[1] = (tv[1][0] * tm[0][0] + tv[1][1] * tm[0][1]) + tv[1][2] * tm[0][2]; /* cov2d[0][1] = col 0,row 1 */
    out[2] = (tv[1][0] * tm[1][0] + tv[1][1] * tm[1][1]) + tv[1][2] * tm[1][2]; /* cov2d[1][1] */
}

/* utils.wesl:54-79 */
static void cov2d_axes(const Uniforms* u, const float pos[3], const float cov3d[6], float std_dev, float axes[4]) {
    float cv[3];
    cov2d(u, pos, cov3d, cv);
    float mid = 0.5f * (cv[0] + cv[2]);
    float hx = 0.5f * (cv[0] - cv[2]);
    float radius = sqrtf(hx * hx + cv[1] * cv[1]);
    float major_lambda = mid + radius;
    float minor_lambda = mid - radius;
    if (minor_lambda < 0.0f) { axes[0] = axes[1] = axes[2] = axes[3] = 0.0f; return; }
    float dx = cv[1], dy = major_lambda - cv[0];
    float ddx, ddy;
    if (dx == 0.0f && dy == 0.0f) { ddx = 0.0f; ddy = 1.0f; }
    else { float l = sqrtf(dx * dx + dy * dy); ddx = dx / l; ddy = dy / l; }
    float major_len = fminf(std_dev * sqrtf(major_lambda), 1024.0f);
    float minor_len = fminf(std_dev * sqrtf(minor_lambda), 1024.0f);
    axes[0] = major_len * ddx; axes[1] = major_len * ddy;
    axes[2] = minor_len * ddy; axes[3] = minor_len * -ddx;
}

/* ------------------------------------------------------------------ preprocess (preprocess.wesl:52-126) */

static void project_centre(const Uniforms* u, const float pos[3], float world[3], float clip[4]) {
    const Mat4* m = &u->model;
    for (int i = 0; i < 3; i++)
        world[i] = ((m->c[0][i] * pos[0] + m->c[1][i] * pos[1]) + m->c[2][i] * pos[2]) + m->c[3][i];
    const Mat4* pv = &u->pv;
    for (int i = 0; i < 4; i++)
        clip[i] = ((pv->c[0][i] * world[0] + pv->c[1][i] * world[1]) + pv->c[2][i] * world[2]) + pv->c[3][i];
}

/* returns 1 when visible; ndc_z receives ndc.z */
static int preprocess_one(const Uniforms* u, const SoModel* model, uint32_t index, float* ndc_z) {
    if (model->selection) { /* preprocess.wesl:68-78 */
        int bit = (model->selection[index / 32u] >> (index % 32u)) & 1u;
        int inverted = model->invert_selection != 0u;
        if (inverted == bit) return 0;
    }
    uint32_t stride = so_pod_stride(model->sh_fmt, model->cov_fmt);
    const uint8_t* pod = (const uint8_t*)model->pods + (size_t)stride * index;
    float pos[3]; memcpy(pos, pod, 12);
    float world[3], clip[4];
    project_centre(u, pos, world, clip);
    float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
    if (cull(nx, ny, nz)) { /* preprocess.wesl:87-99 */
        float cov3d[6], axes[4];
        unpack_cov3d(pod, model->sh_fmt, model->cov_fmt, cov3d);
        cov2d_axes(u, pos, cov3d, u->std_dev * u->gsize, axes);
        float mx = axes[0] * u->std_dev / u->size[0];
        float my = axes[1] * u->std_dev / u->size[1];
        float ndc_major_len = sqrtf(mx * mx + my * my);
        float l = sqrtf(nx * nx + ny * ny);
        float dirx = -nx / l, diry = -ny / l;
        float m = fminf(ndc_major_len, l);
        float bx = nx + m * dirx, by = ny + m * diry;
        if (cull(bx, by, nz)) return 0;
    }
    *ndc_z = nz;
    return 1;
}

uint32_t so_preprocess(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
                       uint32_t* indices, float* keys, uint32_t* visible_mask,
                       uint32_t draw_args[4], uint32_t sort_args[3]) {
    Uniforms u = make_uniforms(cam, &model->model_transform, gt);
    uint32_t n = model->n;
    uint32_t words = (n + 31u) / 32u;
    uint32_t* mask = visible_mask ? visible_mask : (uint32_t*)calloc(words ? words : 1, 4);
    float* zs = (float*)malloc((size_t)(n ? n : 1) * 4);
    memset(mask, 0, (size_t)words * 4);
    #pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < (int64_t)words; w++) {
        uint32_t bits = 0;
        for (uint32_t b = 0; b < 32; b++) {
            uint32_t i = (uint32_t)w * 32u + b;
            if (i >= n) break;
            float z;
            if (preprocess_one(&u, model, i, &z)) { bits |= 1u << b; zs[i] = z; }
        }
        mask[w] = bits;
    }
    uint32_t v = 0;
    for (uint32_t i = 0; i < n; i++) {
        if ((mask[i / 32u] >> (i % 32u)) & 1u) {
            indices[v] = i;
            keys[v] = 1.0f - zs[i]; /* preprocess.wesl:105 */
            v++;
        }
    }
    /* post: preprocess.wesl:108-126 */
    uint32_t blocks = (v + 3839u) / 3840u;
    uint32_t cap = ((n + 3839u) / 3840u) * 3840u;
    uint32_t padded = blocks * 3840u < cap ? blocks * 3840u : cap;
    for (uint32_t i = v; i < padded; i++) keys[i] = 2.0f;
    if (draw_args) { draw_args[0] = 6; draw_args[1] = v; draw_args[2] = 0; draw_args[3] = 0; }
    if (sort_args) { sort_args[0] = blocks; sort_args[1] = 1; sort_args[2] = 1; }
    free(zs);
    if (!visible_mask) free(mask);
    return v;
}

/* ------------------------------------------------------------------ sort (radix_sort.wgsl, semantics) */

void so_radix_sort(uint32_t* keys, uint32_t* payload, uint32_t count) {
    if (count == 0) return;
    uint32_t* k2 = (uint32_t*)malloc((size_t)count * 4);
    uint32_t* p2 = (uint32_t*)malloc((size_t)count * 4);
    uint32_t *ka = keys, *pa = payload, *kb = k2, *pb = p2;
    for (int pass = 0; pass < 4; pass++) { /* 8-bit digits, LSB first, stable */
        uint32_t hist[256]; memset(hist, 0, sizeof hist);
        int shift = pass * 8;
        for (uint32_t i = 0; i < count; i++) hist[(ka[i] >> shift) & 255u]++;
        uint32_t sum = 0;
        for (int d = 0; d < 256; d++) { uint32_t c = hist[d]; hist[d] = sum; sum += c; }
        for (uint32_t i = 0; i < count; i++) {
            uint32_t d = (ka[i] >> shift) & 255u;
            uint32_t o = hist[d]++;
            kb[o] = ka[i]; pb[o] = pa[i];
        }
        uint32_t* t = ka; ka = kb; kb = t;
        t = pa; pa = pb; pb = t;
    }
    free(k2); free(p2); /* 4 passes: result is back in keys/payload */
}

/* ------------------------------------------------------------------ vertex stage (render.wesl:58-130) */

/* utils.wesl:82-135.  Contract: the polynomial is evaluated as one fma chain per degree (WGSL
 * allows contraction); the basis factors are plain individually rounded products. */
static void view_color(const Uniforms* u, const uint8_t* pod, int sh_fmt, const float dir[3], float rgba[4]) {
    const float sh_c1 = 0.4886025f;
    const float sh_c2[5] = { 1.0925484f, -1.0925484f, 0.3153916f, -1.0925484f, 0.5462742f };
    const float sh_c3[7] = { -0.5900436f, 2.8906114f, -0.4570458f, 0.3731763f, -0.4570458f, 1.4453057f, -0.5900436f };
    float x = dir[0], y = dir[1], z = dir[2];
    float col[4];
    for (int c = 0; c < 4; c++) col[c] = (float)pod[12 + c] / 255.0f; /* unpack4x8unorm */
    float sh[45];
    unpack_sh(pod, sh_fmt, sh);
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    float b2[5] = { sh_c2[0] * xy, sh_c2[1] * yz, sh_c2[2] * ((2.0f * zz - xx) - yy), sh_c2[3] * xz, sh_c2[4] * (xx - yy) };
    float b3[7] = { sh_c3[0] * y * (3.0f * xx - yy), sh_c3[1] * xy * z, sh_c3[2] * y * ((4.0f * zz - xx) - yy),
                    sh_c3[3] * z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy), sh_c3[4] * x * ((4.0f * zz - xx) - yy),
                    sh_c3[5] * z * (xx - yy), sh_c3[6] * x * (xx - 3.0f * yy) };
    for (int c = 0; c < 3; c++) {
        float result = u->no_sh0 ? 0.5f : col[c];
        #define SH(i) sh[(i) * 3 + c]
        if (u->sh_deg >= 1 && sh_fmt != SO_SH_NONE) {
            float t = fmaf(SH(1), z, -(SH(0) * y));          /* -sh0*y + sh1*z */
            t = fmaf(-SH(2), x, t);                          /* ... - sh2*x */
            result = fmaf(sh_c1, t, result);
            if (u->sh_deg >= 2) {
                float acc = b2[0] * SH(3);
                for (int k = 1; k < 5; k++) acc = fmaf(b2[k], SH(3 + k), acc);
                result = result + acc;
                if (u->sh_deg >= 3) {
                    acc = b3[0] * SH(8);
                    for (int k = 1; k < 7; k++) acc = fmaf(b3[k], SH(8 + k), acc);
                    result = result + acc;
                }
            }
        }
        #undef SH
        rgba[c] = fmaxf(result, 0.0f);
    }
    rgba[3] = col[3];
}

static void project_one(const Uniforms* u, const SoModel* model, uint32_t index, SoSplat* s) {
    uint32_t stride = so_pod_stride(model->sh_fmt, model->cov_fmt);
    const uint8_t* pod = (const uint8_t*)model->pods + (size_t)stride * index;
    float pos[3]; memcpy(pos, pod, 12);
    float world[3], clip[4];
    project_centre(u, pos, world, clip);
    float nx = clip[0] / clip[3], ny = clip[1] / clip[3];
    memset(s, 0, sizeof *s);
    s->z = clip[2] / clip[3];
    s->cx = (nx + 1.0f) * 0.5f * u->size[0];
    s->cy = (1.0f - ny) * 0.5f * u->size[1];
    /* color(): render.wesl:58-73 */
    float vd[3] = { u->cam_pos[0] - world[0], u->cam_pos[1] - world[1], u->cam_pos[2] - world[2] };
    float md[3];
    for (int i = 0; i < 3; i++)
        md[i] = (u->inv_sr.c[0][i] * vd[0] + u->inv_sr.c[1][i] * vd[1]) + u->inv_sr.c[2][i] * vd[2];
    /* -normalize(v): contract = v * (1 / length(v)) */
    float inv_ml = 1.0f / sqrtf((md[0] * md[0] + md[1] * md[1]) + md[2] * md[2]);
    float dir[3] = { -(md[0] * inv_ml), -(md[1] * inv_ml), -(md[2] * inv_ml) };
    float rgba[4];
    view_color(u, pod, model->sh_fmt, dir, rgba);
    s->r = rgba[0]; s->g = rgba[1]; s->b = rgba[2]; s->a = rgba[3];
    if (u->mode == SO_MODE_POINT) { /* render.wesl:92-104 */
        const Mat4* vm = &u->vm;
        float vp[3];
        for (int i = 0; i < 3; i++)
            vp[i] = ((vm->c[0][i] * pos[0] + vm->c[1][i] * pos[1]) + vm->c[2][i] * pos[2]) + vm->c[3][i];
        float len = sqrtf((vp[0] * vp[0] + vp[1] * vp[1]) + vp[2] * vp[2]);
        float half = 0.01f * u->gsize * 0.5f * u->size[1] / len; /* pixels, both axes */
        float inv = 1.0f / half;
        s->ax = inv; s->ay = 0.0f; s->bx = 0.0f; s->by = inv;
        s->ext_x = half; s->ext_y = half;
        s->valid = (half > 0.0f) && isfinite(inv) && isfinite(s->cx) && isfinite(s->cy);
        return;
    }
    float cov3d[6], axes[4];
    unpack_cov3d(pod, model->sh_fmt, model->cov_fmt, cov3d);
    cov2d_axes(u, pos, cov3d, u->std_dev * u->gsize, axes);
    float mm = axes[0] * axes[0] + axes[1] * axes[1];
    float nn = axes[2] * axes[2] + axes[3] * axes[3];
    /* pixel delta d' = (dx, -dy) = (q.x*major + q.y*minor)/2  =>  q.x = 2 d'.major/|major|^2 */
    s->ax = 2.0f * axes[0] / mm; s->ay = -(2.0f * axes[1] / mm);
    s->bx = 2.0f * axes[2] / nn; s->by = -(2.0f * axes[3] / nn);
    float hs = 0.5f * u->std_dev;
    s->ext_x = hs * sqrtf(axes[0] * axes[0] + axes[2] * axes[2]);
    s->ext_y = hs * sqrtf(axes[1] * axes[1] + axes[3] * axes[3]);
    s->valid = isfinite(s->ax) && isfinite(s->ay) && isfinite(s->bx) && isfinite(s->by) &&
               isfinite(s->cx) && isfinite(s->cy) && isfinite(s->ext_x) && isfinite(s->ext_y);
}

void so_project(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
                const uint32_t* indices, uint32_t count, SoSplat* out) {
    Uniforms u = make_uniforms(cam, &model->model_transform, gt);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)count; i++) project_one(&u, model, indices[i], &out[i]);
}

/* ------------------------------------------------------------------ fragment stage + blending */

/* exp(-x), x >= 0: y = x*(-log2 e); n = rint(y); f = y-n; 2^f by degree-6 Horner in fma;
 * scaled by 2^n through the exponent field.  Every step is one exactly-rounded IEEE op. */
float so_exp_neg_poly(float x) {
    float y = x * -1.44269504f;
    float n = rintf(y);
    float f = y - n;
    float p = 1.54035304e-4f;
    p = fmaf(p, f, 1.33335581e-3f);
    p = fmaf(p, f, 9.61812911e-3f);
    p = fmaf(p, f, 5.55041087e-2f);
    p = fmaf(p, f, 2.40226507e-1f);
    p = fmaf(p, f, 6.93147181e-1f);
    p = fmaf(p, f, 1.0f);
    if (n < -125.0f) return 0.0f;
    uint32_t bits; memcpy(&bits, &p, 4);
    bits += (uint32_t)((int32_t)n << 23);
    memcpy(&p, &bits, 4);
    return p;
}

/* `*Srgb` targets (the reference accepts any TextureFormat: src/renderer.rs:120-155; its doctests use Rgba8UnormSrgb,
 * src/selection/mod.rs:46): the attachment stores sRGB codes; ALPHA_BLENDING decodes the destination to linear, blends, and
 * encodes the result back (Vulkan 1.3 §29.1 / §16 "sRGB"), once per blend.  Tables: oracle/srgb_tables.h (generated). */
#include "srgb_tables.h"
static float srgb_decode(int code) { float f; memcpy(&f, &SO_SRGB_DECODE_BITS[code], 4); return f; }
static int srgb_encode(float x) { /* number of thresholds 1..255 that are <= x */
    int lo = 0, hi = 255;
    while (lo < hi) {
        int mid = (lo + hi + 1) / 2;
        float t; memcpy(&t, &SO_SRGB_THRESHOLD_BITS[mid], 4);
        if (x >= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}
static int fmt_is_srgb(int f) { return f == SO_TARGET_RGBA8_SRGB || f == SO_TARGET_BGRA8_SRGB; }
static int fmt_is_unorm(int f) { return f == SO_TARGET_RGBA8 || f == SO_TARGET_BGRA8 || fmt_is_srgb(f); } /* 8-bit codes in memory */
static int fmt_is_bgra(int f) { return f == SO_TARGET_BGRA8 || f == SO_TARGET_BGRA8_SRGB; }

static int depth_passes(int compare, float z, float d) {
    switch (compare) {
        case 1: return 0;
        case 2: return z < d;
        case 3: return z == d;
        case 4: return z <= d;
        case 5: return z > d;
        case 6: return z != d;
        case 7: return z >= d;
        default: return 1;
    }
}

static void raster_band(const SoSplat* sp, uint32_t count, const Uniforms* u, int strict_exp,
                        int is_unorm, int is_f16, uint32_t width, uint32_t y_lo, uint32_t y_hi,
                        uint32_t row0, float* acc, uint64_t* bbox_px, uint64_t* alive_px,
                        float* depth, int depth_compare, int depth_write) {
    const float sd = u->std_dev;
    const float sd2 = sd * sd;
    const float ol = (sd - 0.1f) * (sd - 0.1f);
    uint64_t nb = 0, na = 0;
    for (uint32_t k = 0; k < count; k++) {
        const SoSplat* s = &sp[k];
        if (!s->valid) continue;
        float fx0 = floorf(s->cx - s->ext_x - 1.0f), fx1 = ceilf(s->cx + s->ext_x + 1.0f);
        float fy0 = floorf(s->cy - s->ext_y - 1.0f), fy1 = ceilf(s->cy + s->ext_y + 1.0f);
        if (fx1 < 0.0f || fy1 < (float)y_lo || fx0 >= (float)width || fy0 >= (float)y_hi) continue;
        uint32_t x0 = fx0 < 0.0f ? 0u : (uint32_t)fx0;
        uint32_t x1 = fx1 >= (float)width ? width - 1 : (uint32_t)fx1;
        uint32_t y0 = fy0 < (float)y_lo ? y_lo : (uint32_t)fy0;
        uint32_t y1 = fy1 >= (float)y_hi ? y_hi - 1 : (uint32_t)fy1;
        /* Fixed-point attachments clamp the SOURCE colour to [0,1] before the blend equation
           (Vulkan 1.3 spec 29.1 "Blending": "If the color attachment is fixed-point, the components of
           the source and destination values and blend factors are each clamped to [0,1]"); SH colours
           can exceed 1 (utils.wesl:82-135 only clamps below).  With c, d <= 255 and alpha in [0,1] the
           post-blend clamp below can never bind after rounding, it is kept as a statement of the spec. */
        float c255[3] = { fminf(s->r * 255.0f, 255.0f), fminf(s->g * 255.0f, 255.0f), fminf(s->b * 255.0f, 255.0f) };
        float cf[3] = { s->r, s->g, s->b };
        for (uint32_t py = y0; py <= y1; py++) {
            float dy = ((float)py + 0.5f) - s->cy;
            float* row = acc + ((size_t)(py - row0) * width) * 4;
            for (uint32_t px = x0; px <= x1; px++) {
                float dx = ((float)px + 0.5f) - s->cx;
                float qx = fmaf(dx, s->ax, dy * s->ay);   /* fma */
                float qy = fmaf(dx, s->bx, dy * s->by);   /* fma */
                nb++;
                float alpha;
                if (u->mode == SO_MODE_POINT) { /* render.wesl:164-166 */
                    if (!(fabsf(qx) <= 1.0f && fabsf(qy) <= 1.0f)) continue;
                    alpha = 1.0f;
                } else {
                    float r2 = fmaf(qx, qx, qy * qy);     /* fma; dot(quad_offset, quad_offset) */
                    if (!(r2 <= sd2)) continue;           /* discard: render.wesl:145,155 */
                    if (u->mode == SO_MODE_SPLAT) {       /* render.wesl:149 */
                        float e = strict_exp ? so_exp_neg_poly(r2) : expf(-r2);
                        alpha = s->a * e;
                    } else {                              /* render.wesl:159-160 */
                        float outline = r2 > ol ? 1.0f : 0.0f;
                        alpha = s->a + (1.0f - s->a) * outline;
                    }
                }
                na++;
                if (depth) { /* depth test after the shader's discard: src/renderer.rs:304 */
                    float* dz = depth + (size_t)(py - row0) * width + px;
                    if (!depth_passes(depth_compare, s->z, *dz)) continue;
                    if (depth_write) *dz = s->z;
                }
                float om = 1.0f - alpha;
                float* d = row + (size_t)px * 4;
                if (is_unorm == 2) {
                    /* sRGB attachment: state = the stored code; decode, blend in linear with the source clamped to [0,1], encode */
                    for (int c = 0; c < 3; c++)
                        d[c] = (float)srgb_encode(fmaf(srgb_decode((int)d[c]), om, fminf(cf[c], 1.0f) * alpha));
                } else if (is_unorm) {
                    /* ALPHA_BLENDING into unorm8 (renderer.rs:296-300): the target is
                       re-quantised after every blend; state kept in 0..255 units. */
                    for (int c = 0; c < 3; c++)
                        d[c] = rintf(fminf(fmaf(d[c], om, c255[c] * alpha), 255.0f));
                } else {
                    for (int c = 0; c < 3; c++) d[c] = fmaf(d[c], om, cf[c] * alpha);
                    d[3] = fmaf(d[3], om, alpha);
                    if (is_f16) for (int c = 0; c < 4; c++) d[c] = f16_to_f32(f32_to_f16(d[c]));
                }
            }
        }
    }
    *bbox_px += nb; *alive_px += na;
}

void so_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int so_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void so_render(const SoModel* models, uint32_t n_models, const SoCameraPod* cam,
               const SoGaussianTransformPod* gt, int target_format, int strict_exp,
               uint32_t row0, uint32_t rows, void* target, SoStats* stats, int n_threads) {
    uint32_t width = (uint32_t)cam->size[0], height = (uint32_t)cam->size[1];
    if (row0 > height) row0 = height;
    if (rows > height - row0) rows = height - row0;
    int is_unorm = fmt_is_unorm(target_format) ? (fmt_is_srgb(target_format) ? 2 : 1) : 0;
    int is_f16 = target_format == SO_TARGET_RGBA16F;
    if (n_threads <= 0) n_threads = so_max_threads();
    /* clear to BLACK = (0,0,0,1): renderer.rs:171-177 */
    float* acc = (float*)malloc((size_t)(rows ? rows : 1) * width * 4 * sizeof(float));
    for (size_t i = 0; i < (size_t)rows * width; i++) {
        acc[i * 4 + 0] = 0.0f; acc[i * 4 + 1] = 0.0f; acc[i * 4 + 2] = 0.0f; acc[i * 4 + 3] = 1.0f;
    }
    SoStats st = { 0, 0, 0 };
    for (uint32_t m = 0; m < n_models; m++) { /* multi_model.rs:491-527: model k+1 over model k */
        const SoModel* model = &models[m];
        uint32_t n = model->n;
        uint32_t cap = ((n + 3839u) / 3840u) * 3840u;
        uint32_t* idx = (uint32_t*)malloc((size_t)(cap ? cap : 1) * 4);
        float* keys = (float*)malloc((size_t)(cap ? cap : 1) * 4);
        uint32_t v = so_preprocess(model, cam, gt, idx, keys, NULL, NULL, NULL);
        so_radix_sort((uint32_t*)keys, idx, v); /* pads (2.0) stay behind the V valid keys */
        SoSplat* sp = (SoSplat*)malloc((size_t)(v ? v : 1) * sizeof(SoSplat));
        so_project(model, cam, gt, idx, v, sp);
        Uniforms u = make_uniforms(cam, &model->model_transform, gt);
        st.visible += v;
        uint64_t nb = 0, na = 0;
        #pragma omp parallel for schedule(static) num_threads(n_threads) reduction(+ : nb, na)
        for (int t = 0; t < n_threads; t++) {
            uint32_t y_lo = row0 + (uint32_t)((uint64_t)rows * (uint32_t)t / (uint32_t)n_threads);
            uint32_t y_hi = row0 + (uint32_t)((uint64_t)rows * ((uint32_t)t + 1) / (uint32_t)n_threads);
            if (y_hi > y_lo)
                raster_band(sp, v, &u, strict_exp, is_unorm, is_f16, width, y_lo, y_hi, row0, acc, &nb, &na, NULL, 0, 0);
        }
        st.bbox_pixels += nb; st.alive_pixels += na;
        free(sp); free(keys); free(idx);
    }
    size_t npx = (size_t)rows * width;
    if (is_unorm) {
        uint8_t* out = (uint8_t*)target;
        int ri = fmt_is_bgra(target_format) ? 2 : 0, bi = 2 - ri;
        for (size_t i = 0; i < npx; i++) {
            out[i * 4 + ri] = (uint8_t)acc[i * 4 + 0];
            out[i * 4 + 1] = (uint8_t)acc[i * 4 + 1];
            out[i * 4 + bi] = (uint8_t)acc[i * 4 + 2];
            out[i * 4 + 3] = 255; /* alpha: 1*(1-a)+a == 1 after quantisation */
        }
    } else if (is_f16) {
        uint16_t* out = (uint16_t*)target;
        for (size_t i = 0; i < npx * 4; i++) out[i] = f32_to_f16(acc[i]);
    } else {
        memcpy(target, acc, npx * 4 * sizeof(float));
    }
    if (stats) *stats = st;
    free(acc);
}

/* squared distance from p to the segment a-b; every step one IEEE op in this order */
static float so_seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    float ex = bx - ax, ey = by - ay, wx = px - ax, wy = py - ay;
    float len2 = ex * ex + ey * ey;
    float t = 0.0f;
    if (len2 > 0.0f) t = fminf(fmaxf((wx * ex + wy * ey) / len2, 0.0f), 1.0f);
    float dx = wx - t * ex, dy = wy - t * ey;
    return dx * dx + dy * dy;
}

/* viewport selection with the brush mask: viewport.wesl:37-69 fed by viewport_texture_brush.wesl, which draws a disc of
 * `radius` at both ends of a stroke segment (frag_main discards dot(uv,uv) > 1) and a quad of half-width `radius`
 * between them: a capsule.  Texel (ix,iy) is taken as set iff its centre is within `radius` of a segment of the
 * polyline; successive strokes accumulate in the texture (accumulate != 0: OR into dest). */
void so_select_brush(const SoModel* model, const SoCameraPod* cam, const float* points_xy, uint32_t n_points,
                     float radius, int accumulate, uint32_t* dest) {
    SoGaussianTransformPod gt = { 1.0f, 0, 0, 0, 255 };
    Uniforms u = make_uniforms(cam, &model->model_transform, &gt);
    uint32_t n = model->n, words = (n + 31u) / 32u;
    uint32_t stride = so_pod_stride(model->sh_fmt, model->cov_fmt);
    float r2 = radius * radius;
    for (uint32_t w = 0; w < words; w++) {
        uint32_t bits = 0;
        for (uint32_t b = 0; b < 32 && w * 32u + b < n; b++) {
            const uint8_t* pod = (const uint8_t*)model->pods + (size_t)stride * (w * 32u + b);
            float pos[3]; memcpy(pos, pod, 12);
            float world[3], clip[4];
            project_centre(&u, pos, world, clip);
            float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
            if (cull(nx, ny, nz)) continue;
            float tx = (nx * 1.0f + 1.0f) * u.size[0] * 0.5f;
            float ty = (ny * -1.0f + 1.0f) * u.size[1] * 0.5f;
            int32_t ix = (int32_t)tx, iy = (int32_t)ty;
            if (ix < 0 || iy < 0 || ix >= (int32_t)u.size[0] || iy >= (int32_t)u.size[1]) continue;
            float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
            int sel = 0;
            if (n_points == 1) sel = so_seg_dist2(px, py, points_xy[0], points_xy[1], points_xy[0], points_xy[1]) <= r2;
            for (uint32_t i = 0; i + 1 < n_points && !sel; i++)
                sel = so_seg_dist2(px, py, points_xy[2 * i], points_xy[2 * i + 1], points_xy[2 * i + 2], points_xy[2 * i + 3]) <= r2;
            if (sel) bits |= 1u << b;
        }
        dest[w] = accumulate ? (dest[w] | bits) : bits;
    }
}

/* target <-> the f32 accumulator the blend runs in (unorm8: 0..255 units) */
static float* acc_load(int target_format, int load, const void* target, size_t npx) {
    int is_unorm = fmt_is_unorm(target_format) ? (fmt_is_srgb(target_format) ? 2 : 1) : 0;
    int is_f16 = target_format == SO_TARGET_RGBA16F;
    int ri = fmt_is_bgra(target_format) ? 2 : 0, bi = 2 - ri;
    float* acc = (float*)malloc((npx ? npx : 1) * 4 * sizeof(float));
    for (size_t i = 0; i < npx; i++) {
        if (!load) {
            acc[i * 4 + 0] = acc[i * 4 + 1] = acc[i * 4 + 2] = 0.0f; acc[i * 4 + 3] = 1.0f;
        } else if (is_unorm) { /* 0..255 units, as the blend keeps them */
            const uint8_t* t = (const uint8_t*)target + i * 4;
            acc[i * 4 + 0] = t[ri]; acc[i * 4 + 1] = t[1]; acc[i * 4 + 2] = t[bi]; acc[i * 4 + 3] = 1.0f;
        } else if (is_f16) {
            const uint16_t* t = (const uint16_t*)target + i * 4;
            for (int c = 0; c < 4; c++) acc[i * 4 + c] = f16_to_f32(t[c]);
        } else {
            memcpy(acc + i * 4, (const float*)target + i * 4, 16);
        }
    }
    return acc;
}

static void acc_store(int target_format, const float* acc, void* target, size_t npx) {
    int is_unorm = fmt_is_unorm(target_format) ? (fmt_is_srgb(target_format) ? 2 : 1) : 0;
    int is_f16 = target_format == SO_TARGET_RGBA16F;
    int ri = fmt_is_bgra(target_format) ? 2 : 0, bi = 2 - ri;
    if (is_unorm) {
        uint8_t* out = (uint8_t*)target;
        for (size_t i = 0; i < npx; i++) {
            out[i * 4 + ri] = (uint8_t)acc[i * 4 + 0];
            out[i * 4 + 1] = (uint8_t)acc[i * 4 + 1];
            out[i * 4 + bi] = (uint8_t)acc[i * 4 + 2];
            out[i * 4 + 3] = 255;
        }
    } else if (is_f16) {
        uint16_t* out = (uint16_t*)target;
        for (size_t i = 0; i < npx * 4; i++) out[i] = f32_to_f16(acc[i]);
    } else {
        memcpy(target, acc, npx * 4 * sizeof(float));
    }
}

/* one indirect instanced draw of `count` instances, instance i = Gaussian indices[i], in that order */
static void draw_instances(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
                           const uint32_t* indices, uint32_t count, int clip_quads, int target_format, int strict_exp,
                           float* acc, float* depth, int depth_compare, int depth_write, int n_threads) {
    uint32_t width = (uint32_t)cam->size[0], height = (uint32_t)cam->size[1];
    int is_unorm = fmt_is_unorm(target_format) ? (fmt_is_srgb(target_format) ? 2 : 1) : 0;
    int is_f16 = target_format == SO_TARGET_RGBA16F;
    if (n_threads <= 0) n_threads = so_max_threads();
    SoSplat* sp = (SoSplat*)malloc((size_t)(count ? count : 1) * sizeof(SoSplat));
    so_project(model, cam, gt, indices, count, sp);
    Uniforms u = make_uniforms(cam, &model->model_transform, gt);
    if (clip_quads) {
        /* The four vertices of a quad share clip z and w (render.wesl:123), so the fixed-function clipper keeps or drops
         * the quad as a whole: kept iff w > 0 and 0 <= z <= w.  x/y are clipped per pixel by the viewport. */
        uint32_t stride = so_pod_stride(model->sh_fmt, model->cov_fmt);
        for (uint32_t i = 0; i < count; i++) {
            float pos[3], world[3], clip[4];
            memcpy(pos, (const uint8_t*)model->pods + (size_t)stride * indices[i], 12);
            project_centre(&u, pos, world, clip);
            float nz = clip[2] / clip[3];
            if (!(clip[3] > 0.0f && nz >= 0.0f && nz <= 1.0f)) sp[i].valid = 0;
        }
    }
    uint64_t nb = 0, na = 0;
    #pragma omp parallel for schedule(static) num_threads(n_threads) reduction(+ : nb, na)
    for (int t = 0; t < n_threads; t++) {
        uint32_t y_lo = (uint32_t)((uint64_t)height * (uint32_t)t / (uint32_t)n_threads);
        uint32_t y_hi = (uint32_t)((uint64_t)height * ((uint32_t)t + 1) / (uint32_t)n_threads);
        if (y_hi > y_lo)
            raster_band(sp, count, &u, strict_exp, is_unorm, is_f16, width, y_lo, y_hi, 0, acc, &nb, &na, depth, depth_compare, depth_write);
    }
    free(sp);
}

void so_render_pass(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
                    int target_format, int strict_exp, int load, void* target, float* depth, int depth_compare,
                    int depth_write, int n_threads) {
    size_t npx = (size_t)(uint32_t)cam->size[1] * (uint32_t)cam->size[0];
    float* acc = acc_load(target_format, load, target, npx);
    uint32_t n = model->n;
    uint32_t cap = ((n + 3839u) / 3840u) * 3840u;
    uint32_t* idx = (uint32_t*)malloc((size_t)(cap ? cap : 1) * 4);
    float* keys = (float*)malloc((size_t)(cap ? cap : 1) * 4);
    uint32_t v = so_preprocess(model, cam, gt, idx, keys, NULL, NULL, NULL);
    so_radix_sort((uint32_t*)keys, idx, v);
    draw_instances(model, cam, gt, idx, v, 0, target_format, strict_exp, acc, depth, depth_compare, depth_write, n_threads);
    free(keys); free(idx);
    acc_store(target_format, acc, target, npx);
    free(acc);
}

/* Renderer<G, ()>::render / render_with_pass on a caller's bind group (src/renderer.rs:321-356): instances
 * 0..count-1 of `indices` in order, no preprocess, no sort. */
void so_draw(const SoModel* model, const SoCameraPod* cam, const SoGaussianTransformPod* gt,
             const uint32_t* indices, uint32_t count, int target_format, int strict_exp, int load, void* target,
             float* depth, int depth_compare, int depth_write, int n_threads) {
    size_t npx = (size_t)(uint32_t)cam->size[1] * (uint32_t)cam->size[0];
    float* acc = acc_load(target_format, load, target, npx);
    draw_instances(model, cam, gt, indices, count, 1, target_format, strict_exp, acc, depth, depth_compare, depth_write, n_threads);
    acc_store(target_format, acc, target, npx);
    free(acc);
}

/* Number of bytes x for which the product's division-free x/255 (one Newton step, fmaf) differs
 * from the correctly rounded x / 255.0f.  Must be 0: lets the kernel avoid the IEEE div sequence. */
int so_unorm8_newton_mismatches(void) {
    int bad = 0;
    for (int x = 0; x < 256; x++) {
        float fx = (float)x;
        float q = fx * 0.00392156886f;
        float r = fmaf(-q, 255.0f, fx);
        float y = fmaf(r, 0.00392156886f, q);
        if (y != fx / 255.0f) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------ viewport selection (f1) */

void so_select_rect(const SoModel* model, const SoCameraPod* cam, float x0, float y0, float x1, float y1,
                    uint32_t* dest) {
    SoGaussianTransformPod gt = { 1.0f, 0, 0, 0, 255 };
    Uniforms u = make_uniforms(cam, &model->model_transform, &gt);
    uint32_t n = model->n, words = (n + 31u) / 32u;
    uint32_t stride = so_pod_stride(model->sh_fmt, model->cov_fmt);
    for (uint32_t w = 0; w < words; w++) {
        uint32_t bits = 0;
        for (uint32_t b = 0; b < 32 && w * 32u + b < n; b++) {
            const uint8_t* pod = (const uint8_t*)model->pods + (size_t)stride * (w * 32u + b);
            float pos[3]; memcpy(pos, pod, 12);
            float world[3], clip[4];
            project_centre(&u, pos, world, clip);
            float nx = clip[0] / clip[3], ny = clip[1] / clip[3], nz = clip[2] / clip[3];
            if (cull(nx, ny, nz)) continue; /* viewport.wesl:55-58: bit cleared */
            /* ndc_to_camera_texture + vec2<i32>() truncation: camera.wesl:18-20, viewport.wesl:60 */
            float tx = (nx * 1.0f + 1.0f) * u.size[0] * 0.5f;
            float ty = (ny * -1.0f + 1.0f) * u.size[1] * 0.5f;
            int32_t ix = (int32_t)tx, iy = (int32_t)ty;
            if (ix < 0 || iy < 0 || ix >= (int32_t)u.size[0] || iy >= (int32_t)u.size[1]) continue; /* OOB textureLoad = 0 */
            /* rectangle mask texel (ix,iy) covered iff its centre lies inside [x0,x1) x [y0,y1) */
            float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
            if (px >= x0 && px < x1 && py >= y0 && py < y1) bits |= 1u << b;
        }
        dest[w] = bits;
    }
}
