// sb_host.cpp — host-only part of the C ABI: pod packers, camera/transform pods, PLY reader.
//
// These replace the CPU-side work the reference does outside the per-frame path
// (core::GaussiansBuffer::new packing, CameraPod::new, ModelTransformPod::new,
// GaussianTransformPod::new, Gaussians::read_from_file) — SURVEY.md §8 row f2.
// Compiled with -ffp-contract=off: every f32 operation is rounded on its own.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <algorithm>
#include <zlib.h>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#endif

#include "../../include/splat_b200.h"

namespace {

uint16_t float_to_half(float f) {  // round-to-nearest-even, IEEE binary16
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t a = x & 0x7fffffffu;
    if (a >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (a > 0x7f800000u ? 0x200u : 0u));
    if (a >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);
    if (a < 0x33000001u) return (uint16_t)sign;
    const int e = (int)(a >> 23) - 127;
    const uint32_t m = (a & 0x7fffffu) | 0x800000u;
    if (e < -14) {
        const uint32_t shift = (uint32_t)(13 + (-14 - e));
        uint32_t out = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (out & 1u))) out++;
        return (uint16_t)(sign | out);
    }
    uint32_t out = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ffu);
    const uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (out & 1u))) out++;
    return (uint16_t)(sign | out);
}

// n floats -> n halves, the same round-to-nearest-even as float_to_half.  On x86 with F16C (every host this library has met) eight at a
// time in hardware — the software conversion made the half-precision pod formats the slowest to pack (1.6 M Gaussians/s against 7.3 M/s
// for single precision on eight cores); NaNs (whose payload the hardware keeps and float_to_half drops) go through the software path.
#if defined(__x86_64__) || defined(__i386__)
__attribute__((target("avx,f16c"))) void floats_to_halves_f16c(const float* src, uint16_t* dst, int n) {
    int k = 0;
    bool any_nan = false;
    for (; k + 8 <= n; k += 8) {
        const __m256 v = _mm256_loadu_ps(src + k);
        any_nan = any_nan || _mm256_movemask_ps(_mm256_cmp_ps(v, v, _CMP_UNORD_Q)) != 0;
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + k), _mm256_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    }
    for (; k < n; k++) dst[k] = float_to_half(src[k]);
    if (any_nan)
        for (k = 0; k < n; k++)
            if (src[k] != src[k]) dst[k] = float_to_half(src[k]);
}
#endif
void floats_to_halves(const float* src, uint16_t* dst, int n) {
#if defined(__x86_64__) || defined(__i386__)
    static const bool f16c = __builtin_cpu_supports("f16c") && __builtin_cpu_supports("avx");
    if (f16c) {
        floats_to_halves_f16c(src, dst, n);
        return;
    }
#endif
    for (int k = 0; k < n; k++) dst[k] = float_to_half(src[k]);
}

float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do {
                man <<= 1;
                e++;
            } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(112 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

uint16_t half_next(uint16_t h, bool up) {
    const bool neg = (h & 0x8000u) != 0;
    if ((h & 0x7fffu) == 0) return up ? 0x0001u : 0x8001u;
    return (uint16_t)((up != neg) ? h + 1 : h - 1);
}

uint32_t sh_size(int sh) { return sh == SB_SH_SINGLE ? 180u : sh == SB_SH_HALF ? 92u : sh == SB_SH_NORM8 ? 52u : 0u; }
uint32_t cov_size(int cov) { return cov == SB_COV_SINGLE ? 24u : cov == SB_COV_HALF ? 12u : 28u; }

// R = Mat3::from_quat (glam), then cov3d = (R S)(R S)^T upper triangle
void covariance(const float q[4], const float s[3], float out[6]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2;
    const float wx = w * x2, wy = w * y2, wz = w * z2;
    const float r[3][3] = {// r[col][row]
                           {1.0f - (yy + zz), xy + wz, xz - wy},
                           {xy - wz, 1.0f - (xx + zz), yz + wx},
                           {xz + wy, yz - wx, 1.0f - (xx + yy)}};
    float m[3][3];
    for (int c = 0; c < 3; c++)
        for (int rr = 0; rr < 3; rr++) m[rr][c] = r[c][rr] * s[c];
    int k = 0;
    for (int i = 0; i < 3; i++)
        for (int j = i; j < 3; j++) out[k++] = (m[i][0] * m[j][0] + m[i][1] * m[j][1]) + m[i][2] * m[j][2];
}

bool valid_fmt(int sh, int cov) { return sh >= 0 && sh <= 3 && cov >= 0 && cov <= 2; }

}  // namespace

extern "C" {

const char* sb_version(void) { return "splat_b200 0.1.0 (sm_100a)"; }

const char* sb_status_string(SbStatus s) {
    switch (s) {
        case SB_OK: return "ok";
        case SB_ERR_INVALID_ARG: return "invalid argument";
        case SB_ERR_CUDA: return "CUDA error";
        case SB_ERR_MODEL_TOO_LARGE: return "model size exceeds the device limit";
        case SB_ERR_MODEL_NOT_FOUND: return "model not found";
        case SB_ERR_BAD_BUFFER_SIZE: return "buffer size mismatch";
        case SB_ERR_IO: return "I/O error";
        case SB_ERR_OVERFLOW: return "tile-duplicate capacity exceeded";
        default: return "unknown status";
    }
}

uint32_t sb_pod_stride(int32_t sh_fmt, int32_t cov_fmt) {
    if (!valid_fmt(sh_fmt, cov_fmt)) return 0;
    return (16u + sh_size(sh_fmt) + cov_size(cov_fmt) + 15u) & ~15u;
}

uint32_t sb_padded_key_count(uint32_t n) { return (uint32_t)(((uint64_t)n + 3839u) / 3840u * 3840u); }

uint64_t sb_keys_buffer_size_bytes(uint32_t n) {
    // keys_buffer_size(n) * RS_KEYVAL_SIZE(4) * BYTES_PER_PAYLOAD_ELEM(4): radix_sorter.rs:922-939
    return (uint64_t)sb_padded_key_count(n) * 4u * 4u;
}

SbStatus sb_pack_gaussians(const SbGaussian* src, uint64_t n, int32_t sh_fmt, int32_t cov_fmt, void* out) {
    if ((!src && n) || (!out && n) || !valid_fmt(sh_fmt, cov_fmt)) return SB_ERR_INVALID_ARG;
    const uint32_t stride = sb_pod_stride(sh_fmt, cov_fmt);
    uint8_t* base = static_cast<uint8_t*>(out);
    std::memset(base, 0, (size_t)stride * n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        const SbGaussian& g = src[i];
        uint8_t* p = base + (size_t)stride * i;
        std::memcpy(p, g.pos, 12);
        std::memcpy(p + 12, g.color, 4);
        uint8_t* q = p + 16;
        if (sh_fmt == SB_SH_SINGLE) {
            std::memcpy(q, g.sh, 180);
        } else if (sh_fmt == SB_SH_HALF) {
            uint16_t h[45];
            floats_to_halves(g.sh, h, 45);
            std::memcpy(q, h, sizeof h);
        } else if (sh_fmt == SB_SH_NORM8) {
            float lo = g.sh[0], hi = g.sh[0];
            for (int k = 1; k < 45; k++) {
                lo = g.sh[k] < lo ? g.sh[k] : lo;
                hi = g.sh[k] > hi ? g.sh[k] : hi;
            }
            uint16_t hlo = float_to_half(lo), hhi = float_to_half(hi);
            if (half_to_float(hlo) > lo) hlo = half_next(hlo, false);
            if (half_to_float(hhi) < hi) hhi = half_next(hhi, true);
            const float flo = half_to_float(hlo), fhi = half_to_float(hhi);
            std::memcpy(q, &hlo, 2);
            std::memcpy(q + 2, &hhi, 2);
            const float range = fhi - flo;
            for (int k = 0; k < 45; k++) {
                const float t = range > 0.0f ? (g.sh[k] - flo) / range : 0.0f;
                float v = std::rint(t * 255.0f);
                v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
                q[4 + k] = (uint8_t)v;
            }
        }
        uint8_t* c = q + sh_size(sh_fmt);
        if (cov_fmt == SB_COV_ROT_SCALE) {
            std::memcpy(c, g.rot, 16);
            std::memcpy(c + 16, g.scale, 12);
        } else {
            float cov[6];
            covariance(g.rot, g.scale, cov);
            if (cov_fmt == SB_COV_SINGLE) {
                std::memcpy(c, cov, 24);
            } else {
                uint16_t h[6];
                floats_to_halves(cov, h, 6);
                std::memcpy(c, h, sizeof h);
            }
        }
    }
    return SB_OK;
}

SbStatus sb_camera_pod(const float pos[3], float yaw, float pitch, float z_near, float z_far, float vertical_fov,
                       uint32_t width, uint32_t height, SbCameraPod* out) {
    if (!pos || !out || width == 0 || height == 0) return SB_ERR_INVALID_ARG;
    // Camera::get_forward (src/camera.rs:71-77)
    float f[3] = {std::cos(pitch) * std::sin(yaw), std::sin(pitch), std::cos(pitch) * std::cos(yaw)};
    // glam Mat4::look_to_rh(eye, dir, Vec3::Y)
    const float fl = 1.0f / std::sqrt((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
    for (float& v : f) v *= fl;
    const float up[3] = {0.0f, 1.0f, 0.0f};
    float s[3] = {f[1] * up[2] - up[1] * f[2], f[2] * up[0] - up[2] * f[0], f[0] * up[1] - up[0] * f[1]};
    const float sl = 1.0f / std::sqrt((s[0] * s[0] + s[1] * s[1]) + s[2] * s[2]);
    for (float& v : s) v *= sl;
    const float u[3] = {s[1] * f[2] - f[1] * s[2], s[2] * f[0] - f[2] * s[0], s[0] * f[1] - f[0] * s[1]};
    const float ds = (pos[0] * s[0] + pos[1] * s[1]) + pos[2] * s[2];
    const float du = (pos[0] * u[0] + pos[1] * u[1]) + pos[2] * u[2];
    const float df = (pos[0] * f[0] + pos[1] * f[1]) + pos[2] * f[2];
    const float view[16] = {s[0], u[0], -f[0], 0.0f, s[1], u[1], -f[1], 0.0f, s[2], u[2], -f[2], 0.0f, -ds, -du, df, 1.0f};
    std::memcpy(out->view, view, sizeof view);
    // glam Mat4::perspective_rh (depth 0..1), aspect = w/h (src/buffer/camera.rs:72-80)
    const float aspect = (float)width / (float)height;
    const float sn = std::sin(0.5f * vertical_fov), cs = std::cos(0.5f * vertical_fov);
    const float h = cs / sn, w = h / aspect, r = z_far / (z_near - z_far);
    const float proj[16] = {w, 0, 0, 0, 0, h, 0, 0, 0, 0, r, -1.0f, 0, 0, r * z_near, 0};
    std::memcpy(out->proj, proj, sizeof proj);
    out->size[0] = (float)width;
    out->size[1] = (float)height;
    out->_padding[0] = out->_padding[1] = 0;
    return SB_OK;
}

SbStatus sb_model_transform_pod(const float pos[3], const float rot[4], const float scale[3], SbModelTransformPod* out) {
    if (!pos || !rot || !scale || !out) return SB_ERR_INVALID_ARG;
    std::memset(out, 0, sizeof *out);
    std::memcpy(out->pos, pos, 12);
    std::memcpy(out->rot, rot, 16);
    std::memcpy(out->scale, scale, 12);
    return SB_OK;
}

SbStatus sb_gaussian_transform_pod(float size, int32_t display_mode, int32_t sh_deg, int32_t no_sh0, float max_std_dev,
                                   SbGaussianTransformPod* out) {
    if (!out || display_mode < 0 || display_mode > 2) return SB_ERR_INVALID_ARG;
    if (sh_deg < 0 || sh_deg > 3) return SB_ERR_INVALID_ARG;                     // GaussianShDegree::new -> None
    if (!(max_std_dev >= 0.0f && max_std_dev <= 3.0f)) return SB_ERR_INVALID_ARG;  // GaussianMaxStdDev::new -> None
    out->size = size;
    out->display_mode = (uint8_t)display_mode;
    out->sh_deg = (uint8_t)sh_deg;
    out->no_sh0 = no_sh0 ? 1 : 0;
    out->max_std_dev = (uint8_t)std::rint(max_std_dev / 3.0f * 255.0f);
    return SB_OK;
}

// INRIA 3DGS .ply (binary little endian, float properties), as coverage/model.ply.
SbStatus sb_read_ply(const char* path, SbGaussian** out, uint64_t* n_out) {
    if (!path || !out || !n_out) return SB_ERR_INVALID_ARG;
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return SB_ERR_IO;
    std::vector<std::string> props;
    uint64_t n = 0;
    bool binary_le = false, in_vertex = false, ok = false;
    char line[512];
    while (std::fgets(line, sizeof line, fp)) {
        std::string l(line);
        while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
        if (l.rfind("format binary_little_endian", 0) == 0) binary_le = true;
        if (l.rfind("element ", 0) == 0) {
            in_vertex = l.rfind("element vertex ", 0) == 0;
            if (in_vertex) n = std::strtoull(l.c_str() + 15, nullptr, 10);
        } else if (l.rfind("property ", 0) == 0 && in_vertex) {
            if (l.rfind("property float ", 0) != 0 && l.rfind("property float32 ", 0) != 0) {
                std::fclose(fp);
                return SB_ERR_IO;
            }
            props.push_back(l.substr(l.rfind(' ') + 1));
        } else if (l == "end_header") {
            ok = true;
            break;
        }
    }
    if (!ok || !binary_le || props.empty()) {
        std::fclose(fp);
        return SB_ERR_IO;
    }
    auto find = [&](const std::string& name) -> int {
        for (size_t i = 0; i < props.size(); i++)
            if (props[i] == name) return (int)i;
        return -1;
    };
    const int ix = find("x"), iy = find("y"), iz = find("z"), iop = find("opacity");
    int idc[3], isc[3], irot[4], irest[45];
    bool have_all = ix >= 0 && iy >= 0 && iz >= 0 && iop >= 0;
    for (int c = 0; c < 3; c++) have_all = ((idc[c] = find("f_dc_" + std::to_string(c))) >= 0) && have_all;
    for (int c = 0; c < 3; c++) have_all = ((isc[c] = find("scale_" + std::to_string(c))) >= 0) && have_all;
    for (int c = 0; c < 4; c++) have_all = ((irot[c] = find("rot_" + std::to_string(c))) >= 0) && have_all;
    // f_rest_*: channel-major, K = count / 3 coefficients per channel (K = 0, 3, 8 or 15 for SH degree 0..3).  The set must
    // be contiguous from f_rest_0 and a whole number of coefficients per channel; anything else is rejected, not guessed at.
    int n_rest = 0;
    for (int c = 0; c < 45; c++) {
        irest[c] = find("f_rest_" + std::to_string(c));
        if (irest[c] >= 0) {
            if (n_rest != c) have_all = false;  // a hole in the set
            n_rest = c + 1;
        }
    }
    if (find("f_rest_45") >= 0 || n_rest % 3 != 0) have_all = false;
    const int rest_per_channel = n_rest / 3;
    if (!have_all) {  // every one of x/y/z/opacity/f_dc_0..2/scale_0..2/rot_0..3 is required
        std::fclose(fp);
        return SB_ERR_IO;
    }
    const size_t np = props.size();
    // the vertex count of the header must fit the file: never trust it for the allocation
    {
        const long data_start = std::ftell(fp);
        if (data_start < 0 || std::fseek(fp, 0, SEEK_END) != 0) {
            std::fclose(fp);
            return SB_ERR_IO;
        }
        const long file_end = std::ftell(fp);
        if (file_end < data_start || std::fseek(fp, data_start, SEEK_SET) != 0 ||
            n > (uint64_t)(file_end - data_start) / (np * 4)) {
            std::fclose(fp);
            return SB_ERR_IO;
        }
    }
    SbGaussian* g = static_cast<SbGaussian*>(std::calloc(n ? n : 1, sizeof(SbGaussian)));
    if (!g) {
        std::fclose(fp);
        return SB_ERR_IO;
    }
    const float SH_C0 = 0.2820948f;
    auto convert = [&](const float* row, SbGaussian& o) {
        o.pos[0] = row[ix]; o.pos[1] = row[iy]; o.pos[2] = row[iz];
        for (int c = 0; c < 3; c++) {
            float v = (0.5f + SH_C0 * row[idc[c]]) * 255.0f;
            v = !(v > 0.0f) ? 0.0f : (v > 255.0f ? 255.0f : v);
            o.color[c] = (uint8_t)v;
        }
        float a = (1.0f / (1.0f + std::exp(-row[iop]))) * 255.0f;
        a = !(a > 0.0f) ? 0.0f : (a > 255.0f ? 255.0f : a);
        o.color[3] = (uint8_t)a;
        for (int k = 0; k < rest_per_channel; k++)  // coefficients past the file's SH degree stay 0
            for (int c = 0; c < 3; c++) o.sh[k * 3 + c] = row[irest[c * rest_per_channel + k]];
        for (int c = 0; c < 3; c++) o.scale[c] = std::exp(row[isc[c]]);
        const float qx = row[irot[1]], qy = row[irot[2]], qz = row[irot[3]], qw = row[irot[0]];
        const float inv = 1.0f / std::sqrt(((qx * qx + qy * qy) + qz * qz) + qw * qw);
        o.rot[0] = qx * inv; o.rot[1] = qy * inv; o.rot[2] = qz * inv; o.rot[3] = qw * inv;
    };
    // blocks of rows: one read, then the rows of the block converted on all host cores (the per-row conversion — four exponentials —
    // is what a 6 M-Gaussian file spends its load time on, not the read)
    const uint64_t kBlockRows = 1u << 16;
    std::vector<float> block((size_t)std::min<uint64_t>(n ? n : 1, kBlockRows) * np);
    for (uint64_t base = 0; base < n; base += kBlockRows) {
        const uint64_t cnt = std::min<uint64_t>(kBlockRows, n - base);
        if (std::fread(block.data(), np * 4, cnt, fp) != cnt) {
            std::free(g);
            std::fclose(fp);
            return SB_ERR_IO;
        }
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < (int64_t)cnt; j++) convert(block.data() + (size_t)j * np, g[base + (uint64_t)j]);
    }
    std::fclose(fp);
    *out = g;
    *n_out = n;
    return SB_OK;
}

// Niantic .spz (gzip stream; versions 2 and 3), the second source `Gaussians::read_from_file` tries (examples/simple.rs:157-160:
// `[GaussiansSource::Ply, GaussiansSource::Spz]`).  The decoder lives in the un-vendored wgpu-3dgs-core; this restates the
// published container: a 16-byte header {magic 'NGSP', version, numPoints, shDegree, fractionalBits, flags, reserved} followed by
// positions (3 x 24-bit fixed point), alphas (u8), colours (3 x u8), scales (3 x u8), rotations (3 x u8 in v2, 4 bytes
// "smallest three" in v3) and SH (shDim x 3 x u8), each as one array over all points.  Mapped onto `Gaussian` like the PLY path:
// colour = (0.5 + SH_C0 * dc) * 255 truncated to u8, alpha = the stored byte, scale = exp(log-scale), rotation normalised xyzw.
// File coordinates are kept as they are (no axis convention change): RECALLED layout, unverifiable here.
SbStatus sb_read_spz(const char* path, SbGaussian** out, uint64_t* n_out) {
    if (!path || !out || !n_out) return SB_ERR_INVALID_ARG;
    gzFile gz = gzopen(path, "rb");
    if (!gz) return SB_ERR_IO;
    std::vector<uint8_t> data;
    {
        uint8_t buf[1 << 16];
        int got;
        while ((got = gzread(gz, buf, sizeof buf)) > 0) {
            data.insert(data.end(), buf, buf + got);
            if (data.size() > (size_t)16 << 30) break;  // refuse absurd streams
        }
        const bool bad = got < 0;
        gzclose(gz);
        if (bad) return SB_ERR_IO;
    }
    if (data.size() < 16) return SB_ERR_IO;
    uint32_t magic, version, n;
    std::memcpy(&magic, &data[0], 4);
    std::memcpy(&version, &data[4], 4);
    std::memcpy(&n, &data[8], 4);
    const uint32_t sh_degree = data[12], frac_bits = data[13];
    if (magic != 0x5053474eu || (version != 2u && version != 3u) || sh_degree > 3u || frac_bits > 30u) return SB_ERR_IO;
    static const uint32_t kShDim[4] = {0, 3, 8, 15};
    const uint32_t sh_dim = kShDim[sh_degree];
    const uint32_t rot_bytes = version == 3u ? 4u : 3u;
    const uint64_t per_point = 9u + 1u + 3u + 3u + rot_bytes + (uint64_t)sh_dim * 3u;
    if (data.size() - 16 < per_point * n) return SB_ERR_IO;  // the header's count must fit the stream
    const uint8_t* pos = &data[16];
    const uint8_t* alpha = pos + (size_t)n * 9;
    const uint8_t* col = alpha + (size_t)n;
    const uint8_t* scl = col + (size_t)n * 3;
    const uint8_t* rot = scl + (size_t)n * 3;
    const uint8_t* sh = rot + (size_t)n * rot_bytes;
    SbGaussian* g = static_cast<SbGaussian*>(std::calloc(n ? n : 1, sizeof(SbGaussian)));
    if (!g) return SB_ERR_IO;
    const float pos_scale = 1.0f / (float)(1u << frac_bits);
    const float SH_C0 = 0.2820948f, kColorScale = 0.15f;
    // (the points are independent: decoded on all host cores; the inflate above is the serial part)
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        SbGaussian& o = g[i];
        for (int c = 0; c < 3; c++) {
            const uint8_t* b = pos + (size_t)i * 9 + c * 3;
            int32_t v = (int32_t)b[0] | ((int32_t)b[1] << 8) | ((int32_t)b[2] << 16);
            if (v & 0x800000) v |= ~0xffffff;  // sign-extend 24 bits
            o.pos[c] = (float)v * pos_scale;
        }
        for (int c = 0; c < 3; c++) {
            const float dc = ((float)col[(size_t)i * 3 + c] / 255.0f - 0.5f) / kColorScale;
            float v = (0.5f + SH_C0 * dc) * 255.0f;
            v = !(v > 0.0f) ? 0.0f : (v > 255.0f ? 255.0f : v);
            o.color[c] = (uint8_t)v;
        }
        o.color[3] = alpha[i];
        for (int c = 0; c < 3; c++) o.scale[c] = std::exp((float)scl[(size_t)i * 3 + c] / 16.0f - 10.0f);
        float q[4];
        if (version == 2u) {
            const uint8_t* r = rot + (size_t)i * 3;
            float sum = 0.0f;
            for (int c = 0; c < 3; c++) {
                q[c] = (float)r[c] / 127.5f - 1.0f;
                sum += q[c] * q[c];
            }
            q[3] = std::sqrt(std::fmax(0.0f, 1.0f - sum));
        } else {
            uint32_t comp;
            std::memcpy(&comp, rot + (size_t)i * 4, 4);
            const uint32_t largest = comp >> 30;
            float sum = 0.0f;
            for (int k = 3; k >= 0; k--) {
                if ((uint32_t)k == largest) continue;
                const uint32_t mag = comp & 0x1ffu, neg = (comp >> 9) & 1u;
                comp >>= 10;
                q[k] = 0.70710678f * (float)mag / 511.0f;
                if (neg) q[k] = -q[k];
                sum += q[k] * q[k];
            }
            q[largest] = std::sqrt(std::fmax(0.0f, 1.0f - sum));
        }
        const float len = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        const float inv = len > 0.0f ? 1.0f / len : 0.0f;
        for (int c = 0; c < 4; c++) o.rot[c] = q[c] * inv;
        if (!(len > 0.0f)) o.rot[3] = 1.0f;
        for (uint32_t k = 0; k < sh_dim; k++)
            for (int c = 0; c < 3; c++) o.sh[k * 3 + c] = ((float)sh[((size_t)i * sh_dim + k) * 3 + c] - 128.0f) / 128.0f;
    }
    *out = g;
    *n_out = n;
    return SB_OK;
}

void sb_free(void* p) { std::free(p); }

}  // extern "C"
