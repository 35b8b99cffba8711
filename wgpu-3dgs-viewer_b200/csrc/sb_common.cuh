// sb_common.cuh — shared device-side types and helpers (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/splat_b200.h"

namespace sb {

constexpr uint32_t kHistoBlockKvs = 3840;  // wgpu_sort::HISTO_BLOCK_KVS, src/radix_sorter.rs:178
constexpr int kTile = 16;                  // raster tile edge in pixels

// Per-frame uniforms, passed to kernels BY VALUE (__grid_constant__): update_* -> render in
// program order then behaves like queue.write_buffer -> submit with no staging hazards.
// Derived on the host in strict f32 (see make_uniforms in sb_api.cu), mirroring what every
// invocation of the reference shaders recomputes from the three uniform buffers.
struct Uniforms {
    float pv[16];      // proj * view              (camera.wesl:13-15)
    float vm[16];      // view * model_mat         (utils.wesl:37)
    float model[16];   // model_transform_mat
    float sr[9];       // model_scale_rot_mat      (column-major mat3)
    float inv_sr[9];   // model_transform_inv_sr_mat
    float w[9];        // mat3(view)
    float size[2];
    float focal[2];
    float cam_pos[3];
    float std_dev;     // gaussian_transform_max_std_dev
    float gsize;       // gaussian_transform.size
    float color_scale; // 255 for unorm8 targets, 1 for float and sRGB targets
    float color_max;   // source-colour clamp after scaling: 255 (unorm8), 1 (sRGB: fixed-point attachment), +inf (float)
    float cut_k;       // exact alpha cut-off (splat mode on unorm8 targets): 1 / kAlphaCut, 0 = off
    uint32_t mode, sh_deg, no_sh0;
    uint32_t width, height;
    uint32_t tiles_x, tiles_y;
};

// On a unorm8 target a blend with alpha < 0.5/255 is the identity EXACTLY: d is an integer in 0..255, the source colour
// c255 <= 255, so |fma(d, 1-alpha, c255*alpha) - d| <= alpha*255 + 1.6e-5 < 0.5 and rint() returns d.  In splat mode
// alpha = a*exp(-r^2) falls below kAlphaCut for r^2 > ln(a / kAlphaCut); the vertex stage shrinks the splat's cull
// extents (and with them its tile bbox) to that radius, so those fragments are never generated.  Bit-identical output.
constexpr float kAlphaCut = 0.00195f;      // < 0.5/255 = 0.0019608 with room for the rounding of the blend
constexpr float kAlphaCutMargin = 0.01f;   // added to ln(a / kAlphaCut): covers the rounding of r^2 and of exp (1 %)

// Projected splat = everything vert_main (render.wesl:76-130) hands to the fragment stage,
// computed once per visible Gaussian instead of 6x per quad.  48 bytes, 3 x 128-bit.
// The axis rows are stored column-wise, (ax, bx) and (ay, by), so the rasterizer evaluates both
// quad_offset components with one packed FMUL2 + one packed FFMA2 (sb_raster.cu).
struct __align__(16) SplatRec {
    float cx, cy;   // pixel-space centre
    float ax, bx;   // quad_offset.x = dx*ax + dy*ay
    float ay, by;   // quad_offset.y = dx*bx + dy*by
    float ex, ey;   // half extent (pixels) of the alive region: used for warp-level culling
    float r, g;     // colour * color_scale (clamped to 255 on unorm8 targets: source-colour clamp)
    float b, a;
};
// Tile bbox of a splat, kept in a separate 8-byte array so the binning gathers (random, in depth
// order) stay L2-resident: x0 | y0 << 16, x1 | y1 << 16; min > max component-wise => no tiles.
struct __align__(8) TileBox {
    uint32_t tmin, tmax;
};
static_assert(sizeof(SplatRec) == 48, "SplatRec must be 48 bytes");

// ---- strict (individually rounded, never contracted) f32 ops for bit-exact artefacts
__device__ __forceinline__ float smul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ssub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float sdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float ssqrt(float a) { return __fsqrt_rn(a); }

// ---- mbarrier / bulk-copy (TMA) PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking form: mbarrier.try_wait suspends the warp in hardware until the phase completes or the time hint
// expires, so a waiting warp issues a handful of instructions instead of spinning on test_wait (the spin was 27 %
// of all instructions the rasterizer executed: profiles/r01_raster_gather4_kernel_hotspots.txt).
__device__ __forceinline__ bool mbar_try_wait_suspend(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait_suspend(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// TMA tile::gather4: four rows of a 2-D tensor (row indices r0..r3, column 0) -> four consecutive
// rows in shared memory; completion (4 * row bytes) is signalled on the mbarrier.  SASS: UTMALDG.
__device__ __forceinline__ void tma_gather4(uint32_t smem_dst, const void* tensor_map, uint64_t* bar, uint32_t r0, uint32_t r1,
                                            uint32_t r2, uint32_t r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
            "r"(smem_dst),
        "l"(tensor_map), "r"(smem_u32(bar)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// 64-bit status words for decoupled look-back: flag in the high word, value in the low word.
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

constexpr unsigned long long kFlagAggregate = 1ull << 32;
constexpr unsigned long long kFlagPrefix = 2ull << 32;

// Warp-cooperative decoupled look-back; returns the exclusive prefix of `aggregate` over tiles.
__device__ __forceinline__ uint32_t lookback(unsigned long long* status, uint32_t tile, uint32_t aggregate, uint32_t lane) {
    if (lane == 0) st_relaxed_u64(&status[tile], (tile == 0 ? kFlagPrefix : kFlagAggregate) | aggregate);
    uint32_t excl = 0;
    if (tile > 0) {
        int base = (int)tile - 1;
        while (true) {
            int idx = base - (int)lane;
            unsigned long long v;
            do {
                v = idx >= 0 ? ld_relaxed_u64(&status[idx]) : kFlagPrefix;
            } while (__any_sync(0xffffffffu, (v >> 32) == 0));
            uint32_t pm = __ballot_sync(0xffffffffu, (v >> 32) == 2);
            int first = pm ? (__ffs(pm) - 1) : 31;
            uint32_t contrib = ((int)lane <= first) ? (uint32_t)v : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            excl += contrib;
            if (pm) break;
            base -= 32;
        }
        if (lane == 0) st_relaxed_u64(&status[tile], kFlagPrefix | (unsigned long long)(excl + aggregate));
    }
    return excl;
}

}  // namespace sb
