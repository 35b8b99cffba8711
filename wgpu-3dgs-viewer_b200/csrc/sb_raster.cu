// sb_raster.cu — K4/K5/K6: tile binning and the tile rasterizer.
//
// Replaces Renderer::render's instanced-quad draw + fixed-function ALPHA_BLENDING
// (src/renderer.rs:163-195, 284-308; fragment stage src/shader/render.wesl:143-181):
//   K4 dup_count / dup_offsets   depth-ordered splats -> tile boxes in depth order, tile sums, then per-splat offsets (reduce-then-scan:
//                  no look-back chain) and the first splat of every 4096-duplicate output window
//   K5 dup_emit2   (tile id, Gaussian index) duplicates in depth order, OUTPUT-driven: a CTA expands one window into shared memory and
//                  writes it with coalesced 128-bit stores       (SB_BIN=v1: the round-1 chained scan + per-splat emit, kept for A/B)
//      tile sort   stable onesweep over the tile-id bits only (sb_sort.cu): because the input is
//                  already in the reference's draw order, stability alone preserves it per tile
//   K5b ranges     per-tile ranges of the sorted duplicates + the rasterizer's longest-list-first tile schedule
//                  (SB_RASTER_PATH=bulk: also a gathered copy of the records in tile order)
//   K6 raster      one CTA per 16x16 tile; the tile's records are staged into shared memory by TMA gather4 straight
//                  from the per-Gaussian array, culled once per CTA against the sixteen 4x4 blocks, and one pixel
//                  per thread composites back-to-front in exactly the reference's order, re-quantising after every
//                  blend on unorm8 targets.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <string>

#include <algorithm>

#include "sb_internal.h"

namespace sb {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;  // 4096 splats per tile: a quarter of the look-back hops of 1024 (the kernel is bound by the latency of that chain and of its gathers)
constexpr int kScanTile = kScanThreads * kScanItems;

struct BinParams {
    const TileBox* tboxes;
    const uint32_t* sorted_indices;
    const uint32_t* sorted_indices_alt;  // the depth sort's alt payload buffer: holds the result when *sort_parity != 0
    const uint32_t* sort_parity;         // nullable: result is in sorted_indices
    const uint32_t* visible_count;
    uint32_t* dup_offsets;
    TileBox* tboxes_sorted;  // v2 binning: the tile boxes in depth order (written by the scan, streamed by the emit)
    uint32_t tiles_prefixed; // v2 binning: 1 = dup_tiles_kernel turned the tile sums into prefixes; 0 = dup_offsets sums them itself
    uint32_t* win_first;     // v2 binning: [w] = the splat (depth rank) whose run of duplicates covers duplicate 4096 w
    uint32_t win_capacity;   //   entries of win_first
    uint32_t* dup_keys;
    uint32_t* dup_vals;
    uint32_t* dup_count;
    uint32_t* overflow;
    uint32_t* needed_host;  // mapped pinned word: duplicates this frame needs (read by the next render call)
    unsigned long long* scan_status;
    uint32_t* scan_counter;
    uint32_t dup_capacity;
    uint32_t tiles_x;
    uint32_t ty_lo, ty_hi;  // tile rows of the rendered strip (inclusive)
    uint32_t max_visible;   // n: the device-side count is clamped to it, indices >= n name no Gaussian (robust-buffer behaviour)
    uint32_t indices_vec4;  // the index list is 16-byte aligned (a caller's sub-buffer need not be)
};

// tiles of one splat, clipped to the strip's tile rows
__device__ __forceinline__ uint32_t box_tiles(uint2 q, uint32_t ty_lo, uint32_t ty_hi, uint32_t& x0, uint32_t& y0, uint32_t& w) {
    const uint32_t tmin = q.x, tmax = q.y;
    x0 = tmin & 0xffffu;
    y0 = tmin >> 16;
    const uint32_t x1 = tmax & 0xffffu;
    uint32_t y1 = tmax >> 16;
    if (x0 > x1 || y0 > y1) return 0;
    y0 = max(y0, ty_lo);
    y1 = min(y1, ty_hi);
    if (y0 > y1) return 0;
    w = x1 - x0 + 1;
    return w * (y1 - y0 + 1);
}
__device__ __forceinline__ uint32_t splat_tiles(const TileBox* tboxes, uint32_t g, uint32_t ty_lo, uint32_t ty_hi, uint32_t& x0,
                                                uint32_t& y0, uint32_t& w) {
    return box_tiles(__ldg(reinterpret_cast<const uint2*>(&tboxes[g])), ty_lo, ty_hi, x0, y0, w);
}

// Decoupled look-back over 62-bit aggregates (flag in the two top bits): the duplicate total of a frame can exceed 2^32 (millions
// of close-up splats covering thousands of tiles each), and an overflowing frame must still be FLAGGED, never wrapped.
constexpr unsigned long long kFlag62Aggregate = 1ull << 62;
constexpr unsigned long long kFlag62Prefix = 2ull << 62;
constexpr unsigned long long kValue62Mask = (1ull << 62) - 1;
__device__ __forceinline__ unsigned long long lookback62(unsigned long long* status, uint32_t tile, unsigned long long aggregate, uint32_t lane) {
    if (lane == 0) st_relaxed_u64(&status[tile], (tile == 0 ? kFlag62Prefix : kFlag62Aggregate) | aggregate);
    unsigned long long excl = 0;
    if (tile > 0) {
        int base = (int)tile - 1;
        while (true) {
            const int idx = base - (int)lane;
            unsigned long long v;
            do {
                v = idx >= 0 ? ld_relaxed_u64(&status[idx]) : kFlag62Prefix;
            } while (__any_sync(0xffffffffu, (v >> 62) == 0));
            const uint32_t pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
            const int first = pm ? (__ffs(pm) - 1) : 31;
            unsigned long long contrib = ((int)lane <= first) ? (v & kValue62Mask) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            excl += contrib;
            if (pm) break;
            base -= 32;
        }
        if (lane == 0) st_relaxed_u64(&status[tile], kFlag62Prefix | (excl + aggregate));
    }
    return excl;
}

__device__ __forceinline__ uint32_t sat32(unsigned long long v) { return v > 0xffffffffull ? 0xffffffffu : (uint32_t)v; }

// K4: chained scan (decoupled look-back) of tiles-per-splat in depth order.  Sums are carried in 64 bits; the offsets it writes
// saturate at 2^32 - 1 (>= any capacity), so the emit's `offset < capacity` guard drops exactly the duplicates that do not fit.
__global__ void __launch_bounds__(kScanThreads) dup_scan_kernel(const BinParams p) {
    __shared__ unsigned long long warp_sums[kScanThreads / 32];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t v = min(*p.visible_count, p.max_visible);
    if (blockIdx.x * kScanTile >= v && blockIdx.x > 0) return;
    if (tid == 0) s_tile = atomicAdd(p.scan_counter, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * kScanTile + tid * kScanItems;
    const uint32_t* __restrict__ sorted = (p.sort_parity && *p.sort_parity) ? p.sorted_indices_alt : p.sorted_indices;
    uint32_t cnt[kScanItems];
    unsigned long long sum = 0;
    static_assert(kScanItems % 4 == 0, "a thread's run of splats is read and written as 128-bit vectors");
    uint32_t gidx[kScanItems];
    const bool whole = base + kScanItems <= v && p.indices_vec4;  // the run starts at a multiple of kScanItems: 16-byte aligned
    if (whole) {
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++) {
            const uint4 t = reinterpret_cast<const uint4*>(sorted + base)[q];
            gidx[4 * q] = t.x; gidx[4 * q + 1] = t.y; gidx[4 * q + 2] = t.z; gidx[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++) gidx[i] = base + i < v ? sorted[base + i] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        const uint32_t r = base + i;
        uint32_t x0, y0, w;
        cnt[i] = (r < v && gidx[i] < p.max_visible) ? splat_tiles(p.tboxes, gidx[i], p.ty_lo, p.ty_hi, x0, y0, w) : 0u;
        sum += cnt[i];
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    unsigned long long wexcl = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        const unsigned long long t = warp_sums[w];
        if (w < (int)warp) wexcl += t;
        total += t;
    }
    if (warp == 0) {
        const unsigned long long excl = lookback62(p.scan_status, tile, total, lane);
        if (lane == 0) s_base = excl;
    }
    __syncthreads();
    unsigned long long off = s_base + wexcl + inc - sum;
    if (whole) {
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++) {
            uint4 o;
            o.x = sat32(off); off += cnt[4 * q];
            o.y = sat32(off); off += cnt[4 * q + 1];
            o.z = sat32(off); off += cnt[4 * q + 2];
            o.w = sat32(off); off += cnt[4 * q + 3];
            reinterpret_cast<uint4*>(p.dup_offsets + base)[q] = o;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; i++) {
            const uint32_t r = base + i;
            if (r < v) p.dup_offsets[r] = sat32(off);
            off += cnt[i];
        }
    }
    const uint32_t last_tile = v == 0 ? 0 : (v - 1) / kScanTile;
    if (tile == last_tile && tid == 0) {
        const unsigned long long d = s_base + total;
        if (p.needed_host) {
            p.needed_host[0] = sat32(d);                      // duplicates this frame needed (saturated)
            if (d > p.dup_capacity) p.needed_host[1] += 1u;   // overflow events (the host compares with the count it has seen)
        }
        if (d > p.dup_capacity) {
            *p.overflow = 1u;
            *p.dup_count = p.dup_capacity;
        } else {
            *p.overflow = 0u;
            *p.dup_count = (uint32_t)d;
        }
    }
}

// K5: emit (tile id, Gaussian index) in depth order.  One lane per splat; splats covering more
// than 32 tiles are expanded by the whole warp.
__global__ void __launch_bounds__(256) dup_emit_kernel(const BinParams p) {
    const uint32_t v = min(*p.visible_count, p.max_visible);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = gridDim.x * (blockDim.x / 32);
    const uint32_t* __restrict__ sorted = (p.sort_parity && *p.sort_parity) ? p.sorted_indices_alt : p.sorted_indices;
    for (uint32_t wbase = (blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * 32; wbase < v; wbase += warps_total * 32) {
        const uint32_t r = wbase + lane;
        uint32_t g = 0, n = 0, x0 = 0, y0 = 0, w = 1, off = 0;
        if (r < v) {
            g = sorted[r];
            n = g < p.max_visible ? splat_tiles(p.tboxes, g, p.ty_lo, p.ty_hi, x0, y0, w) : 0u;
            off = p.dup_offsets[r];
        }
        if (n > 0 && n <= 32) {
            // row-major walk of the box without a division per tile
            uint32_t key = (y0 - p.ty_lo) * p.tiles_x + x0, col = 0;  // strip-relative tile id: fewer key bits for the tile sort
            for (uint32_t j = 0; j < n; j++) {
                const uint32_t o = off + j;
                if (o < p.dup_capacity && o >= off) {  // (off saturates at 2^32 - 1: off + j must not wrap back into range)
                    p.dup_keys[o] = key;
                    p.dup_vals[o] = g;
                }
                ++key;
                if (++col == w) {
                    col = 0;
                    key += p.tiles_x - w;
                }
            }
        }
        uint32_t big = __ballot_sync(0xffffffffu, n > 32);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const uint32_t bn = __shfl_sync(0xffffffffu, n, src), bg = __shfl_sync(0xffffffffu, g, src);
            const uint32_t bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
            const uint32_t bw = __shfl_sync(0xffffffffu, w, src), boff = __shfl_sync(0xffffffffu, off, src);
            // j / bw by multiplication: one division per big splat instead of one per tile.  floor(j / bw) ==
            // umulhi(j, ceil(2^32 / bw)) whenever j * bw < 2^32 (checked; bw == 1 would need 2^32 as the multiplier)
            const bool mul_ok = bw > 1u && (unsigned long long)bn * bw < (1ull << 32);
            const uint32_t inv = mul_ok ? 0xffffffffu / bw + 1u : 0u;
            for (uint32_t j = lane; j < bn; j += 32) {
                const uint32_t o = boff + j;
                if (o < p.dup_capacity && o >= boff) {
                    const uint32_t q = mul_ok ? __umulhi(j, inv) : j / bw;
                    p.dup_keys[o] = (by0 - p.ty_lo + q) * p.tiles_x + (bx0 + (j - q * bw));
                    p.dup_vals[o] = bg;
                }
            }
        }
    }
}


// ---------------------------------------------------------------- binning v2
// K4' — reduce, then scan: three small kernels WITHOUT a look-back chain.  ncu of the chained scan (v1): 31 % of the stall samples
// are the CTA barrier behind warp 0's look-back and the kernel takes ~65 us for 6 M splats whatever its memory traffic — with several
// hundred tiles in flight almost every predecessor a tile inspects holds an aggregate, not a prefix, so every tile walks back ~14
// windows of 32 at one L2 round trip each while its other warps wait.  Here:
//   dup_count   one CTA per 4096 splats, STRIPED (thread t takes splats t, t + 512, ...): the index loads and the stores of the
//               boxes in depth order are coalesced, a thread's eight tile-box gathers are independent and all in flight together;
//               block sum -> tile_sums[tile].  No ordering between CTAs, no ticket.
//   dup_tiles   one CTA: exclusive scan of the tile sums (64-bit), the frame's duplicate total, overflow flag, host words.
//   dup_offsets one CTA per 4096 splats (blocked): streams the sorted boxes, block scan + its tile's base -> per-splat offsets (saturated)
//               and the first splat of every 4096-duplicate emit window.
constexpr int kEmitWinLog2 = 12;
constexpr int kEmitWin = 1 << kEmitWinLog2;  // duplicates per emit window
constexpr int kScan2Threads = 512;
constexpr int kScan2Items = 8;
constexpr int kScan2Tile = kScan2Threads * kScan2Items;
static_assert(kScan2Tile == kScanTile, "both scans use the same tile (scan_status sizing, launch grid)");

__global__ void __launch_bounds__(kScan2Threads) dup_count_kernel(const BinParams p) {
    __shared__ unsigned long long warp_sums[kScan2Threads / 32];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t v = min(*p.visible_count, p.max_visible);
    const uint32_t tile_base = blockIdx.x * kScan2Tile;
    if (tile_base >= v) return;
    const uint32_t* __restrict__ sorted = (p.sort_parity && *p.sort_parity) ? p.sorted_indices_alt : p.sorted_indices;
    uint32_t g[kScan2Items];
#pragma unroll
    for (int i = 0; i < kScan2Items; i++) {
        const uint32_t r = tile_base + i * kScan2Threads + tid;
        g[i] = r < v ? sorted[r] : 0xffffffffu;
    }
    uint2 q[kScan2Items];
#pragma unroll
    for (int i = 0; i < kScan2Items; i++)
        q[i] = g[i] < p.max_visible ? __ldg(reinterpret_cast<const uint2*>(&p.tboxes[g[i]])) : make_uint2(1u, 0u);  // (1, 0): no tiles
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < kScan2Items; i++) {
        const uint32_t r = tile_base + i * kScan2Threads + tid;
        uint32_t x0, y0, w;
        sum += box_tiles(q[i], p.ty_lo, p.ty_hi, x0, y0, w);
        if (r < v) reinterpret_cast<uint2*>(p.tboxes_sorted)[r] = q[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) warp_sums[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kScan2Threads / 32; w++) t += warp_sums[w];
        p.scan_status[blockIdx.x] = t;  // tile_sums (the look-back words of v1: same buffer, plain values here)
    }
}

// exclusive scan of the tile sums, in place; one CTA (1465 tiles for 6 M splats; a chunk of 1024 per iteration)
__global__ void __launch_bounds__(1024) dup_tiles_kernel(const BinParams p) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long s_carry;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t v = min(*p.visible_count, p.max_visible);
    const uint32_t tiles = (v + kScan2Tile - 1) / kScan2Tile;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < tiles; base += 1024) {
        const uint32_t i = base + tid;
        const unsigned long long x = i < tiles ? p.scan_status[i] : 0ull;
        unsigned long long inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[warp] = inc;
        __syncthreads();
        unsigned long long wexcl = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 32; w++) {
            const unsigned long long t = warp_sums[w];
            if (w < (int)warp) wexcl += t;
            total += t;
        }
        const unsigned long long carry = s_carry;
        if (i < tiles) p.scan_status[i] = carry + wexcl + inc - x;
        __syncthreads();
        if (tid == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (tid == 0) {
        const unsigned long long d = s_carry;
        if (p.needed_host) {
            p.needed_host[0] = sat32(d);                      // duplicates this frame needed (saturated)
            if (d > p.dup_capacity) p.needed_host[1] += 1u;   // overflow events (the host compares with the count it has seen)
        }
        if (d > p.dup_capacity) {
            *p.overflow = 1u;
            *p.dup_count = p.dup_capacity;
        } else {
            *p.overflow = 0u;
            *p.dup_count = (uint32_t)d;
        }
    }
}

__global__ void __launch_bounds__(kScan2Threads) dup_offsets_kernel(const BinParams p) {
    __shared__ unsigned long long warp_sums[kScan2Threads / 32];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t v = min(*p.visible_count, p.max_visible);
    const uint32_t tile_base = blockIdx.x * kScan2Tile;
    if (v == 0 && blockIdx.x == 0 && tid == 0 && !p.tiles_prefixed) {  // empty frame
        if (p.needed_host) p.needed_host[0] = 0u;
        *p.overflow = 0u;
        *p.dup_count = 0u;
    }
    if (tile_base >= v) return;
    // BLOCKED: thread t owns splats [8 t, 8 t + 8) of the tile — 64 contiguous bytes of sorted boxes, four 128-bit loads (a warp
    // covers 2 KB: every sector is used, the later loads of a thread hit the lines its first one brought)
    const uint32_t base = tile_base + tid * kScan2Items;
    uint32_t cnt[kScan2Items];
    if (base + kScan2Items <= v) {
        const uint4* bp4 = reinterpret_cast<const uint4*>(p.tboxes_sorted) + (size_t)base / 2;
        uint4 q4[kScan2Items / 2];
#pragma unroll
        for (int i = 0; i < kScan2Items / 2; i++) q4[i] = __ldg(bp4 + i);
#pragma unroll
        for (int i = 0; i < kScan2Items / 2; i++) {
            uint32_t x0, y0, w;
            cnt[2 * i] = box_tiles(make_uint2(q4[i].x, q4[i].y), p.ty_lo, p.ty_hi, x0, y0, w);
            cnt[2 * i + 1] = box_tiles(make_uint2(q4[i].z, q4[i].w), p.ty_lo, p.ty_hi, x0, y0, w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScan2Items; i++) {
            uint32_t x0, y0, w;
            cnt[i] = base + i < v ? box_tiles(__ldg(reinterpret_cast<const uint2*>(p.tboxes_sorted) + base + i), p.ty_lo, p.ty_hi, x0, y0, w) : 0u;
        }
    }
    // my tile's base: the prefix dup_tiles_kernel left, or (few tiles: 1465 for 6 M splats, 12 KB that sit in L2) the sum of the
    // tile sums in front of mine, read by the whole CTA while the boxes are in flight — one launch and one idle SM-array less
    unsigned long long tile_off;
    const uint32_t tiles = (v + kScan2Tile - 1) / kScan2Tile;
    if (p.tiles_prefixed) {
        tile_off = p.scan_status[blockIdx.x];
    } else {
        unsigned long long part = 0;
        for (uint32_t t = tid; t < blockIdx.x; t += kScan2Threads) part += p.scan_status[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) warp_sums[warp] = part;
        __syncthreads();
        tile_off = 0;
#pragma unroll
        for (int w = 0; w < kScan2Threads / 32; w++) tile_off += warp_sums[w];
        __syncthreads();  // warp_sums is reused by the scan below
        if (blockIdx.x == tiles - 1 && tid == 0) {  // the last tile knows the frame's total
            const unsigned long long d = tile_off + p.scan_status[blockIdx.x];
            if (p.needed_host) {
                p.needed_host[0] = sat32(d);
                if (d > p.dup_capacity) p.needed_host[1] += 1u;
            }
            if (d > p.dup_capacity) {
                *p.overflow = 1u;
                *p.dup_count = p.dup_capacity;
            } else {
                *p.overflow = 0u;
                *p.dup_count = (uint32_t)d;
            }
        }
    }
    static_assert(kScan2Items == 8, "two 128-bit vectors per thread");
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < kScan2Items; i++) sum += cnt[i];
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    unsigned long long wexcl = 0;
#pragma unroll
    for (int w = 0; w < kScan2Threads / 32; w++)
        if (w < (int)warp) wexcl += warp_sums[w];
    unsigned long long off = tile_off + wexcl + inc - sum;
    {
        // the splat whose run covers duplicate 4096 w is the first splat of the emit's window w
        unsigned long long o = off;
#pragma unroll
        for (int i = 0; i < kScan2Items; i++) {
            if (cnt[i]) {
                const unsigned long long w_hi = (o + cnt[i] - 1) >> kEmitWinLog2;
                for (unsigned long long w = (o + kEmitWin - 1) >> kEmitWinLog2; w <= w_hi && w < p.win_capacity; w++) p.win_first[w] = base + i;
            }
            o += cnt[i];
        }
    }
    if (base + kScan2Items <= v) {
        uint4 o0, o1;
        o0.x = sat32(off); off += cnt[0];
        o0.y = sat32(off); off += cnt[1];
        o0.z = sat32(off); off += cnt[2];
        o0.w = sat32(off); off += cnt[3];
        o1.x = sat32(off); off += cnt[4];
        o1.y = sat32(off); off += cnt[5];
        o1.z = sat32(off); off += cnt[6];
        o1.w = sat32(off); off += cnt[7];
        reinterpret_cast<uint4*>(p.dup_offsets + base)[0] = o0;
        reinterpret_cast<uint4*>(p.dup_offsets + base)[1] = o1;
    } else {
#pragma unroll
        for (int i = 0; i < kScan2Items; i++) {
            const uint32_t r = base + i;
            if (r < v) p.dup_offsets[r] = sat32(off);
            off += cnt[i];
        }
    }
}

// K5': OUTPUT-driven emit.  A CTA owns a window of 4096 consecutive duplicates: it finds the splats whose runs meet the window by
// binary search in the offsets, expands them into shared memory (thread per splat; splats with more than 32 duplicates in the
// window by a whole warp) and writes the window with 128-bit coalesced stores.  v1 walked the splats and stored every duplicate
// with its own 4-byte store at a lane-private offset — 19 M sector transactions per frame for 77 MB, L1TEX the busiest unit at
// 75 % — and a view with large splats serialised up to 32 such stores per lane.  Work per CTA is now constant whatever the
// splat sizes, and the boxes arrive in depth order (scan) instead of through a second random gather.
constexpr int kEmitUnroll = 4;
constexpr int kEmitBig = 128;  // a window holds at most 4096 / 33 splats with more than 32 duplicates

__global__ void __launch_bounds__(256) dup_emit2_kernel(const BinParams p) {
    __shared__ __align__(16) uint32_t s_keys[kEmitWin];
    __shared__ __align__(16) uint32_t s_vals[kEmitWin];
    __shared__ uint32_t s_big[kEmitBig];
    __shared__ uint32_t s_nbig;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t v = min(*p.visible_count, p.max_visible);
    const uint32_t d = *p.dup_count;  // min(duplicates of the frame, capacity): written by the scan
    const uint32_t* __restrict__ sorted = (p.sort_parity && *p.sort_parity) ? p.sorted_indices_alt : p.sorted_indices;
    const uint32_t* __restrict__ offs = p.dup_offsets;
    const uint2* __restrict__ boxes = reinterpret_cast<const uint2*>(p.tboxes_sorted);
    for (uint32_t win = blockIdx.x; (unsigned long long)win * kEmitWin < d; win += gridDim.x) {
        const uint32_t w0 = win * kEmitWin, w1 = min(w0 + (uint32_t)kEmitWin, d);
        if (tid == 0) s_nbig = 0;
        // the window's splats: from the one that covers its first duplicate to the one that covers the next window's first
        // (both recorded by the scan; the last window runs to the end of the visible list)
        const uint32_t r0 = __ldg(p.win_first + win);
        const uint32_t r1 = (unsigned long long)(win + 1) * kEmitWin < d ? __ldg(p.win_first + win + 1) : v - 1u;
        __syncthreads();
        for (uint32_t rb = r0 + tid; rb <= r1; rb += 256 * kEmitUnroll) {
            // kEmitUnroll splats per thread and iteration: their loads are issued together (a window of small splats is ~2600 of them)
            uint32_t off_[kEmitUnroll], g_[kEmitUnroll];
            uint2 q_[kEmitUnroll];
#pragma unroll
            for (int u = 0; u < kEmitUnroll; u++) {
                const uint32_t r = rb + u * 256;
                const bool ok = r <= r1;
                off_[u] = ok ? __ldg(offs + r) : 0xffffffffu;  // (>= w1: skipped below)
                q_[u] = ok ? __ldg(boxes + r) : make_uint2(1u, 0u);
                g_[u] = ok ? __ldg(sorted + r) : 0u;
            }
#pragma unroll
            for (int u = 0; u < kEmitUnroll; u++) {
                const uint32_t off = off_[u], g = g_[u];
                uint32_t x0, y0, w;
                const uint32_t n = box_tiles(q_[u], p.ty_lo, p.ty_hi, x0, y0, w);
                // (the last window runs to the end of the list: with a frame that overflowed the capacity its tail starts beyond w1)
                const uint32_t j0 = off < w0 ? w0 - off : 0u, j1 = off < w1 ? min(n, w1 - off) : 0u;
                if (j0 >= j1) continue;
                if (j1 - j0 > 32u) {
                    s_big[atomicAdd(&s_nbig, 1u)] = rb + u * 256;
                    continue;
                }
                uint32_t row = 0, col = j0;
                if (j0 >= w) {
                    row = j0 / w;
                    col = j0 - row * w;
                }
                uint32_t key = (y0 - p.ty_lo + row) * p.tiles_x + x0 + col;  // strip-relative tile id
                uint32_t o = off + j0 - w0;
                for (uint32_t j = j0; j < j1; j++, o++) {
                    s_keys[o] = key;
                    s_vals[o] = g;
                    ++key;
                    if (++col == w) {
                        col = 0;
                        key += p.tiles_x - w;
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t nbig = s_nbig;
        for (uint32_t b = warp; b < nbig; b += 8) {
            const uint32_t r = s_big[b];
            const uint32_t off = __ldg(offs + r), g = sorted[r];
            uint32_t x0, y0, w;
            const uint32_t n = box_tiles(__ldg(boxes + r), p.ty_lo, p.ty_hi, x0, y0, w);
            const uint32_t j0 = off < w0 ? w0 - off : 0u, j1 = min(n, w1 - off);
            // j / w by multiplication: floor(j / w) == umulhi(j, ceil(2^32 / w)) whenever j * w < 2^32 (w == 1 would need 2^32)
            const bool mul_ok = w > 1u && (unsigned long long)n * w < (1ull << 32);
            const uint32_t inv = mul_ok ? 0xffffffffu / w + 1u : 0u;
            for (uint32_t j = j0 + lane; j < j1; j += 32) {
                const uint32_t qd = mul_ok ? __umulhi(j, inv) : j / w;
                const uint32_t o = off + j - w0;
                s_keys[o] = (y0 - p.ty_lo + qd) * p.tiles_x + (x0 + (j - qd * w));
                s_vals[o] = g;
            }
        }
        __syncthreads();
        const uint32_t cnt = w1 - w0;
        for (uint32_t i = tid * 4; i < cnt; i += 256 * 4) {
            if (i + 4 <= cnt) {
                *reinterpret_cast<uint4*>(p.dup_keys + w0 + i) = *reinterpret_cast<const uint4*>(s_keys + i);
                *reinterpret_cast<uint4*>(p.dup_vals + w0 + i) = *reinterpret_cast<const uint4*>(s_vals + i);
            } else {
                for (uint32_t k = i; k < cnt; k++) {
                    p.dup_keys[w0 + k] = s_keys[k];
                    p.dup_vals[w0 + k] = s_vals[k];
                }
            }
        }
        __syncthreads();  // the staging buffers are reused by the CTA's next window
    }
}

// K5b: tile ranges from the tile-sorted keys + gather of the records into tile order.
__global__ void __launch_bounds__(256)
    gather_kernel(const uint32_t* __restrict__ dup_keys, const uint32_t* __restrict__ dup_vals, const uint32_t* __restrict__ dup_count,
                  const SplatRec* __restrict__ recs, SplatRec* __restrict__ tile_recs, uint32_t* __restrict__ tile_ranges) {
    const uint32_t d = *dup_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
        const uint32_t t = dup_keys[i];
        if (i == 0 || dup_keys[i - 1] != t) tile_ranges[2 * t] = i;
        if (i == d - 1 || dup_keys[i + 1] != t) tile_ranges[2 * t + 1] = i + 1;
        const float4* src = reinterpret_cast<const float4*>(&recs[dup_vals[i]]);
        float4* dst = reinterpret_cast<float4*>(&tile_recs[i]);
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        dst[0] = a;
        dst[1] = b;
        dst[2] = c;
    }
}

// ---------------------------------------------------------------- K6 rasterizer

constexpr int kBatch = 256;  // splat records per shared-memory stage (12 KB)
// gather4 path: 256 records per batch = 64 gather4s (two per producer lane).  The tensor map declares rows 16 floats wide
// over the 48-byte record stride, so gather m brings records 4m..4m+3 as four 64-byte rows (a record + 16 bytes of its
// successor) into one 256-byte slot (the TMA destination must be 128-byte aligned): record j of the batch sits at 64 j.
#ifndef SB_G4_BATCH
#define SB_G4_BATCH 256
#endif
#ifndef SB_G4_STAGES
#define SB_G4_STAGES 2
#endif
constexpr int kBatchG4 = SB_G4_BATCH;  // 128 or 256: one or two gather4 per producer lane
constexpr int kG4Stages = SB_G4_STAGES;
constexpr int kG4PerLane = kBatchG4 / 128;
constexpr int kG4StageF4 = (kBatchG4 / 4) * 16;

enum { FMT_UNORM8 = 0, FMT_F16 = 1, FMT_F32 = 2, FMT_SRGB8 = 3 };  // FMT_SRGB8: same bytes as FMT_UNORM8, the blend works on decoded values

// sRGB attachments: the pixel state is the stored 8-bit code; a blend decodes it, blends in linear and encodes the result.
// Tables and the exact definition of encode(): scripts/gen_srgb_tables.py (the oracle holds its own generated copy).
#include "sb_srgb_tables.h"
__device__ __forceinline__ float srgb_blend(float code, float om, float src_alpha) {
    const float dl = __uint_as_float(__ldg(&SB_SRGB_DECODE_BITS[(int)code]));
    const float o = __fmaf_rn(dl, om, src_alpha);
    // encode(o) = number of thresholds 1..255 that are <= o.  A fast-math guess, then exact comparisons walk to the answer:
    // the result does not depend on how good the guess is.
    const float s = o <= 0.0031308f ? 12.92f * o : 1.055f * __powf(o, 0.41666667f) - 0.055f;
    int g = min(max(__float2int_rn(s * 255.0f), 0), 255);
    while (g < 255 && o >= __uint_as_float(__ldg(&SB_SRGB_THRESHOLD_BITS[g + 1]))) ++g;
    while (g > 0 && o < __uint_as_float(__ldg(&SB_SRGB_THRESHOLD_BITS[g]))) --g;
    return (float)g;
}

// exp(-x) from exactly rounded steps; mirrors so_exp_neg_poly in oracle/splat_oracle.c
__device__ __forceinline__ float exp_neg_poly(float x) {
    const float y = __fmul_rn(x, -1.44269504f);
    const float n = rintf(y);
    const float f = __fsub_rn(y, n);
    float p = 1.54035304e-4f;
    p = __fmaf_rn(p, f, 1.33335581e-3f);
    p = __fmaf_rn(p, f, 9.61812911e-3f);
    p = __fmaf_rn(p, f, 5.55041087e-2f);
    p = __fmaf_rn(p, f, 2.40226507e-1f);
    p = __fmaf_rn(p, f, 6.93147181e-1f);
    p = __fmaf_rn(p, f, 1.0f);
    if (n < -125.0f) return 0.0f;
    return __uint_as_float(__float_as_uint(p) + ((uint32_t)(int)n << 23));
}

// exp(-x) for 0 <= x <= ~80 on the MUFU: ex2.approx.ftz of -x*log2(e).  __expf adds a range
// fix-up for tiny results that cannot occur here (x <= std_dev^2 <= 9).
__device__ __forceinline__ float exp_neg_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * -1.44269504f));
    return y;
}

// round-to-nearest-even of x in [0, 2^22) without the (quarter-rate) FRND instruction
constexpr float kRintMagic = 12582912.0f;  // 1.5 * 2^23
__device__ __forceinline__ float rint_small(float x) { return __fadd_rn(__fadd_rn(x, kRintMagic), -kRintMagic); }

// ---- packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2): two individually rounded IEEE f32 operations
// per issued instruction, so results stay bit-identical to the scalar oracle.  A pair lives in an
// aligned register pair; ptxas folds a pair built from one scalar into the broadcast operand form
// (`R.F32`) and immediates, so no moves are spent on (alpha, alpha) or the rounding constant.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

struct RasterKernelParams {
    const SplatRec* tile_recs;      // bulk path: records gathered into tile order
    const uint32_t* dup_vals;       // gather4 path: Gaussian index per tile-sorted duplicate
    const uint32_t* tile_ranges;
    float cut_k;                    // 1 / kAlphaCut when the staged records carry cut extents (exact alpha cut-off), else 0
    int no_discard;                 // ND instantiation (see eval_splat); only ever set for splat / unorm8 / fast exp / no depth
    const uint32_t* tile_order;     // nullable: CTA i of the launch rasterizes tile tile_order[order_base + i] (heaviest lists first)
    uint32_t order_base;
    uint8_t* pixels;
    uint32_t pitch;
    uint32_t width, height;
    uint32_t row0, rows;
    uint32_t tiles_x;
    uint32_t ty_lo;
    float sd;       // std_dev
    float sd2;      // std_dev^2
    float outline;  // (std_dev - 0.1)^2
    int bgra;
    int clear;
    int obb_cull;   // 1: cull a splat against the warp's patch along the ellipse axes too
    // depth attachment of a caller's render pass (Renderer::render_with_pass + depth_stencil: src/renderer.rs:123,187-195,304)
    float* depth;              // f32 depth, same strip geometry as pixels; nullptr = no depth test
    uint32_t depth_pitch;      // bytes
    int depth_compare;         // SB_COMPARE_*
    int depth_write;
    const uint8_t* pods;       // the fragment depth of a splat is its centre's ndc z (render.wesl:123: clip_pos.zw = proj_pos.zw),
    uint32_t pod_stride;       //   recomputed from the pod position by the producer warp, exactly as K1 computes it
    float pv[16], model[16];
    unsigned long long* counters;  // COUNT builds: [0] alive fragments, [1] evaluated (pixel,splat) lane pairs,
                                   // [2] (warp,splat) evaluations, [3] of those with at least one alive lane
};

// Destination pixel, kept as two packed pairs: (d0, d1) and (d2, d3); 0..255 units on unorm8 targets.
struct PixelState {
    f32x2 d01, d23;
    float depth = 1.0f;  // depth attachment value under this pixel (DEPTH builds)
    uint32_t n_alive = 0, n_eval = 0, n_warp_eval = 0, n_warp_alive = 0;
    __device__ __forceinline__ PixelState() { d01 = pk2(0.0f, 0.0f); d23 = pk2(0.0f, 1.0f); }
};

template <int FMT>
__device__ __forceinline__ void load_dst(PixelState& st, const uint8_t* row, uint32_t x, int bgra) {
    if constexpr (FMT == FMT_UNORM8 || FMT == FMT_SRGB8) {
        const uchar4 c = reinterpret_cast<const uchar4*>(row)[x];
        st.d01 = pk2(bgra ? c.z : c.x, c.y);
        st.d23 = pk2(bgra ? c.x : c.z, 1.0f);
    } else if constexpr (FMT == FMT_F16) {
        const __half* h = reinterpret_cast<const __half*>(row) + 4 * x;
        st.d01 = pk2(__half2float(h[0]), __half2float(h[1]));
        st.d23 = pk2(__half2float(h[2]), __half2float(h[3]));
    } else {
        const float4 c = reinterpret_cast<const float4*>(row)[x];
        st.d01 = pk2(c.x, c.y);
        st.d23 = pk2(c.z, c.w);
    }
}

template <int FMT>
__device__ __forceinline__ void store_dst(const PixelState& st, uint8_t* row, uint32_t x, int bgra) {
    float d0, d1, d2, d3;
    upk2(st.d01, d0, d1);
    upk2(st.d23, d2, d3);
    if constexpr (FMT == FMT_UNORM8 || FMT == FMT_SRGB8) {
        uchar4 c;
        c.x = (unsigned char)(bgra ? d2 : d0);
        c.y = (unsigned char)d1;
        c.z = (unsigned char)(bgra ? d0 : d2);
        c.w = 255;
        reinterpret_cast<uchar4*>(row)[x] = c;
    } else if constexpr (FMT == FMT_F16) {
        __half2 lo = __floats2half2_rn(d0, d1), hi = __floats2half2_rn(d2, d3);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(row)[x] = o;
    } else {
        reinterpret_cast<float4*>(row)[x] = make_float4(d0, d1, d2, d3);
    }
}

// One (pixel, splat) evaluation + blend: splat 31 - hb of the current round.  Branch-free.
__device__ __forceinline__ bool depth_passes(int compare, float z, float d) {
    switch (compare) {
        case SB_COMPARE_NEVER: return false;
        case SB_COMPARE_LESS: return z < d;
        case SB_COMPARE_EQUAL: return z == d;
        case SB_COMPARE_LESS_EQUAL: return z <= d;
        case SB_COMPARE_GREATER: return z > d;
        case SB_COMPARE_NOT_EQUAL: return z != d;
        case SB_COMPARE_GREATER_EQUAL: return z >= d;
        default: return true;
    }
}

struct DepthArgs {
    const float* zs;  // ndc z of the 32 splats of the current round (shared memory), nullptr when DEPTH is off
    int compare, write;
};

// ND ("no discard", splat mode on unorm8 targets with sd^2 >= 6.3 only): a fragment beyond the quad's alive radius has
// alpha = a*exp(-r^2) < exp(-6.3) < kAlphaCut, so blending it is the identity exactly (sb_common.cuh) and the discard test
// and its select can go.
template <int MODE, int FMT, bool STRICT, bool COUNT, bool PERM, bool DEPTH, bool ND = false, bool FOLD = false>
__device__ __forceinline__ void eval_splat(const char* last, uint32_t hb, f32x2 pxy, bool inside, float sd2, float outline, PixelState& st,
                                           const DepthArgs& da) {
    const char* rp = PERM ? last - 64u * hb : last - 48u * hb;
    const float4 q0 = *reinterpret_cast<const float4*>(rp);       // cx cy ax bx
    const float2 q1 = *reinterpret_cast<const float2*>(rp + 16);  // ay by
    // quad offset (render.wesl:125-128 inverted): q = (fma(dx,ax,dy*ay), fma(dx,bx,dy*by))
    float dx, dy, qx, qy;
    upk2(sub2(pxy, pk2(q0.x, q0.y)), dx, dy);
    upk2(fma2(pk2(q0.z, q0.w), pk2(dx, dx), mul2(pk2(q1.x, q1.y), pk2(dy, dy))), qx, qy);
    if constexpr (COUNT) st.n_eval += inside ? 1u : 0u;
    float alpha;
    bool alive;
    const float4 q2 = *reinterpret_cast<const float4*>(rp + 32);  // r g b a
    if constexpr (MODE == SB_MODE_POINT) {  // render.wesl:164-166
        alive = fabsf(qx) <= 1.0f && fabsf(qy) <= 1.0f;
        alpha = 1.0f;
    } else {
        const float r2 = __fmaf_rn(qx, qx, __fmul_rn(qy, qy));
        alive = r2 <= sd2;  // discard otherwise: render.wesl:145,155
        if constexpr (MODE == SB_MODE_SPLAT) {
            float e;
            if constexpr (FOLD) {
                // the staged axes of this build carry a factor sqrt(log2 e) (the CTA's cull folds it in, once per record and
                // batch): r2 is already -log2 of the falloff, one multiply per evaluation less (and sd2 arrives scaled alike)
                static_assert(!STRICT && MODE == SB_MODE_SPLAT, "the fold belongs to the fast-exp splat builds");
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-r2));
            } else {
                e = STRICT ? exp_neg_poly(r2) : exp_neg_fast(r2);
            }
            alpha = __fmul_rn(q2.w, e);  // render.wesl:149
        } else {
            const float ol = r2 > outline ? 1.0f : 0.0f;  // render.wesl:159-160
            alpha = __fadd_rn(q2.w, __fmul_rn(__fsub_rn(1.0f, q2.w), ol));
        }
    }
    if constexpr (COUNT) {
        st.n_alive += (inside && alive) ? 1u : 0u;
        const uint32_t act = __activemask();  // one or both halves
        const bool any = __any_sync(act, alive);
        st.n_warp_eval += 1u;
        st.n_warp_alive += any ? 1u : 0u;
    }
    if constexpr (DEPTH) {
        // depth test AFTER the shader's discard, against the attachment as earlier fragments of this frame left it
        const float z = da.zs[31u - hb];
        alive = alive && depth_passes(da.compare, z, st.depth);
        if (da.write && alive) st.depth = z;
    }
    // A discarded fragment blends with alpha = 0, which is the identity EXACTLY (d*1 + c*0 = d, and d is
    // already an integer on unorm8 targets): the loop body stays branch-free.
    if constexpr (!ND) alpha = alive ? alpha : 0.0f;
    const float om = __fsub_rn(1.0f, alpha);
    const f32x2 om2 = pk2(om, om), al2 = pk2(alpha, alpha);
    if constexpr (FMT == FMT_UNORM8) {
        // d <- rint(fma(d, 1-alpha, c255*alpha)); the source colour is clamped to 255 per splat (K1) and
        // alpha <= 1, so the blend cannot exceed 255 + rounding and the post-blend clamp never binds
        const f32x2 m = pk2(kRintMagic, kRintMagic), nm = pk2(-kRintMagic, -kRintMagic);
        st.d01 = add2(add2(fma2(st.d01, om2, mul2(pk2(q2.x, q2.y), al2)), m), nm);
        float d2, d3;
        upk2(st.d23, d2, d3);
        d2 = rint_small(__fmaf_rn(d2, om, __fmul_rn(q2.z, alpha)));
        st.d23 = pk2(d2, d3);
    } else if constexpr (FMT == FMT_SRGB8) {
        float d0, d1, d2, d3;
        upk2(st.d01, d0, d1);
        upk2(st.d23, d2, d3);
        d0 = srgb_blend(d0, om, __fmul_rn(q2.x, alpha));  // the source colour was clamped to [0,1] per splat (K1: color_max)
        d1 = srgb_blend(d1, om, __fmul_rn(q2.y, alpha));
        d2 = srgb_blend(d2, om, __fmul_rn(q2.z, alpha));
        st.d01 = pk2(d0, d1);
        st.d23 = pk2(d2, d3);
    } else {
        // (b*alpha, 1*alpha): 1*alpha is exact, so d3 = fma(d3, 1-alpha, alpha) as in the oracle
        st.d01 = fma2(st.d01, om2, mul2(pk2(q2.x, q2.y), al2));
        st.d23 = fma2(st.d23, om2, mul2(pk2(q2.z, 1.0f), al2));
        if constexpr (FMT == FMT_F16) {
            float d0, d1, d2, d3;
            upk2(st.d01, d0, d1);
            upk2(st.d23, d2, d3);
            st.d01 = pk2(__half2float(__float2half_rn(d0)), __half2float(__float2half_rn(d1)));
            st.d23 = pk2(__half2float(__float2half_rn(d2)), __half2float(__float2half_rn(d3)));
        }
    }
}

// Composites one staged batch onto this thread's pixel, in list order.  PERM: record j of the batch
// sits at float4 index 4 j (the layout the gather4 producer writes: rows fetched 64 bytes wide).
// Warp-level culling: each lane tests one splat against the two 4x4 halves of the warp's 8x4 pixel patch — the
// bbox of its alive region and (obb) the two ellipse axes, i.e. the full separating-axis test of a half's
// rectangle against the ellipse's oriented bounding box; each half-warp then evaluates only its own survivors.
// CTA-level cull of the gather4 path: ONE thread tests a splat against all sixteen 4x4 blocks of the tile (bit 4 j + i <->
// block column i, row j; bx0/by0 = centre of block (0, 0)) with the same tests a warp used to run for its own two blocks —
// bbox of the alive region and the separating-axis test against the ellipse's oriented bounding box.  The eight warps then
// pick their two bits from the mask instead of each loading and testing every splat of the list themselves.
__device__ __forceinline__ uint32_t block_mask(const float4* __restrict__ rec, float bx0, float by0, float sd, bool obb, float cut_k) {
    const float4 c0 = rec[0];  // cx cy ax bx
    const float4 c1 = rec[1];  // ay by ex ey
    const float ex = c1.z + 1.51f, ey = c1.w + 1.51f;
    float dl[4], dy[4];
    uint32_t cm = 0, rm = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        dl[i] = (bx0 + 4.0f * (float)i) - c0.x;
        dy[i] = (by0 + 4.0f * (float)i) - c0.y;
        cm |= (fabsf(dl[i]) <= ex ? 1u : 0u) << i;
        rm |= (fabsf(dy[i]) <= ey ? 1u : 0u) << i;
    }
    if (cm == 0u || rm == 0u) return 0u;
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) m |= ((rm >> j) & 1u) ? (cm << (4 * j)) : 0u;
    if (obb) {
        // Radius of the region that can change a pixel, in quad-offset units: sd, or with the exact alpha cut-off
        // (sb_common.cuh) the radius beyond which alpha = a*exp(-r^2) < kAlphaCut — a little more margin than the vertex
        // stage takes, so this never cuts tighter than the extents it wrote.
        float rc = sd;
        if (cut_k > 0.0f) rc = sqrtf(fminf(fmaxf(__logf(rec[2].w * cut_k) + 1.5f * kAlphaCutMargin, 0.0f), sd * sd));
        // separating axes of the ellipse's frame: |q(p)| <= |q(pc)| + 1.5|a_x| + 1.5|a_y| over a block's pixel centres
        const float mx = fmaf(1.5f, fabsf(c0.z) + fabsf(c1.x), rc), my = fmaf(1.5f, fabsf(c0.w) + fabsf(c1.y), rc);
        const float lx = fmaf(mx, 1.0001f, 1.0e-3f), ly = fmaf(my, 1.0001f, 1.0e-3f);
        // disc: |q(p) - q(pc)| <= R = 1.5 max |(ax, bx) +- (ay, by)| over the block, so a block with |q(pc)| > rc + R has
        // no pixel inside the radius (cuts the corners the two axis tests leave)
        const float sx = c0.z + c1.x, sy = c0.w + c1.y, tx_ = c0.z - c1.x, ty_ = c0.w - c1.y;
        const float rr = fmaf(1.5f * 1.0001f, sqrtf(fmaxf(fmaf(sx, sx, sy * sy), fmaf(tx_, tx_, ty_ * ty_))), rc) + 1.0e-3f;
        const float lim2 = rr * rr;
        uint32_t keep = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tx = dy[j] * c1.x, ty = dy[j] * c1.y;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float qx = fmaf(dl[i], c0.z, tx), qy = fmaf(dl[i], c0.w, ty);
                const bool h = fabsf(qx) <= lx && fabsf(qy) <= ly && fmaf(qx, qx, qy * qy) <= lim2;
                keep |= (h ? 1u : 0u) << (4 * j + i);
            }
        }
        m &= keep;
    }
    return m;
}

// SHARED: the hit masks of the batch were computed once per CTA (block_mask) and sit in shared memory; `bl` = this warp's
// left block's bit.
template <int MODE, int FMT, bool STRICT, bool COUNT, bool PERM, bool DEPTH = false, bool ND = false, bool SHARED = false, bool FOLD = false>
__device__ __forceinline__ void composite_batch(const float4* __restrict__ recs, uint32_t cnt, f32x2 pxy, float pcx, float pcy,
                                                uint32_t lane, bool inside, float sd, float sd2, float outline, bool obb,
                                                PixelState& st, DepthArgs da = DepthArgs{nullptr, 0, 0},
                                                const uint16_t* __restrict__ masks = nullptr, uint32_t bl = 0) {
    // Byte addressing.  PERM (gather4 path): record j sits at 64 j (rows are fetched 64 bytes wide, four per 256-byte gather);
    // otherwise at 48 j.  A round is 32 consecutive records starting at a multiple of 32.
    const char* rb = reinterpret_cast<const char*>(recs);
    const char* mine = rb + (PERM ? 64u * lane : 48u * lane);
    constexpr uint32_t kRound = PERM ? 2048u : 1536u;
    for (uint32_t base = 0; base < cnt; base += 32, mine += kRound, rb += kRound, da.zs += DEPTH ? 32 : 0) {
        // each lane tests ONE splat against BOTH 4x4 halves of the warp's 8x4 patch (pcx: centre of the left half)
        bool hit_l = false, hit_r = false;
        if constexpr (SHARED) {
            const uint32_t m = masks[base + lane] >> bl;  // masks past the end of the list are zero
            hit_l = m & 1u;
            hit_r = m & 2u;
        } else if (base + lane < cnt) {
            const float4 c0 = *reinterpret_cast<const float4*>(mine);       // cx cy ax bx
            const float4 c1 = *reinterpret_cast<const float4*>(mine + 16);  // ay by ex ey
            const float dl = pcx - c0.x, dr = dl + 4.0f, ddy = pcy - c0.y;
            const bool hy = fabsf(ddy) <= c1.w + 1.51f;
            hit_l = hy && fabsf(dl) <= c1.z + 1.51f;
            hit_r = hy && fabsf(dr) <= c1.z + 1.51f;
            if (obb) {
                // |q(p)| <= |q(pc)| + 1.5|a_x| + 1.5|a_y| over a half's pixel centres; alive needs |q| <= sd
                const float tx = ddy * c1.x, ty = ddy * c1.y;
                const float mx = fmaf(1.5f, fabsf(c0.z) + fabsf(c1.x), sd), my = fmaf(1.5f, fabsf(c0.w) + fabsf(c1.y), sd);
                const float lx = fmaf(mx, 1.0001f, 1.0e-3f), ly = fmaf(my, 1.0001f, 1.0e-3f);
                hit_l = hit_l && fabsf(fmaf(dl, c0.z, tx)) <= lx && fabsf(fmaf(dl, c0.w, ty)) <= ly;
                hit_r = hit_r && fabsf(fmaf(dr, c0.z, tx)) <= lx && fabsf(fmaf(dr, c0.w, ty)) <= ly;
            }
        }
        // Two independent splat streams per warp: lanes 0-15 (left half) walk the splats that touch the left 4x4 pixels,
        // lanes 16-31 those of the right half; a splat of a few pixels usually touches only one of them, so the loop runs
        // max(n_left, n_right) times instead of n_left-or-right.  Bit hb = 31 - b <-> splat b of the round.
        const uint32_t ml = __ballot_sync(0xffffffffu, hit_l), mr = __ballot_sync(0xffffffffu, hit_r);
        // address of splat b = 31 - hb: rb + 64 b = last - 64 hb   (PERM)
        //                                 rb + 48 b                = last - 48 hb               (else)
        const char* last = rb + (PERM ? 64u * 31u : 48u * 31u);
        // Big splats touch both halves: then one warp-uniform walk over the union costs the same number of iterations
        // and keeps the loop control on the uniform datapath.  Split only when it saves at least two iterations.
        const uint32_t un = ml | mr;
        if (__popc(un) >= max(__popc(ml), __popc(mr)) + 2) {
            uint32_t todo = __brev(lane < 16 ? ml : mr);
            while (todo) {
                uint32_t hb;  // FLO directly; `31 - __clz` is canonicalised back into a clz and costs five more integer ops
                asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(todo));
                todo ^= 1u << hb;
                eval_splat<MODE, FMT, STRICT, COUNT, PERM, DEPTH, ND, FOLD>(last, hb, pxy, inside, sd2, outline, st, da);
            }
        } else {
            uint32_t todo = __brev(un);
            while (todo) {
                uint32_t hb;
                asm("bfind.u32 %0, %1;" : "=r"(hb) : "r"(todo));
                todo ^= 1u << hb;
                eval_splat<MODE, FMT, STRICT, COUNT, PERM, DEPTH, ND, FOLD>(last, hb, pxy, inside, sd2, outline, st, da);
            }
        }
    }
}

template <bool COUNT>
__device__ __forceinline__ void flush_counters(PixelState& st, uint32_t lane, unsigned long long* counters) {
    if constexpr (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st.n_alive += __shfl_xor_sync(0xffffffffu, st.n_alive, o);
            st.n_eval += __shfl_xor_sync(0xffffffffu, st.n_eval, o);
        }
        if (lane == 0 && counters) {
            atomicAdd(&counters[0], (unsigned long long)st.n_alive);
            atomicAdd(&counters[1], (unsigned long long)st.n_eval);
            atomicAdd(&counters[2], (unsigned long long)st.n_warp_eval);
            atomicAdd(&counters[3], (unsigned long long)st.n_warp_alive);
        }
    }
}

// K6 (a): batches are contiguous in tile_recs (written by gather_kernel); one elected thread
// streams them in with 1-D TMA bulk copies, double buffered.
template <int MODE, int FMT, bool STRICT, bool COUNT>
__global__ void __launch_bounds__(256) raster_bulk_kernel(const RasterKernelParams p) {
    __shared__ __align__(128) float4 stage[2][kBatch * 3];
    __shared__ __align__(8) uint64_t full_bar[2];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    uint32_t tile_x = blockIdx.x, tile_y = p.ty_lo + blockIdx.y;
    if (p.tile_order) {  // longest-list-first schedule: the kernel's tail is made of short lists
        const uint32_t t = p.tile_order[p.order_base + blockIdx.y * gridDim.x + blockIdx.x];
        tile_x = t % p.tiles_x;
        tile_y = t / p.tiles_x;
    }
    const uint32_t tile = tile_y * p.tiles_x + tile_x;
    // each warp owns an 8x4 pixel patch: lanes 0-15 its left 4x4 half, lanes 16-31 the right one (pixel centres:
    // half extents 1.5 x 1.5 around the half's centre; pcx = centre of the LEFT half)
    const uint32_t x = tile_x * kTile + (warp & 1u) * 8 + (lane >> 4) * 4 + (lane & 3u);
    const uint32_t y = tile_y * kTile + (warp >> 1) * 4 + ((lane >> 2) & 3u);
    const bool inside = x < p.width && y >= p.row0 && y < p.row0 + p.rows;
    const float px = (float)x + 0.5f, py = (float)y + 0.5f;
    const float pcx = (float)(tile_x * kTile + (warp & 1u) * 8) + 2.0f, pcy = (float)(tile_y * kTile + (warp >> 1) * 4) + 2.0f;

    const uint32_t begin = p.tile_ranges[2 * tile], end = p.tile_ranges[2 * tile + 1];
    const uint32_t total = end - begin;
    const uint32_t batches = (total + kBatch - 1) / kBatch;

    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0 && batches > 0) {
        const uint32_t bytes = min((uint32_t)kBatch, total) * (uint32_t)sizeof(SplatRec);
        mbar_arrive_expect_tx(&full_bar[0], bytes);
        bulk_g2s(stage[0], p.tile_recs + begin, bytes, &full_bar[0]);
    }

    PixelState st;
    uint8_t* dst = p.pixels + (size_t)(y - p.row0) * p.pitch;
    if (!p.clear && inside) load_dst<FMT>(st, dst, x, p.bgra);

    for (uint32_t k = 0; k < batches; k++) {
        const uint32_t s = k & 1u;
        if (tid == 0 && k + 1 < batches) {
            const uint32_t nb = min((uint32_t)kBatch, total - (k + 1) * kBatch) * (uint32_t)sizeof(SplatRec);
            mbar_arrive_expect_tx(&full_bar[s ^ 1u], nb);
            bulk_g2s(stage[s ^ 1u], p.tile_recs + begin + (size_t)(k + 1) * kBatch, nb, &full_bar[s ^ 1u]);
        }
        mbar_wait(&full_bar[s], (k >> 1) & 1u);
        const uint32_t cnt = min((uint32_t)kBatch, total - k * kBatch);
        composite_batch<MODE, FMT, STRICT, COUNT, false>(stage[s], cnt, pk2(px, py), pcx, pcy, lane, inside, p.sd, p.sd2, p.outline, p.obb_cull != 0, st);
        __syncthreads();  // everyone is done with stage[s] before it is refilled
    }
    flush_counters<COUNT>(st, lane, p.counters);
    if (inside) store_dst<FMT>(st, dst, x, p.bgra);
}

// K6 (b): no gathered copy of the records at all.  A producer warp reads the tile's Gaussian indices
// (coalesced) and fetches the 48-byte records straight from the per-Gaussian array with TMA
// tile::gather4 (cp.async.bulk.tensor.2d ... tile::gather4 -> UTMALDG, four rows per instruction, two
// instructions per lane and batch) into a double-buffered ring; eight consumer warps composite.
template <int MODE, int FMT, bool STRICT, bool COUNT, bool DEPTH = false, bool ND = false, bool FOLD = false>
__global__ void __launch_bounds__(288) raster_gather4_kernel(const __grid_constant__ RasterKernelParams p, const __grid_constant__ CUtensorMap recs_map) {
    __shared__ __align__(256) float4 stage[kG4Stages][kG4StageF4];
    __shared__ float zs[DEPTH ? kG4Stages : 1][DEPTH ? kBatchG4 : 1];  // ndc z per staged record (depth-tested passes only)
    __shared__ __align__(8) uint64_t z_bar[kG4Stages];  // DEPTH: producer's generic stores to zs -> consumers (plain arrive/wait pair)
    __shared__ __align__(8) uint64_t full_bar[kG4Stages];
    __shared__ __align__(8) uint64_t empty_bar[kG4Stages];
    __shared__ uint16_t cull_mask[kG4Stages][kBatchG4];  // block_mask of every staged record (consumer thread t owns record t)
    static_assert(kBatchG4 == 256, "one consumer thread per staged record");

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    uint32_t tile_x = blockIdx.x, tile_y = p.ty_lo + blockIdx.y;
    if (p.tile_order) {  // longest-list-first schedule: the kernel's tail is made of short lists
        const uint32_t t = p.tile_order[p.order_base + blockIdx.y * gridDim.x + blockIdx.x];
        tile_x = t % p.tiles_x;
        tile_y = t / p.tiles_x;
    }
    const uint32_t tile = tile_y * p.tiles_x + tile_x;
    const uint32_t begin = p.tile_ranges[2 * tile], end = p.tile_ranges[2 * tile + 1];
    const uint32_t total = end - begin;
    const uint32_t batches = (total + kBatchG4 - 1) / kBatchG4;

    if (tid == 0) {
        for (int s = 0; s < kG4Stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 8);
            mbar_init(&z_bar[s], 32);  // every producer lane arrives for its own stores
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == 8) {
        // ---------------- producer warp: lane t issues gather4 #t (records 4t..4t+3) and #t+32.  The Gaussian
        // indices of batch k+1 are fetched right after the gathers of batch k were issued, so a refill costs one
        // gather round trip, not an index load followed by one.
        uint32_t g[4 * kG4PerLane];
        auto load_indices = [&](uint32_t k) {
            const uint32_t cnt = min((uint32_t)kBatchG4, total - k * kBatchG4);
            const uint32_t* idx = p.dup_vals + begin + (size_t)k * kBatchG4;
            const uint32_t g_first = __ldg(idx);  // padding index for rows past the end of the list
#pragma unroll
            for (int h = 0; h < kG4PerLane; h++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t j = 4u * (lane + 32u * h) + i;
                    g[h * 4 + i] = j < cnt ? __ldg(idx + j) : g_first;
                }
        };
        if (batches > 0) load_indices(0);
        for (uint32_t k = 0; k < batches; k++) {
            const uint32_t s = k % kG4Stages;
            mbar_wait(&empty_bar[s], ((k / kG4Stages) & 1u) ^ 1u);
            const uint32_t cnt = min((uint32_t)kBatchG4, total - k * kBatchG4);
            // only gathers whose first record exists are issued; the barrier is armed with exactly the
            // bytes that will land
            const uint32_t gathers = (cnt + 3u) / 4u;
            if constexpr (DEPTH) {
                // fragment depth = the centre's ndc z, recomputed from the pod position with K1's strict arithmetic
#pragma unroll
                for (int h = 0; h < kG4PerLane; h++)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const float4 head = __ldg(reinterpret_cast<const float4*>(p.pods + (size_t)g[4 * h + i] * p.pod_stride));
                        float world[3], cz, cw;
#pragma unroll
                        for (int c = 0; c < 3; c++)
                            world[c] = sadd(sadd(sadd(smul(p.model[c], head.x), smul(p.model[4 + c], head.y)), smul(p.model[8 + c], head.z)),
                                            p.model[12 + c]);
                        cz = sadd(sadd(sadd(smul(p.pv[2], world[0]), smul(p.pv[6], world[1])), smul(p.pv[10], world[2])), p.pv[14]);
                        cw = sadd(sadd(sadd(smul(p.pv[3], world[0]), smul(p.pv[7], world[1])), smul(p.pv[11], world[2])), p.pv[15]);
                        zs[s][4u * (lane + 32u * h) + i] = sdiv(cz, cw);
                    }
                mbar_arrive(&z_bar[s]);
            }
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], gathers * 4u * 64u);  // rows are fetched 64 bytes wide
            __syncwarp();
            const uint32_t dst0 = smem_u32(&stage[s][lane * 16]);
#pragma unroll
            for (int h = 0; h < kG4PerLane; h++)
                if (lane + 32u * h < gathers)
                    tma_gather4(dst0 + 32u * 256u * h, &recs_map, &full_bar[s], g[4 * h], g[4 * h + 1], g[4 * h + 2], g[4 * h + 3]);
            if (k + 1 < batches) load_indices(k + 1);
        }
        return;
    }

    // ---------------- consumers: each warp owns an 8x4 pixel patch, lanes 0-15 its left 4x4 half, lanes 16-31 the right one
    const uint32_t x = tile_x * kTile + (warp & 1u) * 8 + (lane >> 4) * 4 + (lane & 3u);
    const uint32_t y = tile_y * kTile + (warp >> 1) * 4 + ((lane >> 2) & 3u);
    const bool inside = x < p.width && y >= p.row0 && y < p.row0 + p.rows;
    const float px = (float)x + 0.5f, py = (float)y + 0.5f;
    const float pcx = (float)(tile_x * kTile + (warp & 1u) * 8) + 2.0f, pcy = (float)(tile_y * kTile + (warp >> 1) * 4) + 2.0f;
    const float bx0 = (float)(tile_x * kTile) + 2.0f, by0 = (float)(tile_y * kTile) + 2.0f;  // centre of block (0, 0)
    const uint32_t bl = (warp >> 1) * 4u + (warp & 1u) * 2u;                                   // bit of this warp's left block
    PixelState st;
    uint8_t* dst = p.pixels + (size_t)(y - p.row0) * p.pitch;
    if (!p.clear && inside) load_dst<FMT>(st, dst, x, p.bgra);
    float* zdst = nullptr;
    if constexpr (DEPTH) {
        zdst = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.depth) + (size_t)(y - p.row0) * p.depth_pitch) + x;
        if (inside) st.depth = *zdst;
    }

    for (uint32_t k = 0; k < batches; k++) {
        const uint32_t s = k % kG4Stages;
        mbar_wait(&full_bar[s], (k / kG4Stages) & 1u);
        if constexpr (DEPTH) mbar_wait(&z_bar[s], (k / kG4Stages) & 1u);
        const uint32_t cnt = min((uint32_t)kBatchG4, total - k * kBatchG4);
        // CTA-level cull: thread t tests record t of the batch against the tile's sixteen blocks, once for all warps.  The
        // buffer of stage s is free: the stage was refilled only after every warp had released it (empty_bar).
        cull_mask[s][tid] = tid < cnt ? (uint16_t)block_mask(&stage[s][4u * tid], bx0, by0, p.sd, p.obb_cull != 0, p.cut_k) : (uint16_t)0;
        if constexpr (FOLD) {
            // fast-exp splat builds on unorm8 with sd^2 >= 6.3 (where a fragment at the discard radius no longer changes a pixel,
            // so the rounding of r^2 at that boundary cannot show): fold sqrt(log2 e) into the staged record's inverse axes, so that
            // the quad offset's squared length is the exponent of ex2 directly.  One thread per record and batch; the mask above
            // was computed from the unscaled axes, and the barrier below publishes both.
            if (tid < cnt) {
                constexpr float kS = 1.2011224087864498f;  // sqrt(log2 e)
                float4* r = &stage[s][4u * tid];
                const float2 a0 = *reinterpret_cast<const float2*>(&r[0].z);  // ax bx
                const float2 a1 = *reinterpret_cast<const float2*>(&r[1].x);  // ay by
                *reinterpret_cast<float2*>(&r[0].z) = make_float2(a0.x * kS, a0.y * kS);
                *reinterpret_cast<float2*>(&r[1].x) = make_float2(a1.x * kS, a1.y * kS);
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight consumer warps (the producer warp is not part of it)
        composite_batch<MODE, FMT, STRICT, COUNT, true, DEPTH, ND, true, FOLD>(stage[s], cnt, pk2(px, py), pcx, pcy, lane, inside, p.sd,
                                                                            FOLD ? p.sd2 * 1.44269504f : p.sd2,
                                                                     p.outline, p.obb_cull != 0, st,
                                                                     DepthArgs{DEPTH ? zs[s] : nullptr, p.depth_compare, p.depth_write},
                                                                     cull_mask[s], bl);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    flush_counters<COUNT>(st, lane, p.counters);
    if (inside) store_dst<FMT>(st, dst, x, p.bgra);
    if constexpr (DEPTH)
        if (inside && p.depth_write) *zdst = st.depth;
}

// tile ranges from the tile-sorted keys (gather4 path: no record copy): four keys per thread, one 128-bit load
__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t* __restrict__ dup_keys, const uint32_t* __restrict__ dup_count,
                                                          uint32_t* __restrict__ tile_ranges) {
    const uint32_t d = *dup_count;
    const uint32_t nvec = (d + 3) / 4;  // the buffer is padded to a multiple of four keys
    const uint4* kv = reinterpret_cast<const uint4*>(dup_keys);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        const uint4 q = kv[i];
        const uint32_t base = 4 * i;
        const uint32_t k[6] = {base > 0 ? dup_keys[base - 1] : 0xffffffffu, q.x, q.y, q.z, q.w,
                               base + 4 < d ? dup_keys[base + 4] : 0xffffffffu};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t idx = base + j;
            if (idx < d) {
                const uint32_t t = k[j + 1];
                const uint32_t next = idx + 1 < d ? k[j + 2] : 0xffffffffu;
                if (k[j] != t) tile_ranges[2 * t] = idx;
                if (next != t) tile_ranges[2 * t + 1] = idx + 1;
            }
        }
    }
}

// Longest-processing-time-first schedule for the rasterizer: tiles of the strip ordered by descending list length
// (counting sort on length / 32, one block).  A tile's CTA runs as long as its list, the heaviest 0.2-0.5 ms: started
// in row-major order the late heavy tiles leave most of the machine idle at the end of the kernel.
__global__ void __launch_bounds__(1024) tile_order_kernel(const uint32_t* __restrict__ tile_ranges, uint32_t first_tile, uint32_t nt,
                                                          uint32_t* __restrict__ order) {
    __shared__ uint32_t hist[1024];
    __shared__ uint32_t wsum[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint2* rg = reinterpret_cast<const uint2*>(tile_ranges) + first_tile;
    constexpr int kGroup = 8;  // tiles per thread whose ranges are fetched together (one round trip at 1080p)
    auto bucket = [&](uint32_t i) { return i < nt ? 1023u - min((rg[i].y - rg[i].x) >> 5, 1023u) : 0xffffffffu; };
    hist[tid] = 0;
    __syncthreads();
    // count, warp-aggregated: the empty tiles of a frame (most of them) all fall into one bucket
    for (uint32_t g0 = 0; g0 < nt; g0 += kGroup * 1024u) {
        uint32_t bk[kGroup];
#pragma unroll
        for (int k = 0; k < kGroup; k++) bk[k] = bucket(g0 + k * 1024u + tid);
#pragma unroll
        for (int k = 0; k < kGroup; k++) {
            const uint32_t peers = __match_any_sync(0xffffffffu, bk[k]);  // inactive lanes (0xffffffff) form their own group
            if (bk[k] != 0xffffffffu && (peers & lt) == 0) atomicAdd(&hist[bk[k]], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    const uint32_t c = hist[tid];
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = inc - c;
    for (uint32_t w = 0; w < warp; w++) base += wsum[w];
    __syncthreads();
    hist[tid] = base;
    __syncthreads();
    for (uint32_t g0 = 0; g0 < nt; g0 += kGroup * 1024u) {
        uint32_t bk[kGroup];
#pragma unroll
        for (int k = 0; k < kGroup; k++) bk[k] = bucket(g0 + k * 1024u + tid);
#pragma unroll
        for (int k = 0; k < kGroup; k++) {
            const uint32_t peers = __match_any_sync(0xffffffffu, bk[k]);
            uint32_t pos = 0;
            if (bk[k] != 0xffffffffu && (peers & lt) == 0) pos = atomicAdd(&hist[bk[k]], (uint32_t)__popc(peers));
            pos = __shfl_sync(0xffffffffu, pos, __ffs(peers) - 1);
            if (bk[k] != 0xffffffffu) order[pos + __popc(peers & lt)] = first_tile + g0 + k * 1024u + tid;
        }
    }
}

// A render pass that only clears (Color::BLACK = 0,0,0,1): src/renderer.rs:171-177
__global__ void clear_kernel(uint8_t* pixels, uint32_t pitch, uint32_t width, uint32_t rows, int fmt) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= width || y >= rows) return;
    uint8_t* row = pixels + (size_t)y * pitch;
    if (fmt == FMT_UNORM8) {
        reinterpret_cast<uchar4*>(row)[x] = make_uchar4(0, 0, 0, 255);
    } else if (fmt == FMT_F16) {
        reinterpret_cast<uint2*>(row)[x] = make_uint2(0u, 0x3c000000u);
    } else {
        reinterpret_cast<float4*>(row)[x] = make_float4(0.f, 0.f, 0.f, 1.f);
    }
}

template <int MODE, int FMT, bool STRICT>
void launch_raster(const RasterKernelParams& kp, const CUtensorMap* recs_map, dim3 grid, cudaStream_t stream) {
    if (recs_map && kp.depth) {
        raster_gather4_kernel<MODE, FMT, STRICT, false, true><<<grid, 288, 0, stream>>>(kp, *recs_map);
    } else if (recs_map) {
        if (kp.counters) {
            raster_gather4_kernel<MODE, FMT, STRICT, true><<<grid, 288, 0, stream>>>(kp, *recs_map);
        } else if (kp.no_discard) {
            if constexpr (MODE == SB_MODE_SPLAT && FMT == FMT_UNORM8 && !STRICT)
                raster_gather4_kernel<MODE, FMT, STRICT, false, false, true, true><<<grid, 288, 0, stream>>>(kp, *recs_map);
        } else {
            // the same arithmetic with the exact cut-off switched off (the discard test kept): frames stay bit-identical between the two
            bool folded = false;
            if constexpr (MODE == SB_MODE_SPLAT && FMT == FMT_UNORM8 && !STRICT) {
                if (kp.sd2 >= 6.3f) {
                    raster_gather4_kernel<MODE, FMT, STRICT, false, false, false, true><<<grid, 288, 0, stream>>>(kp, *recs_map);
                    folded = true;
                }
            }
            if (!folded) raster_gather4_kernel<MODE, FMT, STRICT, false><<<grid, 288, 0, stream>>>(kp, *recs_map);
        }
    } else {
        if (kp.counters) raster_bulk_kernel<MODE, FMT, STRICT, true><<<grid, 256, 0, stream>>>(kp);
        else raster_bulk_kernel<MODE, FMT, STRICT, false><<<grid, 256, 0, stream>>>(kp);
    }
}

}  // namespace

cudaError_t launch_clear(const SbTarget& t, cudaStream_t stream) {
    const uint32_t rows = t.rows == 0 ? t.height : t.rows;
    if (rows == 0 || t.width == 0) return cudaSuccess;
    const bool srgb = t.format == SB_TARGET_RGBA8_UNORM_SRGB || t.format == SB_TARGET_BGRA8_UNORM_SRGB;
    const int fmt = (t.format == SB_TARGET_RGBA8_UNORM || t.format == SB_TARGET_BGRA8_UNORM || srgb) ? FMT_UNORM8  // BLACK is code 0 either way
                    : t.format == SB_TARGET_RGBA16_FLOAT                                               ? FMT_F16
                                                                                                       : FMT_F32;
    clear_kernel<<<dim3((t.width + 255) / 256, rows), 256, 0, stream>>>(reinterpret_cast<uint8_t*>(t.d_pixels), t.pitch_bytes, t.width,
                                                                      rows, fmt);
    return cudaGetLastError();
}

cudaError_t launch_bin_and_raster(const RasterParams& p, int num_sms, cudaStream_t stream) {
    const Uniforms& u = p.u;
    const SbTarget& t = p.target;
    const uint32_t rows = t.rows == 0 ? u.height : t.rows;
    const uint32_t row0 = t.rows == 0 ? 0 : t.row0;
    if (rows == 0 || u.width == 0) return cudaSuccess;
    const uint32_t ty_lo = row0 / kTile, ty_hi = (row0 + rows - 1) / kTile;
    const uint32_t num_tiles = u.tiles_x * u.tiles_y;

    cudaError_t e;
    // SB_BIN=v1 keeps the round-1 scan / emit pair (A/B runs); it is also the path when there is no buffer for the sorted boxes
    static const bool bin_v1_env = [] { const char* c = std::getenv("SB_BIN"); return c && std::string(c) == "v1"; }();
    const bool bin_v1 = bin_v1_env || p.buf.tboxes_sorted == nullptr || p.buf.win_first == nullptr;
    // tile ranges zeroed (empty tiles -> begin=end=0); v1 scan state: [counter(16B)] [status...] (the v2 scan writes every word it reads)
    if (bin_v1) {
        const size_t scan_tiles = ((size_t)p.max_visible + kScanTile - 1) / kScanTile + 1;
        e = cudaMemsetAsync(p.buf.scan_counter, 0, 16 + scan_tiles * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
    }
    e = cudaMemsetAsync(p.buf.tile_ranges, 0, (size_t)num_tiles * 2 * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    // tile ids are relative to the strip's first tile row: a strip of a huge frame sorts fewer key bits (an 8K frame has 17,
    // an eighth of it 14: two 8-bit passes instead of three); the ranges land in the frame-wide table through an offset pointer
    const uint32_t strip_tiles = (ty_hi - ty_lo + 1) * u.tiles_x;
    uint32_t* const ranges_rel = p.buf.tile_ranges + 2 * (size_t)ty_lo * u.tiles_x;
    int bits = 1;
    while ((1u << bits) < strip_tiles) bits++;

    BinParams bp;
    bp.tboxes = p.tboxes;
    bp.sorted_indices = p.sorted_indices;
    bp.sorted_indices_alt = p.sorted_indices_alt;
    bp.sort_parity = p.sort_parity;
    bp.visible_count = p.visible_count;
    bp.dup_offsets = p.buf.dup_offsets;
    bp.tboxes_sorted = p.buf.tboxes_sorted;
    bp.tiles_prefixed = 0;
    bp.win_first = p.buf.win_first;
    bp.win_capacity = p.buf.win_capacity;
    bp.dup_keys = p.buf.dup_keys;
    bp.dup_vals = p.buf.dup_vals;
    bp.dup_count = p.buf.dup_count;
    bp.overflow = p.buf.overflow;
    bp.needed_host = p.buf.needed_host;
    bp.scan_status = p.buf.scan_status;
    bp.scan_counter = p.buf.scan_counter;
    bp.dup_capacity = (uint32_t)p.buf.dup_capacity;
    bp.tiles_x = u.tiles_x;
    bp.ty_lo = ty_lo;
    bp.ty_hi = ty_hi;
    bp.max_visible = p.max_visible;
    bp.indices_vec4 = (reinterpret_cast<uintptr_t>(p.sorted_indices) & 15u) == 0 &&
                      (p.sorted_indices_alt == nullptr || (reinterpret_cast<uintptr_t>(p.sorted_indices_alt) & 15u) == 0);

    const unsigned scan_grid = (unsigned)(((size_t)p.max_visible + kScanTile - 1) / kScanTile);
    if (bin_v1) {
        dup_scan_kernel<<<scan_grid > 0 ? scan_grid : 1, kScanThreads, 0, stream>>>(bp);
        dup_emit_kernel<<<num_sms * 8, 256, 0, stream>>>(bp);
    } else {
        // beyond 16.7 M splats a CTA no longer sums the tile sums itself (SB_BIN_TILES_PREFIX=1 forces the prefix kernel: tests)
        static const bool force_prefix = [] { const char* c = std::getenv("SB_BIN_TILES_PREFIX"); return c && std::atoi(c) != 0; }();
        bp.tiles_prefixed = (scan_grid > 4096 || force_prefix) ? 1u : 0u;
        dup_count_kernel<<<scan_grid > 0 ? scan_grid : 1, kScan2Threads, 0, stream>>>(bp);
        if (bp.tiles_prefixed) dup_tiles_kernel<<<1, 1024, 0, stream>>>(bp);
        dup_offsets_kernel<<<scan_grid > 0 ? scan_grid : 1, kScan2Threads, 0, stream>>>(bp);
        dup_emit2_kernel<<<num_sms * 12, 256, 0, stream>>>(bp);  // two waves of six CTAs per SM: the windows of a CTA are spread over the frame
    }
    if (p.events) cudaEventRecord(p.events[0], stream);

    e = launch_sort(p.buf.dup_keys, p.buf.dup_vals, p.buf.dup_count, (uint32_t)p.buf.dup_capacity, 0, bits, p.sort, num_sms, stream);
    if (e != cudaSuccess) return e;
    if (p.events) cudaEventRecord(p.events[1], stream);

    if (p.recs_map)
        tile_ranges_kernel<<<num_sms * 8, 256, 0, stream>>>(p.buf.dup_keys, p.buf.dup_count, ranges_rel);
    else
        gather_kernel<<<num_sms * 8, 256, 0, stream>>>(p.buf.dup_keys, p.buf.dup_vals, p.buf.dup_count, p.recs, p.buf.tile_recs,
                                                      ranges_rel);
    if (p.buf.tile_order)
        tile_order_kernel<<<1, 1024, 0, stream>>>(p.buf.tile_ranges, ty_lo * u.tiles_x, (ty_hi - ty_lo + 1) * u.tiles_x, p.buf.tile_order);
    if (p.events) cudaEventRecord(p.events[2], stream);
    if (p.split_raster && p.bin_done != nullptr && p.raster_stream != stream) {
        // binning done on this model's stream; its raster goes to the frame's draw-order stream
        e = cudaEventRecord(p.bin_done, stream);
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(p.raster_stream, p.bin_done, 0);
        if (e != cudaSuccess) return e;
        stream = p.raster_stream;
    }

    RasterKernelParams kp;
    kp.tile_order = p.buf.tile_order;
    kp.order_base = 0;
    kp.tile_recs = p.buf.tile_recs;
    kp.dup_vals = p.buf.dup_vals;
    kp.tile_ranges = p.buf.tile_ranges;
    kp.pixels = reinterpret_cast<uint8_t*>(t.d_pixels);
    kp.pitch = t.pitch_bytes;
    kp.width = u.width;
    kp.height = u.height;
    kp.row0 = row0;
    kp.rows = rows;
    kp.tiles_x = u.tiles_x;
    kp.ty_lo = ty_lo;
    kp.sd = u.std_dev;
    kp.sd2 = u.std_dev * u.std_dev;
    kp.outline = (u.std_dev - 0.1f) * (u.std_dev - 0.1f);
    kp.bgra = t.format == SB_TARGET_BGRA8_UNORM || t.format == SB_TARGET_BGRA8_UNORM_SRGB;
    kp.clear = p.clear;
    kp.obb_cull = p.obb_cull;
    kp.cut_k = p.cut_k;
    kp.no_discard = p.no_discard && u.mode == SB_MODE_SPLAT && (t.format == SB_TARGET_RGBA8_UNORM || t.format == SB_TARGET_BGRA8_UNORM) &&
                    !p.strict_exp && !p.depth && !p.counters && p.recs_map && u.std_dev * u.std_dev >= 6.3f;
    kp.depth = p.depth;
    kp.depth_pitch = p.depth_pitch;
    kp.depth_compare = p.depth_compare;
    kp.depth_write = p.depth_write;
    kp.pods = p.pods;
    kp.pod_stride = p.pod_stride;
    for (int i = 0; i < 16; i++) {
        kp.pv[i] = u.pv[i];
        kp.model[i] = u.model[i];
    }
    kp.counters = p.counters;
    // The tile rows can be rasterized as several launches (bands): in a pipelined view batch every band boundary is a
    // point where the other view's kernels get SM slots, instead of queueing behind one 8160-CTA grid.
    const uint32_t rows_total = ty_hi - ty_lo + 1;
    const uint32_t bands = (uint32_t)std::min<int>(std::max(p.raster_bands, 1), (int)rows_total);
    const bool srgb = t.format == SB_TARGET_RGBA8_UNORM_SRGB || t.format == SB_TARGET_BGRA8_UNORM_SRGB;
    const int fmt = srgb ? FMT_SRGB8
                    : (t.format == SB_TARGET_RGBA8_UNORM || t.format == SB_TARGET_BGRA8_UNORM) ? FMT_UNORM8
                    : t.format == SB_TARGET_RGBA16_FLOAT                                       ? FMT_F16
                                                                                               : FMT_F32;
    for (uint32_t band = 0; band < bands; band++) {
    const uint32_t y0 = ty_lo + rows_total * band / bands, y1 = ty_lo + rows_total * (band + 1) / bands;
    kp.ty_lo = y0;
    kp.order_base = (y0 - ty_lo) * u.tiles_x;  // with a schedule, a band is a slice of it
    const dim3 grid(u.tiles_x, y1 - y0);
#define SB_RASTER(M, F)                                                        \
    if (u.mode == M && fmt == F) {                                             \
        if (M == SB_MODE_SPLAT && p.strict_exp) launch_raster<M, F, true>(kp, p.recs_map, grid, stream); \
        else launch_raster<M, F, false>(kp, p.recs_map, grid, stream);                     \
    }
    SB_RASTER(SB_MODE_SPLAT, FMT_UNORM8)
    SB_RASTER(SB_MODE_SPLAT, FMT_F16)
    SB_RASTER(SB_MODE_SPLAT, FMT_F32)
    SB_RASTER(SB_MODE_ELLIPSE, FMT_UNORM8)
    SB_RASTER(SB_MODE_ELLIPSE, FMT_F16)
    SB_RASTER(SB_MODE_ELLIPSE, FMT_F32)
    SB_RASTER(SB_MODE_POINT, FMT_UNORM8)
    SB_RASTER(SB_MODE_POINT, FMT_F16)
    SB_RASTER(SB_MODE_POINT, FMT_F32)
    SB_RASTER(SB_MODE_SPLAT, FMT_SRGB8)
    SB_RASTER(SB_MODE_ELLIPSE, FMT_SRGB8)
    SB_RASTER(SB_MODE_POINT, FMT_SRGB8)
#undef SB_RASTER
    }
    if (p.events) cudaEventRecord(p.events[3], stream);
    return cudaGetLastError();
}

}  // namespace sb
