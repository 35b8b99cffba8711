// sb_sort.cu — K2/K3: onesweep LSD radix sort of (u32 key, u32 payload) pairs.
//
// Replaces the reference's 7-dispatch sort (zero_histograms, calculate_histogram,
// prefix_histogram, scatter_even/odd x4: src/shader/radix_sort.wgsl:78-529, recorded by
// src/radix_sorter.rs:526-621) with 1 histogram launch + 1 launch per 8-bit digit:
//   - K2 `histogram_kernel`: one read of the keys, all digit histograms at once;
//   - K3 `onesweep_kernel`: per tile, warp-level multi-split ranking (match.any), chained
//     scan with decoupled look-back over per-tile digit counts, shared-memory reorder and
//     coalesced scatter of keys + payload.
// The sort is STABLE (ties keep input order) like the reference's (radix_sort.wgsl:325-343),
// which is what makes the final order canonical (ascending Gaussian index within equal keys).
// The number of keys is read from device memory (the reference's indirect dispatch).
#include <algorithm>
#include <cstdlib>
#include <string>

#include "sb_internal.h"

namespace sb {

namespace {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 16;
constexpr int kSortTile = kSortThreads * kItems;  // pairs per tile
constexpr int kMaxPasses = 4;

constexpr uint32_t kLbAggregate = 1u << 30;
constexpr uint32_t kLbPrefix = 2u << 30;
constexpr uint32_t kLbValueMask = (1u << 30) - 1;

// internal buffer layout (u32 words):
//   [0, 4*256)                 global digit histograms per pass
//   [1024, 1024+4)             tile tickets per pass
//   [1040, ...)                look-back tables: pass-major, tile-major, 256 words each
constexpr size_t kHistWords = kMaxPasses * kRadix;
constexpr size_t kTicketOffset = kHistWords;
constexpr size_t kLookbackOffset = kHistWords + 16;

// Lanes of the warp holding the same digit as this lane (and the same validity), built from one
// ballot per digit bit: VOTE is a cheap uniform-datapath op, MATCH.ANY measured MIO-bound on B200.
template <int NBITS>
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, bool ok) {
    uint32_t peers = __ballot_sync(0xffffffffu, ok);
    if (!ok) peers = ~peers;
#pragma unroll
    for (int b = 0; b < NBITS; b++) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

// K2: one read of the keys, all digit histograms at once (shared atomics, global merge).
__global__ void __launch_bounds__(512) histogram_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ d_count,
                                                        uint32_t max_count, int begin_bit, int end_bit, int num_passes,
                                                        uint32_t* __restrict__ ghist) {
    __shared__ uint32_t hist[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t count = min(*d_count, max_count);
    const uint32_t nvec = count / 4;
    const uint4* kv = reinterpret_cast<const uint4*>(keys);
    // warp-uniform trip count so the warp votes see every lane
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < nvec; base += stride) {
        const uint32_t i = base + threadIdx.x;
        const bool ok = i < nvec;
        uint4 k = make_uint4(0, 0, 0, 0);
        if (ok) k = __ldg(&kv[i]);
        const uint32_t ks[4] = {k.x, k.y, k.z, k.w};
        for (int p = 0; p < num_passes; p++) {
            const int sh = begin_bit + p * kRadixBits;
            const uint32_t mask = (1u << min(kRadixBits, end_bit - sh)) - 1u;
            // one shared atomic per key and digit; when all 128 keys of the warp's four vectors hold the same digit
            // (the top byte of one frame's depth keys) a single lane adds 128 instead of 32 lanes serialising on one
            // address four times — one shuffle and one vote per pass, not per key
            const uint32_t d0 = (ks[0] >> sh) & mask, d1 = (ks[1] >> sh) & mask, d2 = (ks[2] >> sh) & mask, d3 = (ks[3] >> sh) & mask;
            const uint32_t ref = __shfl_sync(0xffffffffu, d0, 0);
            if (__all_sync(0xffffffffu, ok && d0 == ref && d1 == ref && d2 == ref && d3 == ref)) {
                if ((threadIdx.x & 31u) == 0) atomicAdd(&hist[p][ref], 128u);
            } else if (ok) {
                atomicAdd(&hist[p][d0], 1u);
                atomicAdd(&hist[p][d1], 1u);
                atomicAdd(&hist[p][d2], 1u);
                atomicAdd(&hist[p][d3], 1u);
            }
        }
    }
    if (blockIdx.x == 0) {
        for (uint32_t i = nvec * 4 + threadIdx.x; i < count; i += blockDim.x) {
            const uint32_t key = keys[i];
            for (int p = 0; p < num_passes; p++) {
                const int sh = begin_bit + p * kRadixBits;
                atomicAdd(&hist[p][(key >> sh) & ((1u << min(kRadixBits, end_bit - sh)) - 1u)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < num_passes * kRadix; i += blockDim.x) {
        const uint32_t v = (&hist[0][0])[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

struct SortSmem {
    union {
        struct {
            uint32_t warp_hist[kSortWarps][kRadix];
            uint32_t warp_mask[kSortWarps][2][kRadix];  // peer bitmasks per digit, two planes (even/odd item)
        };
        struct {
            uint32_t keys[kSortTile];
            uint32_t vals[kSortTile];
        } stage;
    };
    uint32_t digit_base[kRadix];   // global destination of the first key of each digit, minus its tile-local start
    uint32_t scan_a[kSortWarps];
    uint32_t scan_b[kSortWarps];
    uint32_t tile;
};

// Device-side sort state (after the histograms/tickets): lets a pass whose digit is the same for
// every key be SKIPPED (depth keys of one frame share their top byte most of the time) while the
// ping-pong direction stays consistent: [0] parity (0: data in keys/payload, 1: in the alt
// buffers), [1 + pass] CTAs finished in that pass.
constexpr size_t kStateOffset = kHistWords + 8;

template <int NBITS, bool FULL>
__device__ __forceinline__ void onesweep_tile(SortSmem& sm, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                              uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t count,
                                              int shift, const uint32_t* __restrict__ ghist, uint32_t* __restrict__ lookback,
                                              uint32_t tile) {
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t tile_base = tile * kSortTile;
    const uint32_t tile_count = FULL ? (uint32_t)kSortTile : (count - tile_base);
    constexpr uint32_t kMask = (1u << NBITS) - 1u;

    // ---- load keys (warp-striped: item i of lane l sits at warp_base + i*32 + l)
    uint32_t key[kItems];
    const uint32_t warp_base = tile_base + warp * (32 * kItems);
    const uint32_t* kp = keys_in + warp_base + lane;
#pragma unroll
    for (int i = 0; i < kItems; i++) key[i] = (FULL || warp_base + i * 32 + lane < count) ? __ldg(kp + i * 32) : 0xffffffffu;

    // ---- warp-level multi-split: rank of each key among equal digits of its warp (stable): one
    // ballot per digit bit finds the peers, the lowest peer claims the run with one shared atomic.
    uint32_t rank2[kItems / 2];  // two 16-bit ranks per register (a rank is < 8192)
    uint32_t* wh = sm.warp_hist[warp];
    uint32_t* wm = sm.warp_mask[warp][0];
    const uint32_t lt = lanemask_lt();
    const uint32_t my_bit = 1u << lane;
    // Peer mask through shared memory: every lane ORs its lane bit into mask[digit] (commutative,
    // hence deterministic), then reads the word back = the lanes holding the same digit.  ~18
    // instructions per key instead of ~54 for a ballot per digit bit; MATCH.ANY is MIO-bound on B200.
    // Two planes alternate so the leader's clear never races the next item's ORs.
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const bool ok = FULL || (warp_base + i * 32 + lane) < count;
        const uint32_t d = (key[i] >> shift) & kMask;
        uint32_t* plane = wm + (i & 1) * kRadix;
        if (ok) atomicOr(&plane[d], my_bit);
        __syncwarp();
        const uint32_t peers = ok ? plane[d] : my_bit;
        __syncwarp();  // every peer has read the mask before its leader clears it (racecheck: read/write hazard otherwise)
        const uint32_t below = __popc(peers & lt);
        uint32_t pre = 0;
        if (below == 0 && ok) {
            pre = atomicAdd(&wh[d], (uint32_t)__popc(peers));
            plane[d] = 0;
        }
        pre = __shfl_sync(0xffffffffu, pre, __ffs(peers) - 1);
        if (i & 1) rank2[i >> 1] |= (pre + below) << 16;
        else rank2[i >> 1] = pre + below;
    }
    // payload: loaded late so it is not live across the ranking loop; its latency hides behind the
    // digit scan and the look-back
    uint32_t val[kItems];
    const uint32_t* vp = vals_in + warp_base + lane;
#pragma unroll
    for (int i = 0; i < kItems; i++) val[i] = (FULL || warp_base + i * 32 + lane < count) ? __ldg(vp + i * 32) : 0u;
    __syncthreads();

    // ---- thread d < 256 owns digit d: scan over warps, publish the tile aggregate, look back
    uint32_t digit_count = 0, gcount = 0;
    if (tid < kRadix) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const uint32_t t = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = run;
            run += t;
        }
        digit_count = run;
        st_relaxed_u32(&lookback[(size_t)tile * kRadix + tid], (tile == 0 ? kLbPrefix : kLbAggregate) | digit_count);
        gcount = ghist[tid];
    }

    // exclusive scans over digits: tile-local start of each digit, and global digit start
    uint32_t local_excl, global_excl;
    {
        uint32_t a = digit_count, b = gcount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o);
            const uint32_t tb = __shfl_up_sync(0xffffffffu, b, o);
            if ((int)lane >= o) {
                a += ta;
                b += tb;
            }
        }
        if (lane == 31) {
            sm.scan_a[warp] = a;
            sm.scan_b[warp] = b;
        }
        __syncthreads();
        uint32_t wa = 0, wb = 0;
#pragma unroll
        for (int w = 0; w < kRadix / 32; w++) {
            if (w < (int)warp) {
                wa += sm.scan_a[w];
                wb += sm.scan_b[w];
            }
        }
        local_excl = a - digit_count + wa;
        global_excl = b - gcount + wb;
    }

    // fold the tile-local digit start into warp_hist: a key's tile-sorted position is
    // warp_hist[w][d] + rank (warp_hist[w][d] = exclusive count of digit d in warps < w)
    if (tid < kRadix) {
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) sm.warp_hist[w][tid] += local_excl;
    }
    __syncthreads();

    // ---- tile-sorted positions, then reorder through shared memory.  None of this needs the
    // global offsets, so it runs BEFORE the look-back: predecessors get time to publish.
    uint32_t base2[kItems / 2];
#pragma unroll
    for (int i = 0; i < kItems; i += 2)
        base2[i >> 1] = wh[(key[i] >> shift) & kMask] | (wh[(key[i + 1] >> shift) & kMask] << 16);
    __syncthreads();  // warp_hist/warp_mask are dead; their storage becomes the staging buffer
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        if (FULL || (warp_base + i * 32 + lane) < count) {
            const uint32_t pos = ((rank2[i >> 1] >> (16 * (i & 1))) & 0xffffu) + ((base2[i >> 1] >> (16 * (i & 1))) & 0xffffu);
            sm.stage.keys[pos] = key[i];
            sm.stage.vals[pos] = val[i];
        }
    }

    if (tid < kRadix) {
        // decoupled look-back.  Tiles of one wave run in lock-step, so the nearest predecessor is often not
        // published yet: poll only that one word (with a short sleep, spinning burnt 27 % of the issued
        // instructions), then read the next seven predecessors in one round trip.
        uint32_t tile_excl = 0;
        int t = (int)tile - 1;
        while (t >= 0) {
            const uint32_t v0 = ld_relaxed_u32(&lookback[(size_t)t * kRadix + tid]);
            if ((v0 >> 30) == 0) {
                __nanosleep(40);
                continue;
            }
            tile_excl += v0 & kLbValueMask;
            --t;
            if ((v0 >> 30) == 2) break;
            uint32_t v[7];
#pragma unroll
            for (int j = 0; j < 7; j++) v[j] = (t - j) >= 0 ? ld_relaxed_u32(&lookback[(size_t)(t - j) * kRadix + tid]) : kLbPrefix;
            bool done = false;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                if (done) break;
                if ((v[j] >> 30) == 0) break;  // not published yet: poll it at the top of the loop
                tile_excl += v[j] & kLbValueMask;
                --t;
                if ((v[j] >> 30) == 2) {
                    done = true;
                    t = -1;
                }
            }
        }
        if (tile > 0) st_relaxed_u32(&lookback[(size_t)tile * kRadix + tid], kLbPrefix | (tile_excl + digit_count));
        sm.digit_base[tid] = global_excl + tile_excl - local_excl;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t j = i * kSortThreads + tid;
        if (FULL || j < tile_count) {
            const uint32_t k = sm.stage.keys[j];
            const uint32_t dst = sm.digit_base[(k >> shift) & kMask] + j;
            keys_out[dst] = k;
            vals_out[dst] = sm.stage.vals[j];
        }
    }
}

// K3: one onesweep digit pass over NBITS significant digit bits.
template <int NBITS>
__global__ void __launch_bounds__(kSortThreads, 4)
    onesweep_kernel(uint32_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a, uint32_t* __restrict__ keys_b,
                    uint32_t* __restrict__ vals_b, const uint32_t* __restrict__ d_count, uint32_t max_count, int shift,
                    const uint32_t* __restrict__ ghist /* this pass, 256 */, uint32_t* __restrict__ ticket,
                    uint32_t* __restrict__ lookback /* this pass: tiles x 256 */, uint32_t* __restrict__ state, int pass) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SortSmem& sm = *reinterpret_cast<SortSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t count = min(*d_count, max_count);
    if (blockIdx.x * kSortTile >= count) return;  // surplus CTA: exactly ceil(count/tile) CTAs take a ticket

    // A pass whose digit is identical for all keys is the identity permutation: skip it.
    const int degenerate = __syncthreads_or(tid < kRadix && ghist[tid] == count);
    if (degenerate) return;

    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
    {
        uint4* z = reinterpret_cast<uint4*>(&sm.warp_hist[0][0]);
        for (int i = tid; i < kSortWarps * kRadix * 3 / 4; i += kSortThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint32_t parity = ld_relaxed_u32(&state[0]);  // stable during the pass: flipped by the last CTA only
    const uint32_t* kin = parity ? keys_b : keys_a;
    const uint32_t* vin = parity ? vals_b : vals_a;
    uint32_t* kout = parity ? keys_a : keys_b;
    uint32_t* vout = parity ? vals_a : vals_b;
    if ((tile + 1) * kSortTile <= count)
        onesweep_tile<NBITS, true>(sm, kin, vin, kout, vout, count, shift, ghist, lookback, tile);
    else
        onesweep_tile<NBITS, false>(sm, kin, vin, kout, vout, count, shift, ghist, lookback, tile);

    // the last CTA of the pass flips the ping-pong parity for the next kernel
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const uint32_t tiles = (count + kSortTile - 1) / kSortTile;
        if (atomicAdd(&state[1 + pass], 1u) == tiles - 1) st_relaxed_u32(&state[0], parity ^ 1u);
    }
}


constexpr size_t kPlanOffset = kHistWords + 4;  // [kPlanOffset + pass]: bit 0 = skip, bit 1 = input lives in the alt buffers

// lanes of the warp whose NBITS-bit digit equals this lane's: one vote per digit bit
template <int NBITS>
__device__ __forceinline__ uint32_t match_digit(uint32_t d) {
#ifdef SB_SORT_USE_MATCH
    return __match_any_sync(0xffffffffu, d);  // tuning variant, measured: 64 M keys 2.38 ms against 1.85 ms with the votes
#endif
    // differ |= lanes whose bit b differs from mine = vote ^ (my bit replicated).  Written in PTX: ptxas turns it
    // into one R2P for all bits + VOTE + predicated NOT + OR per bit; the C form costs six instructions per bit.
    uint32_t differ = 0u;
#pragma unroll
    for (int b = 0; b < NBITS; b++) {
        uint32_t x;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            ".reg .b32 t, m, r;\n"
            "and.b32 t, %1, %2;\n"
            "setp.ne.u32 p, t, 0;\n"
            "vote.sync.ballot.b32 m, p, 0xffffffff;\n"
            "selp.b32 r, -1, 0, p;\n"
            "xor.b32 %0, m, r;\n"
            "}\n"
            : "=r"(x)
            : "r"(d), "r"(1u << b));
        differ |= x;
    }
    return ~differ;
}

// ================================================================ K3 (v3): small-footprint one-tile-per-CTA pass
//
// Lessons of three persistent variants that were built and measured (profiles/r01_sort_variants.txt): any scheme in which a CTA holds
// a ticket while it still has to wait on another tile's look-back feeds that wait back into everybody's look-back and
// ends up slower than the plain kernel.  So: one tile per CTA, ticket at the start, nothing between ticket and
// aggregate but the key load and the ranking — and as many tiles in flight per SM as shared memory allows:
//   256 threads x 16 pairs (4096-pair tiles), <= 40 registers, 34 KB of shared memory -> six CTAs (48 warps) per SM;
//   the payload never enters registers: once a pair's tile-sorted position is known, cp.async (LDGSTS, 4 bytes)
//   moves it from global memory straight to its staging slot, overlapping the look-back;
//   count, plan, digit totals and the ticket are fetched in one round trip; ranking without shared-memory atomics.
constexpr int kV3Threads = 256;
constexpr int kV3Items = 16;
constexpr int kV3Tile = kV3Threads * kV3Items;  // 4096
constexpr int kV3Warps = kV3Threads / 32;
constexpr int kV3CtasPerSm = 6;

struct Sort3Smem {
    union {
        uint32_t warp_hist[kV3Warps][kRadix];
        struct {
            uint32_t keys[kV3Tile];
            uint32_t vals[kV3Tile];
        } stage;
    };
    uint32_t digit_local[kRadix];
    uint32_t digit_base[kRadix];
    uint32_t scan_a[kV3Warps];
    uint32_t scan_b[kV3Warps];
    uint32_t tile;
};

__device__ __forceinline__ void cp_async_4(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int NBITS, bool FULL>
__device__ __forceinline__ void onesweep3_tile(Sort3Smem& sm, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t count,
                                               int shift, uint32_t gcount, uint32_t* __restrict__ lookback, uint32_t tile) {
    constexpr uint32_t kMask = (1u << NBITS) - 1u;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t tile_base = tile * kV3Tile;
    const uint32_t tile_count = FULL ? (uint32_t)kV3Tile : (count - tile_base);

    uint32_t key[kV3Items];
    const uint32_t woff = warp * (32 * kV3Items) + lane;
    const uint32_t* kp = keys_in + tile_base + woff;
#pragma unroll
    for (int i = 0; i < kV3Items; i++) key[i] = (FULL || woff + i * 32 < tile_count) ? __ldg(kp + i * 32) : 0xffffffffu;

    uint32_t* wh = sm.warp_hist[warp];
    {
        uint4* z = reinterpret_cast<uint4*>(wh);
        z[lane] = make_uint4(0, 0, 0, 0);
        z[lane + 32] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();

    uint32_t rank2[kV3Items / 2];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < kV3Items; i++) {
        const bool ok = FULL || (woff + i * 32) < tile_count;
        const uint32_t d = (key[i] >> shift) & kMask;
        uint32_t peers = match_digit<NBITS>(d);
        if (!FULL) {
            const uint32_t okm = __ballot_sync(0xffffffffu, ok);
            peers &= ok ? okm : ~okm;
        }
        const uint32_t below = __popc(peers & lt);
        uint32_t pre = 0;
        if (below == 0 && ok) {
            pre = wh[d];
            wh[d] = pre + (uint32_t)__popc(peers);
        }
        __syncwarp();
        pre = __shfl_sync(0xffffffffu, pre, __ffs(peers) - 1);
        if (i & 1) rank2[i >> 1] |= (pre + below) << 16;
        else rank2[i >> 1] = pre + below;
    }
    __syncthreads();

    // ---- thread d owns digit d: scan over warps, publish the tile aggregate, scans over digits
    uint32_t digit_count, local_excl, global_excl, nearest = 0;
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kV3Warps; w++) {
            const uint32_t t = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = run;
            run += t;
        }
        digit_count = run;
        st_relaxed_u32(&lookback[(size_t)tile * kRadix + tid], (tile == 0 ? kLbPrefix : kLbAggregate) | digit_count);
        if (tile > 0) nearest = ld_relaxed_u32(&lookback[(size_t)(tile - 1) * kRadix + tid]);
        uint32_t a = digit_count, b = gcount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o);
            const uint32_t tb = __shfl_up_sync(0xffffffffu, b, o);
            if ((int)lane >= o) {
                a += ta;
                b += tb;
            }
        }
        if (lane == 31) {
            sm.scan_a[warp] = a;
            sm.scan_b[warp] = b;
        }
        __syncthreads();
        uint32_t wa = 0, wb = 0;
#pragma unroll
        for (int w = 0; w < kV3Warps; w++) {
            if (w < (int)warp) {
                wa += sm.scan_a[w];
                wb += sm.scan_b[w];
            }
        }
        local_excl = a - digit_count + wa;
        global_excl = b - gcount + wb;
        sm.digit_local[tid] = local_excl;
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < kV3Items; i++) {
        const uint32_t d = (key[i] >> shift) & kMask;
        rank2[i >> 1] += (wh[d] + sm.digit_local[d]) << (16 * (i & 1));
    }
    __syncthreads();  // histograms dead: the storage becomes the staging buffer
    const uint32_t vals_smem = smem_u32(sm.stage.vals);
    const uint32_t* vp = vals_in + tile_base + woff;
#pragma unroll
    for (int i = 0; i < kV3Items; i++) {
        if (FULL || (woff + i * 32) < tile_count) {
            const uint32_t pos = (rank2[i >> 1] >> (16 * (i & 1))) & 0xffffu;
            sm.stage.keys[pos] = key[i];
            cp_async_4(vals_smem + pos * 4u, vp + i * 32);  // payload: global -> staging slot, no register
        }
    }

    {
        uint32_t tile_excl = 0;
        int t = (int)tile - 1;
        bool have = true;
        while (t >= 0) {
            const uint32_t v0 = have ? nearest : ld_relaxed_u32(&lookback[(size_t)t * kRadix + tid]);
            have = false;
            if ((v0 >> 30) == 0) {
                __nanosleep(40);
                continue;
            }
            tile_excl += v0 & kLbValueMask;
            --t;
            if ((v0 >> 30) == 2) break;
            uint32_t v[7];
#pragma unroll
            for (int j = 0; j < 7; j++) v[j] = (t - j) >= 0 ? ld_relaxed_u32(&lookback[(size_t)(t - j) * kRadix + tid]) : kLbPrefix;
            bool done = false;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                if (done) break;
                if ((v[j] >> 30) == 0) break;
                tile_excl += v[j] & kLbValueMask;
                --t;
                if ((v[j] >> 30) == 2) {
                    done = true;
                    t = -1;
                }
            }
        }
        if (tile > 0) st_relaxed_u32(&lookback[(size_t)tile * kRadix + tid], kLbPrefix | (tile_excl + digit_count));
        sm.digit_base[tid] = global_excl + tile_excl - local_excl;
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kV3Items; i++) {
        const uint32_t j = i * kV3Threads + tid;
        if (FULL || j < tile_count) {
            const uint32_t k = sm.stage.keys[j];
            const uint32_t dst = sm.digit_base[(k >> shift) & kMask] + j;
            keys_out[dst] = k;
            vals_out[dst] = sm.stage.vals[j];
        }
    }
}

template <int NBITS>
__global__ void __launch_bounds__(kV3Threads, kV3CtasPerSm)
    onesweep3_kernel(uint32_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a, uint32_t* __restrict__ keys_b,
                     uint32_t* __restrict__ vals_b, const uint32_t* __restrict__ d_count, uint32_t max_count, int shift,
                     const uint32_t* __restrict__ ghist, uint32_t* __restrict__ ticket, uint32_t* __restrict__ lookback,
                     const uint32_t* __restrict__ plan) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Sort3Smem& sm = *reinterpret_cast<Sort3Smem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    // one round trip: ticket, count, plan and this digit's total are all in flight together
    uint32_t q = 0;
    if (tid == 0) q = atomicAdd(ticket, 1u);
    const uint32_t count = min(*d_count, max_count);
    const uint32_t pl = *plan;
    const uint32_t gcount = ghist[tid];
    if (pl & 1u) return;
    if (tid == 0) sm.tile = q;
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile * kV3Tile >= count) return;  // surplus CTA
    const bool alt_in = (pl >> 1) & 1u;
    const uint32_t* kin = alt_in ? keys_b : keys_a;
    const uint32_t* vin = alt_in ? vals_b : vals_a;
    uint32_t* kout = alt_in ? keys_a : keys_b;
    uint32_t* vout = alt_in ? vals_a : vals_b;
    if ((tile + 1) * kV3Tile <= count)
        onesweep3_tile<NBITS, true>(sm, kin, vin, kout, vout, count, shift, gcount, lookback, tile);
    else
        onesweep3_tile<NBITS, false>(sm, kin, vin, kout, vout, count, shift, gcount, lookback, tile);
}

// ================================================================ K2'/K3' (v4): key-adaptive onesweep
//
// Depth keys of one frame are f32 values of 1 - ndc.z: most of their 32 bits never vary (sign, the exponent's high bits, and —
// because 1 - ndc.z is a difference of numbers near 1 — the low mantissa bits).  K1 accumulates the OR and the OR-of-complements
// of the visible keys while it writes them; the bits that differ between at least two keys are `or & nand`, and the sort only
// has to cover bit range [lo, hi) of them: 17 bits for the 6 M bench scene from outside, 26 from inside.  One prologue kernel
// turns that into a plan — ceil((hi - lo) / 9) passes of at most 9 bits, widths balanced — builds every pass's histogram in one
// read of the keys and zeroes the look-back tables; the pass kernels read their shift and width from the plan.  Two passes
// instead of four (one of them skipped) for the bench scene, and 3 launches of real work instead of 7.
// Digits are at most 9 bits wide: 512 bins, two per thread (every per-digit quantity is a 64-bit vector: one ld/st each), 16 KB
// of warp histograms in the 32 KB the staging buffer needs anyway, six 256-thread CTAs per SM as before.
constexpr int kV4MaxBits = 9;
constexpr int kV4Bins = 1 << kV4MaxBits;  // 512
constexpr int kV4Threads = 512;
#ifndef SB_V4_ITEMS
#define SB_V4_ITEMS 8
#endif
constexpr int kV4Items = SB_V4_ITEMS;
constexpr int kV4Tile = kV4Threads * kV4Items;  // 4096
constexpr int kV4Warps = kV4Threads / 32;
#ifndef SB_V4_CTAS
#define SB_V4_CTAS 4
#endif
constexpr int kV4CtasPerSm = SB_V4_CTAS;

static_assert(kV4Bins == kV4Threads, "one digit per thread");

// prep words (sb_internal.h kSortPrep*): [0] or, [1] nand, [2] passes, [3] parity, [4..8) plan, [8..12) tickets, [12] prologue CTAs done, [16..) hist[4][512]
constexpr uint32_t kPlanValid = 1u << 17;
constexpr uint32_t kPlanAltIn = 1u << 16;

struct Sort4Smem {
    union {
        uint32_t warp_hist[kV4Warps][kV4Bins];  // 32 KB
        struct {
            uint32_t keys[kV4Tile];
            uint32_t vals[kV4Tile];
        } stage;                                 // 32 KB
    };
    uint32_t digit_off[kV4Bins];  // first the tile-local start of each digit, then (after the look-back) its global base
    uint32_t scan_a[kV4Warps];
    uint32_t tile;
};

// lanes whose digit (< 2^nbits, nbits <= 9, warp-uniform) equals this lane's: one vote per significant digit bit
__device__ __forceinline__ uint32_t match_digit9(uint32_t d, int nbits) {
    uint32_t differ = 0u;
#pragma unroll
    for (int b = 0; b < kV4MaxBits; b++) {
        if (b < kV4MaxBits - 1 || nbits == kV4MaxBits) {  // a bit above the digit's width is 0 in every lane: its vote changes nothing
            uint32_t x;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                ".reg .b32 t, m, r;\n"
                "and.b32 t, %1, %2;\n"
                "setp.ne.u32 p, t, 0;\n"
                "vote.sync.ballot.b32 m, p, 0xffffffff;\n"
                "selp.b32 r, -1, 0, p;\n"
                "xor.b32 %0, m, r;\n"
                "}\n"
                : "=r"(x)
                : "r"(d), "r"(1u << b));
            differ |= x;
        }
    }
    return ~differ;
}

// The sort plan from the varying-bit mask: passes, and for pass p its shift / width / input side.
struct SortPlan4 {
    int passes;
    int shift[kMaxPasses], bits[kMaxPasses];
};
__device__ __forceinline__ SortPlan4 make_plan4(uint32_t varying) {
    SortPlan4 pl;
    pl.passes = 0;
#pragma unroll
    for (int p = 0; p < kMaxPasses; p++) pl.shift[p] = pl.bits[p] = 0;
    if (varying == 0u) return pl;
    const int lo = __ffs(varying) - 1, hi = 32 - __clz(varying);
    const int nbits = hi - lo;
    pl.passes = (nbits + kV4MaxBits - 1) / kV4MaxBits;
    const int base = nbits / pl.passes, rem = nbits % pl.passes;
    int sh = lo;
#pragma unroll
    for (int p = 0; p < kMaxPasses; p++) {
        if (p < pl.passes) {
            pl.shift[p] = sh;
            pl.bits[p] = base + (p < rem ? 1 : 0);
            sh += pl.bits[p];
        }
    }
    return pl;
}

// K2': plan + all histograms in one read of the keys + look-back tables zeroed.  `prep` arrives zeroed apart from or / nand.
// explicit_mask != 0: sort exactly those bits (the standalone sorter's begin/end range, the tile sort's tile-id bits).
__global__ void __launch_bounds__(512) sort4_prologue_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ d_count,
                                                             uint32_t max_count, uint32_t* __restrict__ prep, uint32_t explicit_mask,
                                                             uint32_t* __restrict__ lookback, uint32_t tiles_alloc,
                                                             uint32_t* __restrict__ parity_out) {
    __shared__ uint32_t hist[kMaxPasses][kV4Bins];
    const uint32_t count = min(*d_count, max_count);
    const uint32_t varying = count < 2u ? 0u : (explicit_mask ? explicit_mask : (prep[kSortPrepOr] & prep[kSortPrepNand]));
    const SortPlan4 pl = make_plan4(varying);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t parity = 0;
        for (int p = 0; p < kMaxPasses; p++) {
            uint32_t w = 0;
            if (p < pl.passes) {
                w = (uint32_t)pl.shift[p] | ((uint32_t)pl.bits[p] << 8) | (parity ? kPlanAltIn : 0u) | kPlanValid;
                parity ^= 1u;
            }
            prep[kSortPrepPlan + p] = w;
        }
        prep[kSortPrepPasses] = (uint32_t)pl.passes;
        prep[kSortPrepParity] = parity;
        if (parity_out) *parity_out = parity;
    }
    if (pl.passes == 0) return;
    // zero the look-back tables of the tiles this sort will use
    {
        const uint32_t tiles = (count + kV4Tile - 1) / kV4Tile;
        for (int p = 0; p < pl.passes; p++) {
            uint4* ts = reinterpret_cast<uint4*>(lookback + (size_t)p * tiles_alloc * kV4Bins);
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < tiles * (kV4Bins / 4); i += gridDim.x * blockDim.x) ts[i] = make_uint4(0, 0, 0, 0);
        }
    }
    for (int i = threadIdx.x; i < kMaxPasses * kV4Bins; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t nvec = count / 4;
    const uint4* kv = reinterpret_cast<const uint4*>(keys);
    const uint32_t stride = gridDim.x * blockDim.x;
    constexpr int kU = 2;  // 128-bit loads in flight per thread: the kernel is bound by the latency of its one read of the keys
    for (uint32_t base = blockIdx.x * blockDim.x; base < nvec; base += kU * stride) {  // warp-uniform trip count (votes below)
        uint4 k[kU];
        bool ok[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const uint32_t i = base + u * stride + threadIdx.x;
            ok[u] = i < nvec;
            k[u] = ok[u] ? __ldg(&kv[i]) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
#pragma unroll
            for (int p = 0; p < kMaxPasses; p++) {
                if (p < pl.passes) {
                    const int sh = pl.shift[p];
                    const uint32_t mask = (1u << pl.bits[p]) - 1u;
                    const uint32_t d0 = (k[u].x >> sh) & mask, d1 = (k[u].y >> sh) & mask, d2 = (k[u].z >> sh) & mask, d3 = (k[u].w >> sh) & mask;
                    // a warp whose 128 keys share the digit (sorted or clustered input) adds once instead of serialising 128 atomics
                    const uint32_t ref = __shfl_sync(0xffffffffu, d0, 0);
                    if (__all_sync(0xffffffffu, ok[u] && d0 == ref && d1 == ref && d2 == ref && d3 == ref)) {
                        if ((threadIdx.x & 31u) == 0) atomicAdd(&hist[p][ref], 128u);
                    } else if (ok[u]) {
                        atomicAdd(&hist[p][d0], 1u);
                        atomicAdd(&hist[p][d1], 1u);
                        atomicAdd(&hist[p][d2], 1u);
                        atomicAdd(&hist[p][d3], 1u);
                    }
                }
            }
        }
    }
    if (blockIdx.x == 0) {
        for (uint32_t i = nvec * 4 + threadIdx.x; i < count; i += blockDim.x) {
            const uint32_t key = keys[i];
            for (int p = 0; p < pl.passes; p++) atomicAdd(&hist[p][(key >> pl.shift[p]) & ((1u << pl.bits[p]) - 1u)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pl.passes * kV4Bins; i += blockDim.x) {
        const uint32_t v = (&hist[0][0])[i];
        if (v) atomicAdd(&prep[kSortPrepHist + i], v);
    }
    // The last CTA to get here turns every pass's digit totals into exclusive prefixes (the global base of each digit), in place:
    // the pass kernels then need no scan of their own over the totals.
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        // one cumulative gpu-scope fence after the CTA barrier (the grid-sync pattern) instead of a membar in each of the 512
        // threads: ncu attributed 19 % of this kernel's stall samples to the per-thread fence
        __threadfence();
        s_last = atomicAdd(&prep[kSortPrepDone], 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    static_assert(kV4Bins == 512, "one bin per thread of the 512-thread prologue CTA");
    __shared__ uint32_t wsum[16];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int p = 0; p < pl.passes; p++) {
        const uint32_t c = __ldcg(&prep[kSortPrepHist + p * kV4Bins + threadIdx.x]);
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t base = 0;
        for (uint32_t w = 0; w < warp; w++) base += wsum[w];
        prep[kSortPrepHist + p * kV4Bins + threadIdx.x] = base + inc - c;
        __syncthreads();
    }
}

template <bool FULL>
__device__ __forceinline__ void onesweep4_tile(Sort4Smem& sm, const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t count,
                                               const int shift, const int nbits, const uint32_t* __restrict__ gbase,
                                               uint32_t* __restrict__ lookback, uint32_t tile) {
    const uint32_t kMask = (1u << nbits) - 1u;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t tile_base = tile * kV4Tile;
    const uint32_t tile_count = FULL ? (uint32_t)kV4Tile : (count - tile_base);

    uint32_t key[kV4Items];
    const uint32_t woff = warp * (32 * kV4Items) + lane;
    const uint32_t* kp = keys_in + tile_base + woff;
#pragma unroll
    for (int i = 0; i < kV4Items; i++) key[i] = (FULL || woff + i * 32 < tile_count) ? __ldg(kp + i * 32) : 0xffffffffu;

    uint32_t* wh = sm.warp_hist[warp];
    {
        uint4* z = reinterpret_cast<uint4*>(wh);
#pragma unroll
        for (int q = 0; q < kV4Bins / 128; q++) z[lane + 32 * q] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();

    // ---- warp-level multi-split: rank among the equal digits of the warp (stable), no shared-memory atomics
    uint32_t rank2[kV4Items / 2];
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < kV4Items; i++) {
        const bool ok = FULL || (woff + i * 32) < tile_count;
        const uint32_t d = (key[i] >> shift) & kMask;
        uint32_t peers = match_digit9(d, nbits);
        if (!FULL) {
            const uint32_t okm = __ballot_sync(0xffffffffu, ok);
            peers &= ok ? okm : ~okm;
        }
        const uint32_t below = __popc(peers & lt);
        uint32_t pre = 0;
        if (below == 0 && ok) {
            pre = wh[d];
            wh[d] = pre + (uint32_t)__popc(peers);
        }
        __syncwarp();
        pre = __shfl_sync(0xffffffffu, pre, __ffs(peers) - 1);
        if (i & 1) rank2[i >> 1] |= (pre + below) << 16;
        else rank2[i >> 1] = pre + below;
    }
    __syncthreads();

    // ---- thread d owns digit d: scan over the warps, publish the tile aggregate, scan over the digits
    uint32_t digit_count, local_excl, nearest = 0;
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kV4Warps; w++) {
            const uint32_t t = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = run;
            run += t;
        }
        digit_count = run;
        st_relaxed_u32(&lookback[(size_t)tile * kV4Bins + tid], (tile == 0 ? kLbPrefix : kLbAggregate) | digit_count);
        if (tile > 0) nearest = ld_relaxed_u32(&lookback[(size_t)(tile - 1) * kV4Bins + tid]);
        uint32_t a = digit_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ta = __shfl_up_sync(0xffffffffu, a, o);
            if ((int)lane >= o) a += ta;
        }
        if (lane == 31) sm.scan_a[warp] = a;
        __syncthreads();
        // totals of the warps before mine: lanes 0..15 fetch one each, one REDUX adds them
        const uint32_t part = (lane < warp && lane < (uint32_t)kV4Warps) ? sm.scan_a[lane] : 0u;
        local_excl = a - digit_count + __reduce_add_sync(0xffffffffu, part);
        sm.digit_off[tid] = local_excl;
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < kV4Items; i++) {
        const uint32_t d = (key[i] >> shift) & kMask;
        rank2[i >> 1] += (wh[d] + sm.digit_off[d]) << (16 * (i & 1));
    }
    __syncthreads();  // histograms and local starts are dead: the storage becomes the staging buffer / the global bases
    const uint32_t vals_smem = smem_u32(sm.stage.vals);
    const uint32_t* vp = vals_in + tile_base + woff;
#pragma unroll
    for (int i = 0; i < kV4Items; i++) {
        if (FULL || (woff + i * 32) < tile_count) {
            const uint32_t pos = (rank2[i >> 1] >> (16 * (i & 1))) & 0xffffu;
            sm.stage.keys[pos] = key[i];
            cp_async_4(vals_smem + pos * 4u, vp + i * 32);  // payload: global -> staging slot, no register
        }
    }

    {
        // decoupled look-back (thread per digit): the nearest predecessor was prefetched before the staging writes; then
        // seven per round trip.  (A two-level variant — groups of 16 tiles with group aggregates / prefixes — was built and
        // measured slower at 6 M and at 64 M keys: profiles/r02_sort_variants.txt.)
        const uint32_t global_excl = gbase[tid];  // exclusive digit prefix (prologue); the L2 round trip overlaps the look-back
        uint32_t tile_excl = 0;
        int t = (int)tile - 1;
        bool have = true;
        while (t >= 0) {
            const uint32_t v0 = have ? nearest : ld_relaxed_u32(&lookback[(size_t)t * kV4Bins + tid]);
            have = false;
            if ((v0 >> 30) == 0) {
                __nanosleep(40);
                continue;
            }
            tile_excl += v0 & kLbValueMask;
            --t;
            if ((v0 >> 30) == 2) break;
            uint32_t v[7];
#pragma unroll
            for (int j = 0; j < 7; j++) v[j] = (t - j) >= 0 ? ld_relaxed_u32(&lookback[(size_t)(t - j) * kV4Bins + tid]) : kLbPrefix;
            bool done = false;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                if (done) break;
                if ((v[j] >> 30) == 0) break;
                tile_excl += v[j] & kLbValueMask;
                --t;
                if ((v[j] >> 30) == 2) {
                    done = true;
                    t = -1;
                }
            }
        }
        if (tile > 0) st_relaxed_u32(&lookback[(size_t)tile * kV4Bins + tid], kLbPrefix | (tile_excl + digit_count));
        sm.digit_off[tid] = global_excl + tile_excl - local_excl;
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kV4Items; i++) {
        const uint32_t j = i * kV4Threads + tid;
        if (FULL || j < tile_count) {
            const uint32_t k = sm.stage.keys[j];
            const uint32_t dst = sm.digit_off[(k >> shift) & kMask] + j;
            keys_out[dst] = k;
            vals_out[dst] = sm.stage.vals[j];
        }
    }
}

// K3': one digit pass; shift, width and input side come from the plan the prologue wrote.
__global__ void __launch_bounds__(kV4Threads, kV4CtasPerSm)
    onesweep4_kernel(uint32_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a, uint32_t* __restrict__ keys_b,
                     uint32_t* __restrict__ vals_b, const uint32_t* __restrict__ d_count, uint32_t max_count,
                     uint32_t* __restrict__ prep, uint32_t* __restrict__ lookback_base, uint32_t tiles_alloc, int pass) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Sort4Smem& sm = *reinterpret_cast<Sort4Smem*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    // launched with programmatic stream serialization: the launch itself overlaps the previous kernel's tail; nothing the
    // previous kernel wrote is read before this returns (a no-op for an ordinary launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t pl = prep[kSortPrepPlan + pass];
    if (!(pl & kPlanValid)) return;  // fewer passes than launches: nothing to do
    // one round trip: ticket and count are in flight together
    uint32_t q = 0;
    if (tid == 0) q = atomicAdd(&prep[kSortPrepTickets + pass], 1u);
    const uint32_t count = min(*d_count, max_count);
    if (tid == 0) sm.tile = q;
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile * kV4Tile >= count) return;  // surplus CTA
    const int shift = (int)(pl & 0xffu), nbits = (int)((pl >> 8) & 0xffu);
    const bool alt_in = (pl & kPlanAltIn) != 0;
    const uint32_t* kin = alt_in ? keys_b : keys_a;
    const uint32_t* vin = alt_in ? vals_b : vals_a;
    uint32_t* kout = alt_in ? keys_a : keys_b;
    uint32_t* vout = alt_in ? vals_a : vals_b;
    uint32_t* lookback = lookback_base + (size_t)pass * tiles_alloc * kV4Bins;
    const uint32_t* gbase = prep + kSortPrepHist + pass * kV4Bins;
    if ((tile + 1) * kV4Tile <= count)
        onesweep4_tile<true>(sm, kin, vin, kout, vout, count, shift, nbits, gbase, lookback, tile);
    else
        onesweep4_tile<false>(sm, kin, vin, kout, vout, count, shift, nbits, gbase, lookback, tile);
}

// One block after the histograms: which passes are the identity (one digit holds every key) and where each
// remaining pass reads its input; state[0] = 1 when the result ends in the alt buffers.
__global__ void sort_plan_kernel(uint32_t* __restrict__ internal, const uint32_t* __restrict__ d_count, uint32_t max_count, int num_passes,
                                 uint32_t* __restrict__ parity_out) {
    __shared__ int skip[kMaxPasses];
    const uint32_t count = min(*d_count, max_count);
    if (threadIdx.x < kMaxPasses) skip[threadIdx.x] = 0;
    __syncthreads();
    for (int p = 0; p < num_passes; p++)
        if (internal[p * kRadix + threadIdx.x] == count) skip[p] = 1;  // threadIdx.x < 256
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t parity = 0;
        for (int p = 0; p < num_passes; p++) {
            const uint32_t s = (count == 0 || skip[p]) ? 1u : 0u;
            internal[kPlanOffset + p] = s | (parity << 1);
            if (!s) parity ^= 1u;
        }
        internal[kStateOffset] = parity;
        if (parity_out) *parity_out = parity;
    }
}

// brings the result home when an odd number of passes ran (parity == 1)
__global__ void sort_finish_kernel(uint32_t* __restrict__ keys_a, uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ keys_b,
                                   const uint32_t* __restrict__ vals_b, const uint32_t* __restrict__ d_count, uint32_t max_count,
                                   const uint32_t* __restrict__ state) {
    if (state[0] == 0) return;
    const uint32_t count = min(*d_count, max_count);
    const uint32_t nvec = count / 4;
    const uint4* kb = reinterpret_cast<const uint4*>(keys_b);
    const uint4* vb = reinterpret_cast<const uint4*>(vals_b);
    uint4* ka = reinterpret_cast<uint4*>(keys_a);
    uint4* va = reinterpret_cast<uint4*>(vals_a);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        ka[i] = kb[i];
        va[i] = vb[i];
    }
    if (blockIdx.x == 0)
        for (uint32_t i = nvec * 4 + threadIdx.x; i < count; i += blockDim.x) {
            keys_a[i] = keys_b[i];
            vals_a[i] = vals_b[i];
        }
}

// zeroes the histograms, tickets and the part of the look-back tables this sort will use
__global__ void sort_init_kernel(uint32_t* __restrict__ internal, const uint32_t* __restrict__ d_count, uint32_t max_count,
                                 int num_passes, uint32_t tiles_alloc, uint32_t tile_size) {
    const uint32_t count = min(*d_count, max_count);
    const uint32_t tiles = (count + tile_size - 1) / tile_size;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < kLookbackOffset; i += gridDim.x * blockDim.x) internal[i] = 0;
    for (int p = 0; p < num_passes; p++) {
        uint32_t* lb = internal + kLookbackOffset + (size_t)p * tiles_alloc * kRadix;
        const uint32_t words = tiles * kRadix;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) lb[i] = 0;
    }
}

template <int NBITS>
void launch_pass_n(unsigned tiles, cudaStream_t stream, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb,
                   const uint32_t* d_count, uint32_t max_count, int shift, const uint32_t* ghist, uint32_t* ticket, uint32_t* lb,
                   uint32_t* state, int pass) {
    onesweep_kernel<NBITS><<<tiles, kSortThreads, sizeof(SortSmem), stream>>>(ka, va, kb, vb, d_count, max_count, shift, ghist, ticket,
                                                                            lb, state, pass);
}

void launch_pass(int nbits, unsigned tiles, cudaStream_t stream, uint32_t* kin, uint32_t* vin, uint32_t* kout, uint32_t* vout,
                 const uint32_t* d_count, uint32_t max_count, int shift, const uint32_t* ghist, uint32_t* ticket, uint32_t* lb,
                 uint32_t* state, int pass) {
    switch (nbits) {
#define SB_PASS(N) case N: launch_pass_n<N>(tiles, stream, kin, vin, kout, vout, d_count, max_count, shift, ghist, ticket, lb, state, pass); break;
        SB_PASS(1) SB_PASS(2) SB_PASS(3) SB_PASS(4) SB_PASS(5) SB_PASS(6) SB_PASS(7)
        default: launch_pass_n<8>(tiles, stream, kin, vin, kout, vout, d_count, max_count, shift, ghist, ticket, lb, state, pass); break;
#undef SB_PASS
    }
}

template <int NBITS>
void launch_pass3_n(unsigned grid, cudaStream_t stream, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb,
                    const uint32_t* d_count, uint32_t max_count, int shift, const uint32_t* ghist, uint32_t* ticket, uint32_t* lb,
                    const uint32_t* plan) {
    onesweep3_kernel<NBITS><<<grid, kV3Threads, sizeof(Sort3Smem), stream>>>(ka, va, kb, vb, d_count, max_count, shift, ghist, ticket, lb, plan);
}

void launch_pass3(int nbits, unsigned grid, cudaStream_t stream, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb,
                  const uint32_t* d_count, uint32_t max_count, int shift, const uint32_t* ghist, uint32_t* ticket, uint32_t* lb,
                  const uint32_t* plan) {
    switch (nbits) {
#define SB_PASS(N) case N: launch_pass3_n<N>(grid, stream, ka, va, kb, vb, d_count, max_count, shift, ghist, ticket, lb, plan); break;
        SB_PASS(1) SB_PASS(2) SB_PASS(3) SB_PASS(4) SB_PASS(5) SB_PASS(6) SB_PASS(7)
        default: launch_pass3_n<8>(grid, stream, ka, va, kb, vb, d_count, max_count, shift, ghist, ticket, lb, plan); break;
#undef SB_PASS
    }
}

cudaError_t set_smem_all() {
    cudaError_t e = cudaSuccess;
#define SB_ATTR(N)                                                                                                                   \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(onesweep_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem)); \
    if (e == cudaSuccess)                                                                                                            \
        e = cudaFuncSetAttribute(onesweep3_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sort3Smem));
    SB_ATTR(1) SB_ATTR(2) SB_ATTR(3) SB_ATTR(4) SB_ATTR(5) SB_ATTR(6) SB_ATTR(7) SB_ATTR(8)
#undef SB_ATTR
    return e;
}

// SB_SORT_IMPL = v1 (8192-pair tiles, three CTAs per SM, shared-memory peer masks) | v3 (4096-pair tiles, six CTAs per SM,
// vote ranking, cp.async payload).  Default: v3 for sorts of three or more digit passes (the depth sort: +5 % at 64 M keys),
// v1 for the two-pass tile sort, whose 5-bit second pass measured 8 % faster with the larger tiles.
int sort_impl_forced() {
    static const int forced = [] {
        const char* c = std::getenv("SB_SORT_IMPL");
        if (c && std::string(c) == "v1") return 1;
        if (c && std::string(c) == "v3") return 3;
        if (c && std::string(c) == "v4") return 4;
        return 0;
    }();
    return forced;
}
int sort_impl(int num_passes) { return sort_impl_forced() ? sort_impl_forced() : (num_passes >= 3 ? 3 : 1); }
constexpr uint32_t kSmallSort = 1500000;

}  // namespace

size_t sort_internal_bytes(uint32_t capacity) {
    const size_t tiles = ((size_t)capacity + kV3Tile - 1) / kV3Tile;  // the smaller of the tile sizes
    const size_t v3 = (kLookbackOffset + (size_t)kMaxPasses * tiles * kRadix) * sizeof(uint32_t);
    const size_t v4 = ((size_t)kSortPrepWords + (size_t)kMaxPasses * tiles * kV4Bins) * sizeof(uint32_t);
    return v3 > v4 ? v3 : v4;
}

size_t sort_prep_bytes() { return (size_t)kSortPrepWords * sizeof(uint32_t); }

cudaError_t launch_sort_adaptive(uint32_t* keys, uint32_t* payload, const uint32_t* d_count, uint32_t max_count, uint32_t* prep,
                                 bool prep_zeroed, int begin_bit, int end_bit, const SortScratch& scratch, int num_sms,
                                 cudaStream_t stream, uint32_t* parity_out) {
    if (max_count == 0) return parity_out ? cudaMemsetAsync(parity_out, 0, 4, stream) : cudaSuccess;
    const size_t tiles = ((size_t)max_count + kV4Tile - 1) / kV4Tile;
    uint32_t explicit_mask = 0;
    if (begin_bit >= 0) {
        if (end_bit > 32 || begin_bit >= end_bit) return cudaErrorInvalidValue;
        explicit_mask = (end_bit - begin_bit == 32) ? 0xffffffffu : (((1u << (end_bit - begin_bit)) - 1u) << begin_bit);
    }
    uint32_t* lookback = scratch.internal;
    if (!prep) {  // no caller-owned prep words: they live in front of the look-back tables
        prep = scratch.internal;
        lookback = scratch.internal + kSortPrepWords;
        prep_zeroed = false;
    }
    const size_t need = ((size_t)(lookback - scratch.internal) + (size_t)kMaxPasses * tiles * kV4Bins) * sizeof(uint32_t);
    if (need > scratch.internal_bytes) return cudaErrorInvalidValue;
    cudaError_t e;
    {   // per-device function attribute
        static bool configured[64] = {};
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            e = cudaFuncSetAttribute(onesweep4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sort4Smem));
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    if (!prep_zeroed) {
        // keep or / nand (K1 wrote them) when the caller owns the prep words; everything else starts from zero
        uint32_t* from = (prep == scratch.internal) ? prep : prep + kSortPrepPasses;
        e = cudaMemsetAsync(from, 0, (size_t)(prep + kSortPrepWords - from) * sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
    }
    const int pro_grid = (int)std::min<size_t>((size_t)num_sms * 4, (tiles * kV4Tile / 4 + 511) / 512);
    sort4_prologue_kernel<<<pro_grid > 0 ? pro_grid : 1, 512, 0, stream>>>(keys, d_count, max_count, prep, explicit_mask, lookback,
                                                                          (uint32_t)tiles, parity_out);
    // the number of passes is decided on the device; launches beyond it exit at once
    const int max_passes = begin_bit >= 0 ? (end_bit - begin_bit + kV4MaxBits - 1) / kV4MaxBits : kMaxPasses;
    static const bool pdl = [] { const char* c = std::getenv("SB_SORT_PDL"); return !(c && std::string(c) == "0"); }();
    for (int p = 0; p < max_passes; p++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)tiles);
        cfg.blockDim = dim3(kV4Threads);
        cfg.dynamicSmemBytes = sizeof(Sort4Smem);
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = pdl ? 1 : 0;
        e = cudaLaunchKernelEx(&cfg, onesweep4_kernel, keys, payload, scratch.keys_alt, scratch.payload_alt, d_count, max_count, prep,
                               lookback, (uint32_t)tiles, p);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

cudaError_t launch_sort_finish(uint32_t* keys, uint32_t* payload, const SortScratch& scratch, const uint32_t* d_count, uint32_t max_count,
                               const uint32_t* parity, int num_sms, cudaStream_t stream) {
    if (max_count == 0) return cudaSuccess;
    sort_finish_kernel<<<num_sms * 4, 256, 0, stream>>>(keys, payload, scratch.keys_alt, scratch.payload_alt, d_count, max_count, parity);
    return cudaGetLastError();
}

cudaError_t launch_sort(uint32_t* keys, uint32_t* payload, const uint32_t* d_count, uint32_t max_count, int begin_bit,
                        int end_bit, const SortScratch& scratch, int num_sms, cudaStream_t stream, uint32_t* parity_out,
                        uint32_t expected_count) {
    if (max_count == 0) return cudaSuccess;
    const int num_passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
    if (num_passes < 1 || num_passes > kMaxPasses) return cudaErrorInvalidValue;
    if (sort_impl_forced() == 4) {  // the key-adaptive kernels on an explicit bit range (<= 9-bit digits)
        uint32_t* parity = parity_out ? parity_out : scratch.internal + kSortPrepParity;
        cudaError_t e4 = launch_sort_adaptive(keys, payload, d_count, max_count, nullptr, false, begin_bit, end_bit, scratch, num_sms, stream, parity_out);
        if (e4 != cudaSuccess || parity_out) return e4;
        return launch_sort_finish(keys, payload, scratch, d_count, max_count, parity, num_sms, stream);
    }
    int impl = sort_impl(num_passes);
    // Few keys: a pass is bound by the look-back chain (one hop per tile), so the 8192-pair kernel wins (measured: 0.63 M depth
    // keys 0.106 ms with the 4096-pair kernel, 0.076 ms with this one; 6 M keys 0.173 against 0.203 ms).
    if (impl == 3 && expected_count != 0 && expected_count < kSmallSort && sort_impl_forced() == 0) impl = 1;
    const bool v2 = impl == 3;
    const size_t tile_size = impl == 3 ? kV3Tile : kSortTile;
    const size_t tiles = ((size_t)max_count + tile_size - 1) / tile_size;
    const size_t need = (kLookbackOffset + (size_t)num_passes * tiles * kRadix) * sizeof(uint32_t);
    if (need > scratch.internal_bytes) return cudaErrorInvalidValue;
    cudaError_t e;
    {   // the opt-in shared-memory size is a per-device function attribute: set once per device, not once per process
        static bool configured[64] = {};
        int dev = 0;
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            e = set_smem_all();
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    uint32_t* ghist = scratch.internal;
    uint32_t* tickets = scratch.internal + kTicketOffset;
    uint32_t* lookback = scratch.internal + kLookbackOffset;
    sort_init_kernel<<<num_sms * 2, 256, 0, stream>>>(scratch.internal, d_count, max_count, num_passes, (uint32_t)tiles, (uint32_t)tile_size);
    const int hist_grid = (int)min((size_t)num_sms * 4, (tiles * tile_size / 4 + 511) / 512);
    histogram_kernel<<<hist_grid > 0 ? hist_grid : 1, 512, 0, stream>>>(keys, d_count, max_count, begin_bit, end_bit, num_passes, ghist);

    uint32_t* state = scratch.internal + kStateOffset;
    if (v2) {
        sort_plan_kernel<<<1, kRadix, 0, stream>>>(scratch.internal, d_count, max_count, num_passes, parity_out);
        const unsigned grid = (unsigned)tiles;
        for (int p = 0; p < num_passes; p++) {
            const int nbits = min(kRadixBits, end_bit - (begin_bit + p * kRadixBits));
            launch_pass3(nbits, grid, stream, keys, payload, scratch.keys_alt, scratch.payload_alt, d_count, max_count,
                         begin_bit + p * kRadixBits, ghist + p * kRadix, tickets + p, lookback + (size_t)p * tiles * kRadix,
                         scratch.internal + kPlanOffset + p);
        }
    } else {
        for (int p = 0; p < num_passes; p++) {
            const int nbits = min(kRadixBits, end_bit - (begin_bit + p * kRadixBits));
            launch_pass(nbits, (unsigned)tiles, stream, keys, payload, scratch.keys_alt, scratch.payload_alt, d_count, max_count,
                        begin_bit + p * kRadixBits, ghist + p * kRadix, tickets + p, lookback + (size_t)p * tiles * kRadix, state, p);
        }
    }
    // result sits in the alt buffers when an odd number of passes actually ran: bring it home — unless the caller takes
    // the parity and reads the result where it lies (v3 plan only; launch_sort_finish brings it home later)
    if (!(v2 && parity_out))
        sort_finish_kernel<<<num_sms * 4, 256, 0, stream>>>(keys, payload, scratch.keys_alt, scratch.payload_alt, d_count, max_count, state);
    if (!v2 && parity_out) return cudaMemsetAsync(parity_out, 0, 4, stream);
    return cudaGetLastError();
}

}  // namespace sb
