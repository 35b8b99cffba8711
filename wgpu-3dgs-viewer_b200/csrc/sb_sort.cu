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
#ifndef SB_MATCH_EVERY
#define SB_MATCH_EVERY 0
#endif
constexpr int kMatchEvery = SB_MATCH_EVERY;  // every k-th item is ranked with MATCH.ANY (0 = never)

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
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // one shared atomic per key and digit; when the whole warp holds the same digit
                // (the top byte of one frame's depth keys) a single lane adds 32 instead of 32
                // lanes serialising on one address
                const uint32_t d = (ks[j] >> sh) & mask;
                const uint32_t d0 = __shfl_sync(0xffffffffu, d, 0);
                if (__all_sync(0xffffffffu, ok && d == d0)) {
                    if ((threadIdx.x & 31u) == 0) atomicAdd(&hist[p][d0], 32u);
                } else if (ok) {
                    atomicAdd(&hist[p][d], 1u);
                }
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
                                 int num_passes, uint32_t tiles_alloc) {
    const uint32_t count = min(*d_count, max_count);
    const uint32_t tiles = (count + kSortTile - 1) / kSortTile;
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

cudaError_t set_smem_all() {
    cudaError_t e = cudaSuccess;
#define SB_ATTR(N) if (e == cudaSuccess) e = cudaFuncSetAttribute(onesweep_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
    SB_ATTR(1) SB_ATTR(2) SB_ATTR(3) SB_ATTR(4) SB_ATTR(5) SB_ATTR(6) SB_ATTR(7) SB_ATTR(8)
#undef SB_ATTR
    return e;
}

}  // namespace

size_t sort_internal_bytes(uint32_t capacity) {
    const size_t tiles = ((size_t)capacity + kSortTile - 1) / kSortTile;
    return (kLookbackOffset + (size_t)kMaxPasses * tiles * kRadix) * sizeof(uint32_t);
}

cudaError_t launch_sort(uint32_t* keys, uint32_t* payload, const uint32_t* d_count, uint32_t max_count, int begin_bit,
                        int end_bit, const SortScratch& scratch, int num_sms, cudaStream_t stream) {
    if (max_count == 0) return cudaSuccess;
    const int num_passes = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
    if (num_passes < 1 || num_passes > kMaxPasses) return cudaErrorInvalidValue;
    const size_t tiles = ((size_t)max_count + kSortTile - 1) / kSortTile;
    const size_t need = (kLookbackOffset + (size_t)num_passes * tiles * kRadix) * sizeof(uint32_t);
    if (need > scratch.internal_bytes) return cudaErrorInvalidValue;
    cudaError_t e;
    static bool configured = false;
    if (!configured) {
        e = set_smem_all();
        if (e != cudaSuccess) return e;
        configured = true;
    }
    sort_init_kernel<<<num_sms * 2, 256, 0, stream>>>(scratch.internal, d_count, max_count, num_passes, (uint32_t)tiles);
    uint32_t* ghist = scratch.internal;
    uint32_t* tickets = scratch.internal + kTicketOffset;
    uint32_t* lookback = scratch.internal + kLookbackOffset;

    const int hist_grid = (int)min((size_t)num_sms * 2, (tiles * kSortTile / 4 + 511) / 512);
    histogram_kernel<<<hist_grid > 0 ? hist_grid : 1, 512, 0, stream>>>(keys, d_count, max_count, begin_bit, end_bit, num_passes, ghist);

    uint32_t* state = scratch.internal + kStateOffset;
    for (int p = 0; p < num_passes; p++) {
        const int nbits = min(kRadixBits, end_bit - (begin_bit + p * kRadixBits));
        launch_pass(nbits, (unsigned)tiles, stream, keys, payload, scratch.keys_alt, scratch.payload_alt, d_count, max_count,
                    begin_bit + p * kRadixBits, ghist + p * kRadix, tickets + p, lookback + (size_t)p * tiles * kRadix, state, p);
    }
    // result sits in the alt buffers when an odd number of passes actually ran: bring it home
    sort_finish_kernel<<<num_sms * 4, 256, 0, stream>>>(keys, payload, scratch.keys_alt, scratch.payload_alt, d_count, max_count, state);
    return cudaGetLastError();
}

}  // namespace sb
