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
constexpr int kSortTile = kSortThreads * kItems;  // 4096 pairs per tile
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

__global__ void __launch_bounds__(256) histogram_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ d_count,
                                                        uint32_t max_count, int begin_bit, int num_passes,
                                                        uint32_t* __restrict__ ghist) {
    __shared__ uint32_t hist[kMaxPasses][kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t count = min(*d_count, max_count);
    const uint32_t nvec = count / 4;
    const uint4* kv = reinterpret_cast<const uint4*>(keys);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        const uint4 k = __ldg(&kv[i]);
        const uint32_t ks[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t key = ks[j];
            for (int p = 0; p < num_passes; p++) atomicAdd(&hist[p][(key >> (begin_bit + p * kRadixBits)) & (kRadix - 1)], 1u);
        }
    }
    if (blockIdx.x == 0) {
        for (uint32_t i = nvec * 4 + threadIdx.x; i < count; i += blockDim.x) {
            const uint32_t key = keys[i];
            for (int p = 0; p < num_passes; p++) atomicAdd(&hist[p][(key >> (begin_bit + p * kRadixBits)) & (kRadix - 1)], 1u);
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
        uint32_t warp_hist[kSortWarps][kRadix + 32];  // +32: bin 256 collects out-of-range lanes
        struct {
            uint32_t keys[kSortTile];
            uint32_t vals[kSortTile];
        } stage;
    };
    uint32_t digit_base[kRadix];   // global destination of the first key of each digit, minus its tile-local start
    uint32_t scan_tmp[kSortWarps];
    uint32_t tile;
};

__global__ void __launch_bounds__(kSortThreads)
    onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                    uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ d_count, uint32_t max_count, int shift,
                    const uint32_t* __restrict__ ghist /* this pass, 256 */, uint32_t* __restrict__ ticket,
                    uint32_t* __restrict__ lookback /* this pass: tiles x 256 */) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SortSmem& sm = *reinterpret_cast<SortSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t count = min(*d_count, max_count);

    if (tid == 0) sm.tile = atomicAdd(ticket, 1u);
    for (int i = tid; i < kSortWarps * (kRadix + 32); i += kSortThreads) (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    const uint32_t tile_base = tile * kSortTile;
    if (tile_base >= count) return;
    const uint32_t tile_count = min((uint32_t)kSortTile, count - tile_base);

    // ---- load (warp-striped: item i of lane l sits at warp_base + i*32 + l)
    uint32_t key[kItems], val[kItems];
    const uint32_t warp_base = tile_base + warp * (32 * kItems);
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t idx = warp_base + i * 32 + lane;
        const bool ok = idx < count;
        key[i] = ok ? __ldg(&keys_in[idx]) : 0xffffffffu;
        val[i] = ok ? __ldg(&vals_in[idx]) : 0u;
    }

    // ---- warp-level multi-split: rank of each key among equal digits of its warp, stable
    uint32_t rank[kItems];
    uint32_t* wh = sm.warp_hist[warp];
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const bool ok = (warp_base + i * 32 + lane) < count;
        const uint32_t d = ok ? ((key[i] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t below = __popc(peers & lanemask_lt());
        const uint32_t pre = wh[d];
        __syncwarp();
        if (below == 0) wh[d] = pre + __popc(peers);
        __syncwarp();
        rank[i] = pre + below;
    }
    __syncthreads();

    // ---- thread d owns digit d: scan over warps, publish the tile aggregate, look back
    uint32_t digit_count;
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const uint32_t t = sm.warp_hist[w][tid];
            sm.warp_hist[w][tid] = run;
            run += t;
        }
        digit_count = run;
    }
    uint32_t* lb = lookback + (size_t)tile * kRadix;
    st_relaxed_u32(&lb[tid], (tile == 0 ? kLbPrefix : kLbAggregate) | digit_count);

    // exclusive scan over digits: tile-local start of each digit, and global digit start
    uint32_t gcount = ghist[tid];
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
            sm.scan_tmp[warp] = a;
            sm.digit_base[warp] = b;  // temporary use
        }
        __syncthreads();
        uint32_t wa = 0, wb = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            if (w < (int)warp) {
                wa += sm.scan_tmp[w];
                wb += sm.digit_base[w];
            }
        }
        local_excl = a - digit_count + wa;
        global_excl = b - gcount + wb;
        __syncthreads();
    }

    uint32_t tile_excl = 0;
    if (tile > 0) {
        int t = (int)tile - 1;
        while (true) {
            uint32_t v;
            do {
                v = ld_relaxed_u32(&lookback[(size_t)t * kRadix + tid]);
            } while ((v >> 30) == 0);
            tile_excl += v & kLbValueMask;
            if ((v >> 30) == 2 || t == 0) break;
            --t;
        }
        st_relaxed_u32(&lb[tid], kLbPrefix | (tile_excl + digit_count));
    }
    sm.digit_base[tid] = global_excl + tile_excl - local_excl;
    // warp_hist[w][d] now holds the exclusive count of digit d in warps < w; fold in the
    // tile-local digit start so a key's tile-sorted position is warp_hist[w][d] + rank.
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) sm.warp_hist[w][tid] += local_excl;
    __syncthreads();

    // ---- tile-sorted positions, then reorder through shared memory
    uint32_t pos[kItems];
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t d = (key[i] >> shift) & (kRadix - 1);
        pos[i] = wh[d] + rank[i];
    }
    __syncthreads();  // warp_hist is dead; its storage becomes the staging buffer
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const bool ok = (warp_base + i * 32 + lane) < count;
        if (ok) {
            sm.stage.keys[pos[i]] = key[i];
            sm.stage.vals[pos[i]] = val[i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kItems; i++) {
        const uint32_t j = i * kSortThreads + tid;
        if (j < tile_count) {
            const uint32_t k = sm.stage.keys[j];
            const uint32_t dst = sm.digit_base[(k >> shift) & (kRadix - 1)] + j;
            keys_out[dst] = k;
            vals_out[dst] = sm.stage.vals[j];
        }
    }
}

__global__ void copy_pairs_kernel(const uint32_t* __restrict__ k_in, const uint32_t* __restrict__ v_in, uint32_t* __restrict__ k_out,
                                  uint32_t* __restrict__ v_out, const uint32_t* __restrict__ d_count, uint32_t max_count) {
    const uint32_t count = min(*d_count, max_count);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        k_out[i] = k_in[i];
        v_out[i] = v_in[i];
    }
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
    cudaError_t e = cudaMemsetAsync(scratch.internal, 0, need, stream);
    if (e != cudaSuccess) return e;

    static bool configured = false;
    if (!configured) {
        e = cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortSmem));
        if (e != cudaSuccess) return e;
        configured = true;
    }
    uint32_t* ghist = scratch.internal;
    uint32_t* tickets = scratch.internal + kTicketOffset;
    uint32_t* lookback = scratch.internal + kLookbackOffset;

    const int hist_grid = (int)min((size_t)num_sms * 4, (tiles * kSortTile / 4 + 255) / 256);
    histogram_kernel<<<hist_grid > 0 ? hist_grid : 1, 256, 0, stream>>>(keys, d_count, max_count, begin_bit, num_passes, ghist);

    uint32_t* kin = keys;
    uint32_t* vin = payload;
    uint32_t* kout = scratch.keys_alt;
    uint32_t* vout = scratch.payload_alt;
    for (int p = 0; p < num_passes; p++) {
        onesweep_kernel<<<(unsigned)tiles, kSortThreads, sizeof(SortSmem), stream>>>(
            kin, vin, kout, vout, d_count, max_count, begin_bit + p * kRadixBits, ghist + p * kRadix, tickets + p,
            lookback + (size_t)p * tiles * kRadix);
        uint32_t* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (num_passes & 1) {  // result sits in the alt buffers: bring it home
        copy_pairs_kernel<<<num_sms * 4, 256, 0, stream>>>(kin, vin, keys, payload, d_count, max_count);
    }
    return cudaGetLastError();
}

}  // namespace sb
