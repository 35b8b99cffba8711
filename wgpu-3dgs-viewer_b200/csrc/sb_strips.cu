// sb_strips.cu — one frame sharded over the GPUs of a node as screen strips, with the Preprocessor's work PARTITIONED too.
//
// Rendering a strip costs 1/G of the sort, binning and rasterizer work, but a rank that culls the whole scene for its strip still
// streams every pod (K1: 0.31 ms for 6 M Gaussians whatever survives) — at G = 8 a third of the frame.  Here rank r preprocesses
// only Gaussians [n r / G, n (r + 1) / G) (the ordinary fused K1 on its slice: the reference's full-frame cull, bit for bit) and
// hands every rank the splats that rank's strip needs, straight into that rank's memory over NVLink (peer-mapped buffers):
//
//   strip_scatter_kernel   grid (tiles of the slice's visible list, G destinations).  Block (t, d) keeps the entries whose tile box
//                          meets strip d — order-preserving compaction: ballots, block scan, decoupled look-back per destination —
//                          and PACKS them as 64-byte parcels {index, key, tile box, 48-byte record} into a local outbox segment
//                          (its own strip's parcels go straight into its own inbox).  The last tile leaves the segment's count.
//   strip_send_kernel      one coalesced 128-bit copy of every outbox segment into segment r of the destination's inbox: NVLink
//                          sees full-line stores (scattering 16-byte pieces straight into the peers' arrays measured 0.5 ms at
//                          two ranks — small remote writes waste the link).
//   [one barrier across the ranks: every peer store above has landed]
//   strip_concat_kernel    rank d strings its G inbox segments together in source order and unpacks the records / tile boxes into
//                          its own recs[index] / tboxes[index].  Slices are ascending index ranges and each segment is in
//                          ascending index order, so the result is exactly the strip's visible list as a single-GPU cull would
//                          have compacted it; it also rebuilds what K1 leaves for the stages behind it (indirect args, visible
//                          count, pad keys, the depth sort's varying-bit masks).
//
// From there the rank runs the unchanged depth sort, binning and rasterizer on its own list.  The reference has no multi-GPU
// code; the unit being sharded is Viewer::render (src/lib.rs:266-275).
#include "sb_internal.h"

namespace sb {

namespace {

constexpr int kXThreads = 512;
constexpr int kXItems = 4;
constexpr int kXTile = kXThreads * kXItems;  // 2048 list entries per block

__global__ void __launch_bounds__(kXThreads) strip_scatter_kernel(const __grid_constant__ StripScatterParams p) {
    constexpr int NW = kXThreads / 32;
    __shared__ uint32_t counts[kXItems * NW];
    __shared__ uint32_t s_tile, s_base, s_total;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t d = blockIdx.y;  // destination rank
    const uint32_t v = min(*p.visible_count, p.max_visible);
    if (blockIdx.x * kXTile >= v && blockIdx.x > 0) return;  // exactly max(1, ceil(v / tile)) blocks per destination take a ticket
    if (tid == 0) s_tile = atomicAdd(&p.tickets[d], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t ylo = p.ty_lo[d], yhi = p.ty_hi[d];
    const bool has_strip = ylo <= yhi;  // a rank without a strip (more ranks than tile rows) is encoded as lo > hi

    uint32_t gs[kXItems];
    float ks[kXItems];
    uint32_t hit_bits = 0, ranks = 0;
#pragma unroll
    for (int i = 0; i < kXItems; i++) {
        const uint32_t e = tile * kXTile + i * kXThreads + tid;  // ascending (item, warp, lane) = ascending list position
        bool hit = e < v;
        gs[i] = hit ? p.indices[e] : 0u;
        ks[i] = hit ? p.keys[e] : 0.0f;
        if (hit) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(&p.tboxes[gs[i]]));
            const uint32_t x0 = q.x & 0xffffu, y0 = q.x >> 16, x1 = q.y & 0xffffu, y1 = q.y >> 16;
            hit = has_strip && x0 <= x1 && y0 <= y1 && y1 >= ylo && y0 <= yhi;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) counts[i * NW + warp] = __popc(bal);
        hit_bits |= (hit ? 1u : 0u) << i;
        ranks |= (uint32_t)__popc(bal & lanemask_lt()) << (8 * i);
    }
    __syncthreads();
    static_assert(kXItems * NW == 64, "two (item, warp) counts per lane of warp 0");
    if (warp == 0) {
        const uint32_t a = counts[2 * lane], b = counts[2 * lane + 1];
        uint32_t inc = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        counts[2 * lane] = inc - a - b;
        counts[2 * lane + 1] = inc - b;
        const uint32_t excl = lookback(p.status + (size_t)d * p.status_stride, tile, total, lane);
        if (lane == 0) {
            s_base = excl;
            s_total = total;
        }
    }
    __syncthreads();
    const uint32_t base = s_base;
    // parcels for a remote strip are packed locally and shipped in bulk; this rank's own strip goes straight into its own inbox
    uint4* __restrict__ out = (d == p.rank ? p.inbox : p.outbox) + ((size_t)(d == p.rank ? p.rank : d) * p.segment_capacity) * 4;
#pragma unroll
    for (int i = 0; i < kXItems; i++) {
        if ((hit_bits >> i) & 1u) {
            const uint32_t g = gs[i];
            const uint32_t slot = base + counts[i * NW + warp] + ((ranks >> (8 * i)) & 0xffu);
            if (slot < p.segment_capacity) {
                const uint2 tb = *reinterpret_cast<const uint2*>(&p.tboxes[g]);
                const uint4* src = reinterpret_cast<const uint4*>(&p.recs[g]);
                uint4* dst = out + (size_t)slot * 4;
                dst[0] = make_uint4(g, __float_as_uint(ks[i]), tb.x, tb.y);
                dst[1] = src[0];
                dst[2] = src[1];
                dst[3] = src[2];
            }
        }
    }
    const uint32_t last_tile = v == 0 ? 0 : (v - 1) / kXTile;
    if (tile == last_tile && tid == 0) {
        const uint32_t c = min(base + s_total, p.segment_capacity);
        if (d == p.rank) p.inbox_counts[p.rank] = c;
        else p.outbox_counts[d] = c;
    }
}

// outbox segment d -> segment `rank` of destination d's inbox, 128-bit coalesced; block (x, d)
__global__ void __launch_bounds__(256) strip_send_kernel(const __grid_constant__ StripScatterParams p) {
    const uint32_t d = blockIdx.y;
    if (d == p.rank) return;
    const uint32_t c = p.outbox_counts[d];
    const uint4* __restrict__ src = p.outbox + ((size_t)d * p.segment_capacity) * 4;
    uint4* __restrict__ dst = p.peer_inbox[d] + ((size_t)p.rank * p.segment_capacity) * 4;
    const size_t nvec = (size_t)c * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) p.peer_inbox_counts[d][p.rank] = c;
}

__global__ void __launch_bounds__(512) strip_concat_kernel(const __grid_constant__ StripConcatParams p) {
    __shared__ uint32_t s_prefix[kMaxStripRanks + 1];
    __shared__ uint32_t s_or, s_nand;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    if (tid == 0) {
        uint32_t acc = 0;
        for (uint32_t r = 0; r < p.world; r++) {
            s_prefix[r] = acc;
            acc += min(p.inbox_counts[r], p.segment_capacity);
        }
        s_prefix[p.world] = acc;
        s_or = s_nand = 0u;
    }
    __syncthreads();
    const uint32_t v = min(s_prefix[p.world], p.max_visible);
    uint32_t key_or = 0u, key_nand = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + tid; i < v; i += gridDim.x * blockDim.x) {
        uint32_t r = 0;
        while (r + 1 < p.world && i >= s_prefix[r + 1]) ++r;
        const uint4* parcel = p.inbox + ((size_t)r * p.segment_capacity + (i - s_prefix[r])) * 4;
        const uint4 h = parcel[0];  // index, key bits, tile box
        p.indices[i] = h.x;
        p.keys[i] = __uint_as_float(h.y);
        key_or |= h.y;
        key_nand |= ~h.y;
        if (r != p.rank) {  // a splat another rank preprocessed: its record and tile box, by Gaussian index
            uint4* rec = reinterpret_cast<uint4*>(&p.recs[h.x]);
            rec[0] = parcel[1];
            rec[1] = parcel[2];
            rec[2] = parcel[3];
            *reinterpret_cast<uint2*>(&p.tboxes[h.x]) = make_uint2(h.z, h.w);
        }
    }
    if (p.sort_prep != nullptr) {
        key_or = __reduce_or_sync(0xffffffffu, key_or);
        key_nand = __reduce_or_sync(0xffffffffu, key_nand);
        if (lane == 0) {
            if (key_or) atomicOr(&s_or, key_or);
            if (key_nand) atomicOr(&s_nand, key_nand);
        }
        __syncthreads();
        if (tid == 0) {
            if (s_or) atomicOr(&p.sort_prep[kSortPrepOr], s_or);
            if (s_nand) atomicOr(&p.sort_prep[kSortPrepNand], s_nand);
        }
    }
    if (blockIdx.x == 0) {
        // what K1's `post` leaves (preprocess.wesl:108-126), now for the strip's own visible set
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
        for (uint32_t i = v + tid; i < padded; i += blockDim.x) p.keys[i] = 2.0f;
    }
}

}  // namespace

size_t strip_scatter_scratch_bytes(uint32_t max_visible, uint32_t world) {
    const size_t tiles = ((size_t)max_visible + kXTile - 1) / kXTile + 1;
    return 64 + (size_t)world * tiles * sizeof(unsigned long long);
}

cudaError_t launch_strip_scatter(StripScatterParams& p, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
    const size_t tiles = ((size_t)p.max_visible + kXTile - 1) / kXTile + 1;
    if (p.world > (uint32_t)kMaxStripRanks || scratch_bytes < 64 + (size_t)p.world * tiles * sizeof(unsigned long long)) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(scratch, 0, scratch_bytes, stream);  // tickets + look-back status of every destination
    if (e != cudaSuccess) return e;
    p.tickets = reinterpret_cast<uint32_t*>(scratch);
    p.status = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(scratch) + 64);
    p.status_stride = (uint32_t)tiles;
    const dim3 grid((unsigned)(tiles - 1 > 0 ? tiles - 1 : 1), p.world);
    strip_scatter_kernel<<<grid, kXThreads, 0, stream>>>(p);
    if (p.world > 1) strip_send_kernel<<<dim3(64, p.world), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_strip_concat(const StripConcatParams& p, int num_sms, cudaStream_t stream) {
    if (p.world > (uint32_t)kMaxStripRanks) return cudaErrorInvalidValue;
    strip_concat_kernel<<<num_sms * 2, 512, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sb
