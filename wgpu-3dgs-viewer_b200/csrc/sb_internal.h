// sb_internal.h — launcher interfaces between the C-ABI layer (sb_api.cu) and the kernels.
#pragma once

#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "sb_common.cuh"

namespace sb {

// ---------------------------------------------------------------- preprocess (sb_preprocess.cu)
struct PreParams {
    const uint8_t* gaussians;          // packed pods
    uint32_t n;
    uint32_t num_tiles;                // ceil(n / records-per-tile)
    const uint32_t* selection;         // nullptr = selection feature off
    uint32_t invert_selection;
    uint32_t* indices;                 // IndirectIndicesBuffer
    float* keys;                       // GaussiansDepthBuffer (first ceil(n/3840)*3840 floats used)
    uint32_t keys_capacity;            // ceil(n/3840)*3840
    SbDrawIndirectArgs* draw_args;     // IndirectArgsBuffer
    SbDispatchIndirectArgs* sort_args; // RadixSortIndirectArgsBuffer
    SplatRec* recs;                    // per-Gaussian projected record (indexed by Gaussian index)
    TileBox* tboxes;                   // per-Gaussian tile bbox
    uint32_t* tile_counter;            // dynamic tile ticket (zeroed before launch)
    unsigned long long* tile_status;   // decoupled look-back state (zeroed before launch)
    uint32_t* visible_count;           // V for downstream kernels
    uint32_t* visible_host;            // nullable: device alias of a mapped pinned word that also receives V
    uint32_t index_base;               // gaussians / recs / tboxes point at Gaussian `index_base` of the model (a slice is preprocessed):
                                       // indices written and selection bits looked up are index_base + the slice-local index
    uint32_t strip_on;                 // 1: keep only the visible splats whose tile box meets tile rows [strip_ty_lo, strip_ty_hi]
    uint32_t strip_ty_lo, strip_ty_hi; //    (one frame split into screen strips over several GPUs; needs recs)
    uint32_t* sort_prep;               // nullable: [kSortPrepOr] |= key, [kSortPrepNand] |= ~key over the visible keys (zeroed before launch)
    Uniforms u;
};

int preprocess_records_per_tile(int sh_fmt, int cov_fmt);
// scratch bytes needed for tile_counter + tile_status
size_t preprocess_scratch_bytes(uint32_t n, int sh_fmt, int cov_fmt);
// When PreParams.sort_prep is set it must point at the START of `scratch`, whose first preprocess_scratch_prefix_bytes() bytes
// are the depth sort's prep words; K1's own words follow (scratch_bytes = prefix + preprocess_scratch_bytes).
size_t preprocess_scratch_prefix_bytes();
cudaError_t launch_preprocess(int sh_fmt, int cov_fmt, PreParams& p, void* scratch, size_t scratch_bytes,
                              int num_sms, cudaStream_t stream);

// Vertex stage alone (standalone Renderer on caller-owned buffers): recs/tboxes of the first *count Gaussians named by
// `indices`; p needs gaussians, n, recs, tboxes and u only.
cudaError_t launch_vertex_stage(int sh_fmt, int cov_fmt, PreParams& p, const uint32_t* indices, const uint32_t* count, int num_sms,
                                cudaStream_t stream);

// ---------------------------------------------------------------- radix sort (sb_sort.cu)
struct SortScratch {
    uint32_t* keys_alt;
    uint32_t* payload_alt;
    uint32_t* internal;      // histograms + look-back tables + tickets
    size_t internal_bytes;
};
size_t sort_internal_bytes(uint32_t capacity);
// Stable ascending LSD sort of the first *d_count (clamped to max_count) (key,payload) pairs over
// bits [begin_bit, end_bit).  Result lands in keys/payload when the pass count is even, and is
// copied back otherwise.
// parity_out (device word, nullable): the result is left where the last pass wrote it — *parity_out = 1: in the alt
// buffers — and launch_sort_finish brings it home when somebody needs it there.
// expected_count (0 = unknown): a hint for the choice of pass kernel — small sorts are bound by the length of the look-back
// chain and take the kernel with the larger tiles.
cudaError_t launch_sort(uint32_t* keys, uint32_t* payload, const uint32_t* d_count, uint32_t max_count,
                        int begin_bit, int end_bit, const SortScratch& scratch, int num_sms, cudaStream_t stream,
                        uint32_t* parity_out = nullptr, uint32_t expected_count = 0);
// Key-adaptive sort (v4).  `prep` (kSortPrepWords u32, nullable): [kSortPrepOr] / [kSortPrepNand] hold the OR of the keys and the
// OR of their complements (K1 accumulates them while it writes the keys); everything after them must be zero when prep_zeroed.
// begin_bit >= 0: sort exactly bits [begin_bit, end_bit) and ignore or / nand.  The result is left where the last pass wrote it;
// *parity_out (device) = 1: in scratch.keys_alt / payload_alt (launch_sort_finish brings it home).
constexpr int kSortPrepOr = 0, kSortPrepNand = 1, kSortPrepPasses = 2, kSortPrepParity = 3, kSortPrepPlan = 4, kSortPrepTickets = 8,
              kSortPrepDone = 12, kSortPrepHist = 16, kSortPrepWords = 16 + 4 * 512;
size_t sort_prep_bytes();
cudaError_t launch_sort_adaptive(uint32_t* keys, uint32_t* payload, const uint32_t* d_count, uint32_t max_count, uint32_t* prep,
                                 bool prep_zeroed, int begin_bit, int end_bit, const SortScratch& scratch, int num_sms,
                                 cudaStream_t stream, uint32_t* parity_out);
cudaError_t launch_sort_finish(uint32_t* keys, uint32_t* payload, const SortScratch& scratch, const uint32_t* d_count, uint32_t max_count,
                               const uint32_t* parity, int num_sms, cudaStream_t stream);

// ---------------------------------------------------------------- binning + raster (sb_raster.cu)
struct RasterBuffers {
    uint32_t* dup_offsets;    // [n+1] exclusive scan of tiles-per-splat in sorted order
    TileBox* tboxes_sorted;   // [n] tile boxes in depth order (nullable: the v1 scan / emit pair gathers them twice instead)
    uint32_t* win_first;      // [win_capacity] first splat (depth rank) of every 4096-duplicate emit window
    uint32_t win_capacity;
    uint32_t* dup_keys;       // [dup_capacity] tile id
    uint32_t* dup_vals;       // [dup_capacity] Gaussian index (in depth order within a tile)
    uint32_t* tile_ranges;    // [tiles*2] begin,end into dup arrays
    uint32_t* tile_order;     // [tiles] raster schedule: tiles by descending list length (nullable: row-major)
    uint32_t* dup_count;      // D (device)
    uint32_t* overflow;       // flag
    uint32_t* needed_host;    // device alias of a mapped pinned word (may be null)
    unsigned long long* scan_status;
    uint32_t* scan_counter;
    SplatRec* tile_recs;      // [dup_capacity] records gathered in tile order (TMA staging source)
    uint64_t dup_capacity;
};

struct RasterParams {
    const SplatRec* recs;            // indexed by Gaussian index
    const TileBox* tboxes;
    const uint32_t* sorted_indices;  // depth order
    const uint32_t* sorted_indices_alt;  // where the depth sort left them when *sort_parity != 0 (both nullable)
    const uint32_t* sort_parity;
    const uint32_t* visible_count;
    uint32_t max_visible;            // n
    RasterBuffers buf;
    SortScratch sort;                // for the tile sort
    Uniforms u;
    SbTarget target;
    int strict_exp;
    int clear;                       // 1: clear to BLACK first (first model of a frame)
    int obb_cull;                    // 1: warp-level cull also tests the ellipse axes (SAT), 0: bbox only
    float cut_k;                     // 1 / kAlphaCut when recs carry cut extents (the cull may then use each splat's cut radius), else 0
    int no_discard;                  // allow the rasterizer to drop the discard test where it is provably redundant (exact cut-off on)
    int raster_bands;                // rasterize the tile rows as this many launches (<= 1: one)
    float* depth;                    // optional f32 depth attachment (strip geometry of the target); needs recs_map
    uint32_t depth_pitch;
    int depth_compare, depth_write;
    const uint8_t* pods;             // pod array + stride: the depth-tested pass recomputes each splat's ndc z
    uint32_t pod_stride;
    const CUtensorMap* recs_map;     // non-null: fetch records with TMA gather4 (no gathered copy); host pointer, passed by value
    unsigned long long* counters;    // optional instrumentation: [0] alive fragments, [1] evaluated lane pairs
    cudaEvent_t* events;             // optional: [0]=after scan+emit, [1]=after tile sort, [2]=after gather, [3]=after raster
    // Optional split (multi-model frames): binning runs on the stream passed to launch_bin_and_raster, `bin_done` is recorded
    // there, and the raster kernel is enqueued on `raster_stream` after waiting for it — models are binned concurrently while
    // their rasters stay in draw order on one stream.
    bool split_raster;               // (a null raster_stream is the legacy default stream, so the split has its own flag)
    cudaStream_t raster_stream;
    cudaEvent_t bin_done;
};
cudaError_t launch_bin_and_raster(const RasterParams& p, int num_sms, cudaStream_t stream);
cudaError_t launch_clear(const SbTarget& target, cudaStream_t stream);

// ---------------------------------------------------------------- partitioned strips (sb_strips.cu)
constexpr int kMaxStripRanks = 16;
struct StripScatterParams {
    const uint32_t* indices;        // this rank's visible list (global Gaussian indices, ascending) and keys
    const float* keys;
    const uint32_t* visible_count;
    uint32_t max_visible;           // slice size
    const SplatRec* recs;           // this rank's arrays, indexed by global Gaussian index
    const TileBox* tboxes;
    uint32_t world, rank;
    uint32_t segment_capacity;      // entries per (destination, source) inbox segment
    uint32_t ty_lo[kMaxStripRanks], ty_hi[kMaxStripRanks];  // tile rows of every rank's strip (lo > hi: no strip)
    // A parcel is 64 bytes: {index, key bits, tile box min, max} + the 48-byte record.  Segments hold segment_capacity parcels.
    uint4* inbox;                            // this rank's inbox: [world] segments (segment r = from rank r)
    uint32_t* inbox_counts;                  //   and their [world] counts
    uint4* outbox;                           // this rank's outbox: [world] segments (segment d = for rank d), local
    uint32_t* outbox_counts;
    uint4* peer_inbox[kMaxStripRanks];       // destination d's inbox and counts, peer-mapped
    uint32_t* peer_inbox_counts[kMaxStripRanks];
    uint32_t* tickets;              // set by the launcher
    unsigned long long* status;
    uint32_t status_stride;
};
struct StripConcatParams {
    const uint4* inbox;             // this rank's inbox (64-byte parcels) and segment counts
    const uint32_t* inbox_counts;
    uint32_t rank;
    SplatRec* recs;                 // this rank's arrays: records / tile boxes of splats other ranks preprocessed are unpacked here
    TileBox* tboxes;
    uint32_t world, segment_capacity, max_visible;
    uint32_t* indices;
    float* keys;
    uint32_t keys_capacity;
    SbDrawIndirectArgs* draw_args;
    SbDispatchIndirectArgs* sort_args;
    uint32_t* visible_count;
    uint32_t* visible_host;
    uint32_t* sort_prep;            // zeroed by the caller; receives or / nand of the strip's keys
};
size_t strip_scatter_scratch_bytes(uint32_t max_visible, uint32_t world);
cudaError_t launch_strip_scatter(StripScatterParams& p, void* scratch, size_t scratch_bytes, cudaStream_t stream);
cudaError_t launch_strip_concat(const StripConcatParams& p, int num_sms, cudaStream_t stream);

// ---------------------------------------------------------------- viewport selection (sb_select.cu)
// selection::viewport::main with an analytic rectangle mask; writes ceil(n/32) words.
cudaError_t launch_select_rect(const uint8_t* gaussians, uint32_t n, uint32_t stride, const Uniforms& u, float x0, float y0, float x1,
                               float y1, uint32_t* words, cudaStream_t stream);

// selection::viewport::main with an analytic brush mask: texels within `radius` of the stroke polyline
// (n_points <= SB_BRUSH_MAX_POINTS host floats x,y); accumulate != 0 ORs into the existing words.
cudaError_t launch_select_brush(const uint8_t* gaussians, uint32_t n, uint32_t stride, const Uniforms& u, const float* points_xy,
                                uint32_t n_points, float radius, int accumulate, uint32_t* words, cudaStream_t stream);

// editor-style colour override: snapshot of the colour words, then selected -> override / others -> source
cudaError_t launch_snapshot_colors(const uint8_t* gaussians, uint32_t n, uint32_t stride, uint32_t* orig, cudaStream_t stream);
cudaError_t launch_rgb_override(uint8_t* gaussians, uint32_t n, uint32_t stride, const uint32_t* orig, const uint32_t* selection,
                                const float rgb[3], float alpha, cudaStream_t stream);

// editor BasicColorModifiers on the selected Gaussians (rgb override or HSV, then contrast / exposure / gamma, alpha scale)
cudaError_t launch_basic_color_modifiers(uint8_t* gaussians, uint32_t n, uint32_t stride, const uint32_t* orig, const uint32_t* selection,
                                         int rgb_override, const float rgb_or_hsv[3], float alpha, float contrast, float exposure,
                                         float gamma, cudaStream_t stream);

// ---------------------------------------------------------------- measured peaks (sb_probe.cu)
cudaError_t probe_fp32_peak(int num_sms, cudaStream_t stream, double* lane_ops_per_s);
cudaError_t probe_smem_peak(int num_sms, cudaStream_t stream, double* bytes_per_s);

}  // namespace sb
