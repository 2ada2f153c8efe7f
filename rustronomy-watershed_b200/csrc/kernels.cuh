// kernels.cuh -- launch wrappers of the sm_100a kernels (kernels.cu).
// All functions enqueue on `stream` and return the CUDA error of the launch.
#pragma once

#include "common.cuh"

namespace ws {

// Control block of the flood and of the label kernels (device memory, u32 words).
enum FloodCtrl {
  FC_ACTIVATIONS = 0,  // tiles taken from the worklist and iterated            (statistics: words 0..7 are
  FC_PHASES = 1,       // in-tile phases run                                      reset before every flood launch
  FC_STALE = 2,        // worklist entries dropped because nothing new had arrived  of a strip import)
  FC_IDLE = 3,         // producer warps: kilo-cycles between finding the worklist empty and the next claim
  FC_WAIT_KCYC = 4,    // consumers: kilo-cycles waiting for a staged tile (tail at the end of the flood excluded)
  FC_BUSY_KCYC = 5,    // consumers: kilo-cycles iterating
  FC_SEED_DUP = 7,     // seed_init saw a seed position twice (the statistics words 0..7 are reset per launch)
  FC_ERROR = 8,        // bit 0: seed out of bounds, bit 1: hop overflow, bit 2: orphan pixel, bit 3: slot never written,
                       // bit 4: flood watchdog, bit 5: a ring slot was overwritten while still in use
  FC_JUMP_FLAG0 = 9,   // [9..11] rotating "still unresolved" flags of the pointer jumping
  FC_JUMP_ROUNDS = 12,
  FC_STRIP_CHANGED = 13,  // a halo row of arrival times got lower on import
  FC_STRIP_PENDING = 14,  // owned pixels whose label is still a pointer
  FC_OUTSTANDING = 15,
  FC_SEED_UNSORTED = 24,  // the seed list is not strictly ascending in (slice, row, column), or holds a seed outside
                          // the image: init takes the general path (seed_init) and label_tile reads the seeds'
                          // colours from the label plane instead of deriving them from the list order    // worklist entries queued or being processed; the flood ends when it reaches 0
  FC_QAVAIL0 = 64,        // [64]  per bucket: entries fully published and not yet claimed (a semaphore)
  FC_QHEAD0 = 128,        // [64]  per bucket: next slot to hand out
  FC_QTAIL0 = 192,        // [64]  per bucket: next slot to fill
  FC_WORDS = 256
};

// Worklist of the flood: FLOOD_BUCKETS FIFO rings of tile ids, bucket = (water level of the wake-up) >> shift.
// Tiles are always taken from the lowest non-empty bucket, i.e. roughly in the order in which the
// reference's level loop would reach them, which is what keeps a tile from being iterated on values
// that a lower level is going to overwrite.
constexpr int FLOOD_BUCKETS = 63;
constexpr uint32_t Q_EMPTY = 0xFFFFFFFFu;
constexpr unsigned long long Q_DIRTY = 1ull << 63;  // qmask bit: a neighbour changed since the tile was last taken

struct FloodBuffers {
  uint32_t* T;       // [n_img][t_rows][t_pitch] arrival times, padded layout (common.cuh)
  uint8_t* pix;      // [n_img][pix_rows][pix_pitch] image bytes re-encoded for the flood (255 = never floods)
  uint32_t* lab;     // [px_total] label / parent words
  uint8_t* lvl;      // [px_total] level of colouring
  uint32_t* rim;     // [tiles_total][188] label words of the tiles' rim pixels + [2][cols] pending halo slots (labels.cu)
  uint32_t* qslots;  // [FLOOD_BUCKETS][qcap] ring buffers of tile ids, Q_EMPTY = not written yet
  unsigned long long* qmask;  // [tiles_total] bit b: an entry for the tile sits in bucket b; bit 63: Q_DIRTY
  uint32_t* ctrl;    // [FC_WORDS]
  uint32_t qcap;     // slots per ring = tiles_total + FLOOD_QSLACK
  // sorted seed lists (what find_local_minima returns): colours follow from the order, nothing is scattered
  uint32_t* row_start;        // [n_img * rows + 1] index of the first seed at or after (slice, row)
  uint32_t* rowbase;          // [n_img * rows][2 * tiles_x] index of the first seed of the row at or right of a 32-column block
  const uint32_t* seed_off;   // [n_img + 1] of the current run
  uint32_t colour_base;       // strips: colour of local seed i = colour_base + i + 1
};
constexpr uint32_t FLOOD_QSLACK = 4096;  // > CTAs that can sit between claiming a slot and clearing it

// Number of co-resident CTAs a cooperative launch of each persistent kernel may use.
int flood_max_grid(int device);
int jump_max_grid(int device);
int union_max_grid(int device);

// --- seeds (find_local_minima, lib.rs:1178-1197) ---------------------------
constexpr int MINIMA_GROUPS = 4;     // every thread handles 4 groups of 4 pixels, 1024 columns apart
constexpr int MINIMA_CHUNK = 1024 * MINIMA_GROUPS;  // columns per chunk (one CTA of 256 threads)
size_t minima_num_chunks(const ImageDims& d);
cudaError_t launch_minima_count(const uint8_t* img, ImageDims d, uint32_t* chunk_counts, cudaStream_t s);
// exclusive scan in place; writes seed_off[n_img+1] and total[0]
cudaError_t launch_minima_scan(uint32_t* chunk_counts, size_t n_chunks, ImageDims d, uint32_t* seed_off,
                               uint32_t* total, cudaStream_t s);
cudaError_t launch_minima_write(const uint8_t* img, ImageDims d, const uint32_t* chunk_offsets,
                                uint32_t* out_rc, uint32_t cap, cudaStream_t s);

// --- flood (find_flooded_px + write-back over all levels, lib.rs:196-257, 1379-1438) ---
// state of a run: arrival times, flood image, empty worklist, and -- when the seed list is sorted -- the seeds
// themselves (launch_seed_init then returns at once)
cudaError_t launch_fill_state(FloodBuffers b, ImageDims d, const uint8_t* img, uint32_t lmax, const uint32_t* seeds_rc,
                              const uint32_t* seed_off, uint32_t nseeds, int sms, cudaStream_t s);
cudaError_t launch_seed_init(FloodBuffers b, ImageDims d, const uint32_t* seeds_rc, const uint32_t* seed_off,
                             uint32_t nseeds, uint32_t colour_base, cudaStream_t s);
// row-strip decomposition: boundary rows of arrival times / labels (flood.cu)
cudaError_t launch_strip_export_T(const uint32_t* T, ImageDims d, int ra, int rb, uint32_t* out_a, uint32_t* out_b,
                                  cudaStream_t s);
cudaError_t launch_strip_import_T(FloodBuffers b, ImageDims d, int row, int nb_row, const uint32_t* in,
                                  int bucket_shift, cudaStream_t s);
// (a word that is still a reference is looked up in the rim array: the label plane may be unfinished)
cudaError_t launch_strip_export_lab(const uint32_t* lab, const uint32_t* rim, ImageDims d, int ra, int rb,
                                    uint32_t* out_a, uint32_t* out_b, cudaStream_t s);
// unresolved entries of the rim array incl. the pending halo slots, added to ctrl[FC_STRIP_PENDING]
cudaError_t launch_strip_count_pending_rim(const uint32_t* rim, size_t n, uint32_t* ctrl, int sms, cudaStream_t s);
cudaError_t launch_strip_import_lab(FloodBuffers b, ImageDims d, int row, const uint32_t* in, cudaStream_t s);
cudaError_t launch_strip_count_pending(const uint32_t* lab, ImageDims d, int r0, int r1, uint32_t* ctrl,
                                       cudaStream_t s);
cudaError_t launch_seeds_convert(const uint64_t* in, uint32_t* out, size_t nseeds, size_t rows, size_t cols,
                                 cudaStream_t s);
// bucket_shift: worklist bucket = wake-up level >> shift (flood_bucket_shift() picks it per run)
// tensor_maps: the two 128-byte maps made by flood_make_tensor_maps for this plan's planes
constexpr size_t FLOOD_TENSOR_MAP_BYTES = 256;
cudaError_t flood_make_tensor_maps(const FloodBuffers& b, const ImageDims& d, void* out_maps);
cudaError_t launch_flood(FloodBuffers b, ImageDims d, int check_overflow, int bucket_shift, int grid,
                         const void* tensor_maps, cudaStream_t s);
int flood_bucket_shift(size_t nseeds, const ImageDims& d);
cudaError_t launch_unpad_T(const uint32_t* Tp, ImageDims d, uint32_t* out, cudaStream_t s);

// --- labels (colour decision of lib.rs:235-255 with the `col0` tie-break) ---
size_t rim_words(const ImageDims& d);
// also counts, per slice, the owned pixels that hold a seed = colours present on the canvas (ndistinct[n_img])
// tie_random: draw the parent uniformly among the earlier neighbours (lib.rs:250-253) instead of taking the first
cudaError_t launch_parent(FloodBuffers b, ImageDims d, uint32_t* ndistinct, bool tie_random, uint64_t tie_seed,
                          const void* tensor_maps,
                          cudaStream_t s);
cudaError_t launch_jump(FloodBuffers b, ImageDims d, int grid, int finish, cudaStream_t s);
cudaError_t launch_label_finish(FloodBuffers b, ImageDims d, int sms, cudaStream_t s);

// --- merging (find_merge + make_colour_map + recolour, lib.rs:393-542, 590-592) ---
struct MergeBuffers {
  uint32_t* level_hist;   // [256] edges per level, then exclusive offsets [257]
  uint32_t* level_cursor; // [256]
  uint2* red_ab;          // [cap] forest edges of the tiles (global colour ids), unsorted
  uint8_t* red_w;         // [cap] their levels
  uint32_t* red_count;    // [16]: [0] edges emitted, [12] tiles on ovf_list, the rest statistics
  uint32_t* ovf_list;     // [tiles_total] tiles that did not fit merge_reduce's small size
  uint2* edges;           // [cap] the same edges bucketed by level
  uint32_t* parent;       // [nseeds] union-find with path halving
  uint32_t* hook_to;      // [nseeds] immutable link written once when a root is hooked
  uint8_t* hook_lvl;      // [nseeds] level of that link, 255 = still a root
  uint32_t* unions;       // [n_img][256] successful unions per level
  uint32_t* fin_hist;     // [n_img][256] FINAL forest edges per level (contracted in their tile, never unioned globally)
  uint32_t* ndistinct;    // [n_img] colours present on the canvas (counted by label_tile)
  uint32_t* counts;       // [n_img][256] lakes per level
};
// per-tile contraction + spanning-forest reduction in shared memory (merge.cu).  Edges: .x / .y = global
// colour ids, bit 31 of .y = FINAL (a certain forest edge, only to be counted).  contract = 0: no FINAL edges.
size_t merge_reduce_capacity(const ImageDims& d);
// ovf_list [tiles_total]: scratch for the tiles that need the full-size kernel
cudaError_t launch_merge_reduce(const uint32_t* lab, const uint8_t* lvl, ImageDims d, const uint32_t* seed_off,
                                int contract, uint2* red_ab, uint8_t* red_w, uint32_t* red_count, uint32_t* ovf_list,
                                cudaStream_t s);
// counting sort by level: level_hist[0..256] = exclusive offsets, edges = buckets.  all = 0: only DEFERRED
// edges are bucketed, FINAL ones are counted into fin_hist[slice][level]; all = 1: every edge is bucketed.
cudaError_t launch_red_sort(const uint2* red_ab, const uint8_t* red_w, const uint32_t* red_count, int all,
                            const uint32_t* seed_off, int n_img, uint32_t* level_hist, uint32_t* level_cursor,
                            uint32_t* fin_hist, uint2* edges, cudaStream_t s);
cudaError_t launch_uf_init(MergeBuffers m, ImageDims d, uint32_t nseeds, cudaStream_t s);
cudaError_t launch_uf_reset(MergeBuffers m, uint32_t n, cudaStream_t s);
cudaError_t launch_count_present(const uint32_t* lab, ImageDims d, const uint32_t* seeds_rc, uint32_t nseeds,
                                 uint32_t colour_base, uint32_t* out, cudaStream_t s);
cudaError_t launch_union_levels(MergeBuffers m, const uint32_t* seed_off, int n_img, uint32_t lmax, int grid,
                                cudaStream_t s);
cudaError_t launch_lake_counts(MergeBuffers m, int n_img, uint32_t lmax, cudaStream_t s);
// rep[g] = representative of colour g at `level` (start from rep if `incremental`)
cudaError_t launch_rep_table(const uint32_t* hook_to, const uint8_t* hook_lvl, uint32_t nseeds, uint32_t level,
                             int incremental, uint32_t* rep, cudaStream_t s);

// --- minimum spanning forest of the DEFERRED graph by Boruvka rounds (forest.cu) ---
struct ForestBuffers {
  uint32_t* parent;            // [ncolours] hook pointers (written in the hook phase only)
  uint32_t* link;              // [ncolours] identity of a closed node that moved along a FINAL edge
  unsigned long long* best;    // [ncolours] lightest offer: (255 - round) << 56 | level << 32 | position
  uint8_t* open_;              // [ncolours] 1: the node has edges this list does not hold (strip boundary basins)
  uint2* ab[2];                // live edges as the current roots of their ends, alternating by round
  uint8_t* w[2];
  uint2* orig[2];              // strips: the edges' own ends
  uint32_t* count;             // [2] live edges in ab[0] / ab[1]
  uint32_t* n_deferred;        // DEFERRED picks (strips only)
  uint32_t* rounds;
  uint32_t* error;             // bit 0: round limit, bit 1: DEFERRED list full, bit 2: identity chain too long
  uint2* def_ab;               // [def_cap] DEFERRED picks, between identities after forest_ident
  uint8_t* def_w;
  uint32_t def_cap;
  uint32_t* tile_hist;         // [n_img][256] FINAL edges of the tiles per level
  uint32_t* forest_hist;       // [n_img][256] FINAL picks of the rounds per level
};
int forest_max_grid(int device);
// with_open: the run marks open nodes (strips); else every node is closed
cudaError_t launch_forest_init(ForestBuffers f, uint32_t ncolours, int n_img, int with_open, int sms, cudaStream_t s);
// open_[colour - 1] = 1 for the colours of `count` label words
cudaError_t launch_forest_mark_open(const uint32_t* lab, size_t count, uint32_t ncolours, uint8_t* open_,
                                    cudaStream_t s);
// in_*: the edge list to start from; skip_final: entries with bit 31 of .y are FINAL tile edges (counted only)
cudaError_t launch_forest(ForestBuffers f, const uint2* in_ab, const uint8_t* in_w, const uint32_t* in_count,
                          int skip_final, int with_open, const uint32_t* seed_off, int n_img, int grid, int sms,
                          cudaStream_t s);
// strips: *dst |= (*src != 0) (as_flag) or *dst += *src
cudaError_t launch_ctrl_accumulate(const uint32_t* src, uint32_t* dst, int as_flag, cudaStream_t s);
// strips: header + DEFERRED picks of a finished forest run -> one packet; packets -> one edge list + summed header
cudaError_t launch_strip_packet(ForestBuffers f, const uint32_t* ndistinct, uint32_t cap, void* packet, cudaStream_t s);
cudaError_t launch_strip_unpack(const void* packets, uint32_t n_packets, uint32_t cap, ForestBuffers f,
                                uint32_t* out_count, uint32_t* ndistinct, cudaStream_t s);
cudaError_t launch_forest_lake_counts(const uint32_t* ndistinct, ForestBuffers f, int n_img, uint32_t lmax,
                                      uint32_t* counts, cudaStream_t s);

// --- per-level outputs (hooks of transform_history / transform_to_list) -----
// out[p] = lvl[p] <= level ? label : 0, label optionally mapped through rep[] (merging)
cudaError_t launch_snapshot(const uint32_t* lab, const uint8_t* lvl, size_t n_px, uint32_t level,
                            const uint32_t* rep, uint32_t colour_base, uint64_t* out, cudaStream_t s);
cudaError_t launch_snapshot32(const uint32_t* lab, const uint8_t* lvl, size_t n_px, uint32_t level,
                              const uint32_t* rep, uint32_t colour_base, uint32_t* out, cudaStream_t s);
cudaError_t launch_widen_labels(const uint32_t* lab, size_t n_px, uint64_t* out, cudaStream_t s);
cudaError_t launch_strip_labels(const uint32_t* lab, size_t n_px, uint32_t* out, cudaStream_t s);
// lvl_hist[b][256]: pixels coloured at each level (255 = never)
cudaError_t launch_level_hist(const uint8_t* lvl, ImageDims d, uint32_t* lvl_hist, cudaStream_t s);
// cnt[l][c] (row length ncol = nseeds+1): pixels of colour c coloured at level l   (single slice)
cudaError_t launch_colour_level_count(const uint32_t* lab, const uint8_t* lvl, size_t n_px, uint32_t ncol,
                                      uint32_t* cnt, cudaStream_t s);
// sizes[l][c] = sum_{l' <= l} cnt[l'][c] for c >= 1, sizes[l][0] = n_px - sum_c  (find_lake_sizes rows)
cudaError_t launch_sizes_cumulate(const uint32_t* cnt, uint32_t ncol, uint32_t nlevels, size_t n_px,
                                  uint64_t* sizes, cudaStream_t s);
// merging: fold the row of level l through rep[] in place (sizes[l][rep(c)] += sizes[l][c])
cudaError_t launch_sizes_fold(uint64_t* sizes_row, const uint32_t* rep, uint32_t ncol, uint64_t* scratch_row,
                              cudaStream_t s);
// MergingWatershed::transform (lib.rs:1524-1536): interior 123, border 0
cudaError_t launch_fill_const123(uint64_t* out, int rows, int cols, cudaStream_t s);
// edge correction (lib.rs:1340-1356): copy into a zeroed canvas one pixel larger on every side
cudaError_t launch_pad_image(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s);

// pre_processor (lib.rs:1081-1173): dtype 0 f32, 1 f64, 2 i32, 3 u16, 4 i16, 5 u8, 6 i64
size_t pre_processor_scratch_bytes();
cudaError_t launch_pre_processor(int dtype, const void* in, size_t n, uint32_t maxv, void* scratch, double* minmax,
                                 uint8_t* out, cudaStream_t s);

}  // namespace ws
