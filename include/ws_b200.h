/*
 * ws_b200.h -- C ABI of the B200-native watershed engine.
 *
 * This is the drop-in boundary for the hot path of smups/rustronomy-watershed
 * v0.4.1 (all citations are into the reference's src/lib.rs).  The reference is
 * a pure-Rust crate with no FFI of its own; the entry points below are what a
 * `extern "C"` block in a shim crate binds so that `TransformBuilder`,
 * `Watershed::{transform, transform_with_hook, transform_to_list,
 * transform_history}` and `WatershedUtils::find_local_minima` keep their
 * signatures while the work runs as hand-written sm_100a CUDA
 * (INTEGRATION.md shows that shim).
 *
 * Conventions
 *   - plain pointers and sizes only; images are u8, axis 0 = row ("x" in the
 *     reference), strides are in ELEMENTS and may be negative (ArrayView2);
 *   - seeds are `[n][2]` uint64 = (row, col), the repacked `&[(usize, usize)]`;
 *   - labels ("colours") are uint64 like the reference's usize, 0 = UNCOLOURED
 *     (lib.rs:138), colour of seed i = i + 1 (lib.rs:1360-1367);
 *   - every function returns a ws_status; nothing throws, nothing aborts.  The
 *     reference's run-time failures are panics (e.g. out-of-bounds seed,
 *     lib.rs:1366); the shim turns a non-zero status back into a panic;
 *   - there is NO CPU fallback: without a CUDA device ws_ctx_create fails;
 *   - a ws_ctx is used by one thread at a time (the shim keeps one per thread);
 *     per-level hooks run on the calling thread, in level order.
 */
#ifndef WS_B200_H
#define WS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WS_ABI_VERSION 1

/* lib.rs:138-141 */
#define WS_UNCOLOURED 0u
#define WS_NORMAL_MAX 254u
#define WS_ALWAYS_FILL 0u
#define WS_NEVER_FILL 255u

typedef enum ws_status {
  WS_OK = 0,
  WS_ERR_INVALID_ARG = 1,
  WS_ERR_MAX_TOO_HIGH = 2, /* BuildErr::MaxToHigh, lib.rs:1000-1001, 1052 */
  WS_ERR_MAX_TOO_LOW = 3,  /* BuildErr::MaxToLow,  lib.rs:1002-1003, 1053 */
  WS_ERR_SEED_OOB = 4,     /* the reference panics: lib.rs:1366 / 1676     */
  WS_ERR_NO_DEVICE = 5,
  WS_ERR_CUDA = 6,
  WS_ERR_OOM = 7,
  WS_ERR_TOO_LARGE = 8,    /* image or seed list beyond the engine's limits */
  WS_ERR_HOP_OVERFLOW = 9, /* a flood path longer than 2^24 - 2 steps in one level */
  WS_ERR_INTERNAL = 10
} ws_status;

typedef enum ws_kind {
  WS_SEGMENTING = 0, /* SegmentingWatershed, lib.rs:1609-1848 */
  WS_MERGING = 1     /* MergingWatershed,    lib.rs:1297-1562 */
} ws_kind;

/* The fields of TransformBuilder that reach the hot path (lib.rs:917-919).    */
/* Colour of a pixel whose coloured neighbours disagree (lib.rs:246-254).  The reference draws one of the
 * coloured 4-neighbours uniformly with thread_rng(); WS_TIE_FIRST (the default) takes the first of them
 * in the reference's neighbour order down, right, left, up (`col0`, lib.rs:190, 245) -- always one of the
 * reference's possible outcomes, and reproducible.  WS_TIE_RANDOM draws like the reference (uniform over
 * the coloured neighbours, multiplicity included) from a counter-based generator keyed by
 * ws_ctx_set_tie_seed() and the pixel's position, so a given seed reproduces the image.  Merging results
 * (partitions, lake counts, sizes) do not depend on the tie-break.                                      */
typedef enum ws_tie_break { WS_TIE_FIRST = 0, WS_TIE_RANDOM = 1 } ws_tie_break;

typedef struct ws_config {
  uint8_t kind;            /* ws_kind                                          */
  uint8_t max_water_level; /* 1..=254, default NORMAL_MAX (lib.rs:942)         */
  uint8_t edge_correction; /* enable_edge_correction(), lib.rs:958-961         */
  uint8_t tie_break;       /* ws_tie_break (was `reserved`, 0 = WS_TIE_FIRST)  */
} ws_config;

/* ArrayView2<u8>: base pointer of element (0,0), shape, strides in elements.  */
typedef struct ws_image {
  const uint8_t *data;
  size_t rows, cols;
  ptrdiff_t row_stride, col_stride;
} ws_image;

typedef struct ws_ctx ws_ctx; /* device, streams, workspaces                   */

/* ---- context ------------------------------------------------------------ */
ws_status ws_ctx_create(int device, ws_ctx **out);
void ws_ctx_destroy(ws_ctx *ctx);
/* Message of the last failure on this ctx ("" if none); valid until the next call. */
const char *ws_last_error(const ws_ctx *ctx);
const char *ws_status_str(ws_status s);
int ws_abi_version(void);
/* Frees memory handed out by the library (ws_find_local_minima).              */
void ws_free(void *p);
/* Key of the WS_TIE_RANDOM generator (default: drawn from the OS when the ctx is made, like thread_rng). */
ws_status ws_ctx_set_tie_seed(ws_ctx *ctx, uint64_t seed);

/* Host threads of this ctx that move PAGEABLE caller memory (an ordinary Array2 / Vec) through a page-locked
 * ring while the copy engines run, and do the format changes of the boundary on the way (usize <-> u32).
 * 0 = default: min(16, CPUs this process may run on), or the environment variable WS_HOST_THREADS.
 * Page-locked caller memory is copied directly and needs no host thread.                              */
ws_status ws_ctx_set_host_threads(ws_ctx *ctx, int nthreads);
typedef enum ws_option {
  /* usize label outputs in PAGE-LOCKED caller memory: 0 (default) = widened on the device, 8 bytes per pixel
   * over the link, no host work; 1 = 4 bytes per pixel over the link, widened by the host threads (what
   * pageable destinations always get).                                                                */
  WS_OPT_PINNED_HOST_WIDEN = 1
} ws_option;
ws_status ws_ctx_set_option(ws_ctx *ctx, ws_option opt, int value);

/* Pass as `nseeds` (with seeds_rc == NULL) to the transforms below: the starting points are
 * WatershedUtils::find_local_minima(image) (lib.rs:1178-1197), found on the device, so that only the image
 * crosses the link -- what `transform(img, &find_local_minima(img))` of the README computes.  Not available
 * with edge correction (the reference's caller finds the seeds on the unpadded image, lib.rs:1365-1367).  */
#define WS_SEEDS_AUTO ((size_t)-1)

/* ---- TransformBuilder::build_segmenting / build_merging (lib.rs:998-1046) -- */
/* Only the validation is left to do: 1 <= max_water_level <= 254.             */
ws_status ws_config_validate(const ws_config *cfg);
/* Shape of every label image the transform returns: the input shape, or two
 * more in each axis with edge correction (lib.rs:1330-1337; the padding is NOT
 * removed from hook contexts and results, and seeds are NOT shifted, 1365-1367). */
ws_status ws_output_shape(const ws_config *cfg, size_t rows, size_t cols,
                          size_t *out_rows, size_t *out_cols);

/* ---- WatershedUtils::find_local_minima (lib.rs:1178-1197) ---------------- */
/* Interior pixels strictly GREATER than all 8 neighbours (the code, not its
 * doc comment), no plateau resolution, in row-major order.  *out_rc is a
 * library-owned `[n][2]` array of (row, col); release it with ws_free.        */
ws_status ws_find_local_minima(ws_ctx *ctx, const ws_image *img,
                               uint64_t **out_rc, size_t *out_n);

/* ---- Watershed::transform (lib.rs:1208) ---------------------------------- */
/* Segmenting (lib.rs:1810-1822): label image after the last water level.  The
 * reference indexes element 0 of the hook results and panics ("no output?");
 * the intended value -- the colours at max_water_level -- is returned here.
 * Ties between differently coloured neighbours (random in the reference,
 * lib.rs:250-253) resolve to the FIRST coloured neighbour in the reference's
 * order down, right, left, up (lib.rs:190, `col0` of line 245).
 * Merging (lib.rs:1524-1536): interior 123, border 0, seeds ignored.
 * out_labels: out_rows * out_cols uint64, C order.                            */
ws_status ws_transform(ws_ctx *ctx, const ws_config *cfg, const ws_image *img,
                       const uint64_t *seeds_rc, size_t nseeds,
                       uint64_t *out_labels);

/* ---- Watershed::transform_history (lib.rs:1233; 1538-1549; 1824-1835) ---- */
/* One snapshot per water level 0..=max_water_level, in order.
 * out_levels: max+1 bytes.  out_labels: (max+1) * out_rows * out_cols uint64.
 * Merging labels: the representative of a merged lake is its smallest seed
 * colour; the reference's choice (region[0], lib.rs:539) depends on an
 * unspecified sort order, so parity is "equal up to a renumbering per level". */
ws_status ws_transform_history(ws_ctx *ctx, const ws_config *cfg,
                               const ws_image *img, const uint64_t *seeds_rc,
                               size_t nseeds, uint8_t *out_levels,
                               uint64_t *out_labels);

/* ---- Watershed::transform_to_list (lib.rs:1220; 1551-1561; 1837-1847) ----- */
/* Per level the histogram of the label image exactly as find_lake_sizes
 * builds it (lib.rs:629-635): length out_rows*out_cols + 1, index 0 counts the
 * uncoloured pixels.  out_sizes: (max+1) * (out_rows*out_cols + 1) uint64.    */
ws_status ws_transform_to_list(ws_ctx *ctx, const ws_config *cfg,
                               const ws_image *img, const uint64_t *seeds_rc,
                               size_t nseeds, uint8_t *out_levels,
                               uint64_t *out_sizes);

/* ---- Watershed::transform_with_hook (lib.rs:1214; 1328-1522; 1638-1808) --- */
/* HookCtx (lib.rs:844-850).  `image` is the (padded) input, `colours` the
 * label image after this level's merge, both C order and host-resident for the
 * duration of the call; `seeds` is `[nseeds][3]` = (colour, row, col).        */
typedef struct ws_hook_ctx {
  uint8_t water_level;
  uint8_t max_water_level;
  const uint8_t *image;
  const uint64_t *colours;
  size_t rows, cols;
  const uint64_t *seeds;
  size_t nseeds;
} ws_hook_ctx;
typedef void (*ws_level_hook)(void *user, const ws_hook_ctx *hctx);
/* Calls `hook` once per level on the calling thread, in level order.          */
ws_status ws_transform_with_hook(ws_ctx *ctx, const ws_config *cfg,
                                 const ws_image *img, const uint64_t *seeds_rc,
                                 size_t nseeds, ws_level_hook hook, void *user);

/* ---- WatershedUtils::pre_processor / pre_processor_with_max (lib.rs:1081-1173) -- */
typedef enum ws_dtype {
  WS_F32 = 0, WS_F64 = 1, WS_I32 = 2, WS_U16 = 3, WS_I16 = 4, WS_U8 = 5, WS_I64 = 6,
  /* big-endian storage, swapped on the device: the data unit of a FITS file (BITPIX -32, -64, 16, 32, 64) goes
   * up as it is on disk (tests/integration.rs:72-94 reads its cubes from FITS)                          */
  WS_F32_BE = 7, WS_F64_BE = 8, WS_I16_BE = 9, WS_I32_BE = 10, WS_I64_BE = 11
} ws_dtype;
/* `n` elements of a standard-layout array of any dimension -> u8, exactly as the reference does it:
 * min/max folded from ZERO over the finite values (1147-1156); only values whose f64 image
 * `is_normal()` (1161) are scaled to [0, max] with truncation (1163-1164), so 0.0 and subnormals
 * become NEVER_FILL like NaN and -inf (1168-1170); +inf becomes ALWAYS_FILL (1165-1167).
 * max_value: 1..=254 (`MAX` of pre_processor_with_max, asserted at 1143-1144; pre_processor uses
 * NORMAL_MAX).  Host pointers.                                                              */
ws_status ws_pre_processor(ws_ctx *ctx, ws_dtype dtype, const void *data, size_t n,
                           uint8_t max_value, uint8_t *out);
/* Same with device pointers (enqueued on the ctx stream, returns without synchronising). */
ws_status ws_dev_pre_processor(ws_ctx *ctx, ws_dtype dtype, const void *d_data, size_t n,
                               uint8_t max_value, uint8_t *d_out);

/* ---- compact results (extensions; same computation, smaller outputs) ------ */
/* Per level: number of lakes (distinct non-zero labels) and of uncoloured
 * pixels -- what a caller derives from transform_to_list, without the
 * (rows*cols+1)-long rows.  out_*: max+1 uint64 each (either may be NULL).    */
ws_status ws_transform_lake_counts(ws_ctx *ctx, const ws_config *cfg,
                                   const ws_image *img, const uint64_t *seeds_rc,
                                   size_t nseeds, uint64_t *out_lake_counts,
                                   uint64_t *out_uncoloured);
/* Per level the lake sizes without the (rows*cols+1)-long rows of transform_to_list: out_sizes is
 * [max+1][nseeds+1] uint64, column 0 = uncoloured pixels, column c = pixels of colour c (a merged lake is
 * held by its representative, the other columns of the lake are 0) -- the first nseeds+1 entries of every
 * find_lake_sizes row (lib.rs:629-635); the rest of those rows is zero by construction.
 * out_lake_counts (optional): max+1 uint64, number of non-empty lakes per level.                        */
ws_status ws_transform_lake_sizes_compact(ws_ctx *ctx, const ws_config *cfg,
                                          const ws_image *img, const uint64_t *seeds_rc,
                                          size_t nseeds, uint64_t *out_lake_counts,
                                          uint64_t *out_sizes);
/* Final segmenting labels as uint32 plus, per pixel, the water level at which
 * it was coloured (255 = never): snapshot L == (level <= L ? label : 0).      */
ws_status ws_transform_compact(ws_ctx *ctx, const ws_config *cfg,
                               const ws_image *img, const uint64_t *seeds_rc,
                               size_t nseeds, uint32_t *out_labels,
                               uint8_t *out_level);

/* ---- batches of equally shaped slices (one launch set for all of them) ---- */
/* imgs: n_img contiguous C-order slices.  seeds_rc: the slices' seed lists
 * back to back, seed_offsets[n_img+1] delimiting them.  Outputs are optional:
 * out_labels [n_img][out_rows*out_cols] (segmenting final labels),
 * out_lake_counts [n_img][max+1] (lakes per level; the merging statistic).    */
ws_status ws_transform_batch(ws_ctx *ctx, const ws_config *cfg,
                             const uint8_t *imgs, size_t n_img, size_t rows,
                             size_t cols, const uint64_t *seeds_rc,
                             const uint64_t *seed_offsets, uint64_t *out_labels,
                             uint64_t *out_lake_counts);
/* find_local_minima over a batch; *out_rc as in ws_find_local_minima,
 * out_offsets[n_img+1] is caller-allocated.                                   */
ws_status ws_find_local_minima_batch(ws_ctx *ctx, const uint8_t *imgs,
                                     size_t n_img, size_t rows, size_t cols,
                                     uint64_t **out_rc, uint64_t *out_offsets);

/* ---- device-resident pipeline (inputs and results stay in HBM) ------------ */
/* For callers that already hold the data on the GPU (a torch tensor, a decoded
 * FITS cube) and for timing the kernels without PCIe.  All pointers below are
 * device pointers on the ctx's device; work is enqueued on ws_ctx_stream().   */
typedef struct ws_plan ws_plan;
void *ws_ctx_stream(ws_ctx *ctx); /* cudaStream_t */
ws_status ws_ctx_synchronize(ws_ctx *ctx);
/* Workspace for n_img slices of rows x cols (the shape actually flooded).     */
ws_status ws_plan_create(ws_ctx *ctx, size_t n_img, size_t rows, size_t cols,
                         ws_plan **out);
void ws_plan_destroy(ws_plan *plan);
/* Seeds of every slice: d_seeds_rc `[cap][2]` uint32 receives (row, col) in
 * row-major order per slice, d_seed_off[n_img+1] uint32 the offsets.
 * *out_total (host) is the number found; WS_ERR_TOO_LARGE if it exceeds cap.
 * d_seeds_rc == NULL: count only (offsets and *out_total are still written).  */
ws_status ws_plan_find_local_minima(ws_plan *plan, const uint8_t *d_imgs,
                                    uint32_t *d_seeds_rc, size_t cap,
                                    uint32_t *d_seed_off, size_t *out_total);
/* Flood all levels and resolve labels.  kind = WS_MERGING also runs the
 * per-level union-find and fills the lake counts.                             */
ws_status ws_plan_run(ws_plan *plan, const ws_config *cfg, const uint8_t *d_imgs,
                      const uint32_t *d_seeds_rc, const uint32_t *d_seed_off,
                      size_t nseeds_total);
/* Results of the last ws_plan_run (owned by the plan, valid until the next run):
 * labels  [n_img][rows*cols] uint32, segmenting labels (0 = uncoloured);
 * levels  [n_img][rows*cols] uint8, level of colouring (255 = never);
 * counts  [n_img][256] uint32, lakes at each level (merging runs only).       */
const uint32_t *ws_plan_labels(const ws_plan *plan);
const uint8_t *ws_plan_levels(const ws_plan *plan);
const uint32_t *ws_plan_lake_counts(const ws_plan *plan);
/* Diagnostic: arrival times [n_img][rows*cols] uint32 = (level << 24) | hop, hop being
 * the index of the flood pass (lib.rs:1394 'colouring_loop) inside the level;
 * seeds 0, never-coloured pixels >= 0xFF000000.                               */
const uint32_t *ws_plan_arrival_times(const ws_plan *plan);
/* Label image of slice `i` at water level `level` as uint64 into d_out
 * (rows*cols).  Merging snapshots need the run to have been WS_MERGING.       */
ws_status ws_plan_snapshot(ws_plan *plan, ws_kind kind, size_t i, uint8_t level,
                           uint64_t *d_out);
/* Counters of the last run: [0] stale worklist entries dropped, [1] tile activations,
 * [2] pointer-jumping rounds, [3] merge edges, [4] kernels launched,
 * [5] in-tile iteration phases of the flood, [6] / [7] kilo-cycles the flood's consumer
 * warps spent waiting for a staged tile / iterating, summed over the CTAs.     */
ws_status ws_plan_stats(ws_plan *plan, uint64_t out[8]);

/* ---- one large field as row strips over several plans / GPUs --------------
 * The arrival-time fixed point does not depend on the decomposition, so strips stay bit-exact.
 * Each strip is a plan over rows [row_offset, row_offset + rows) of the field INCLUDING one halo
 * row per neighbouring strip; the caller moves boundary rows between strips (NCCL send/recv over
 * NVLink, or plain device copies when the strips share a GPU) and loops until nothing changes.
 * rustronomy-watershed_b200/strips.py is that driver.                                          */
typedef struct ws_strip {
  size_t global_rows;    /* rows of the whole field                                            */
  size_t row_offset;     /* field row of this plan's row 0                                     */
  uint8_t halo_top;      /* 1: row 0 is a copy of the upper neighbour's last owned row         */
  uint8_t halo_bottom;   /* 1: the last row is a copy of the lower neighbour's first owned row */
  uint32_t colour_base;  /* colour of this strip's seed i = colour_base + i + 1                */
} ws_strip;
/* State init, seeds of the OWNED rows (plan-local coordinates), local flood to its fixed point.  */
ws_status ws_plan_strip_begin(ws_plan *plan, const ws_config *cfg, const ws_strip *strip,
                              const uint8_t *d_img, const uint32_t *d_seeds_rc, size_t nseeds);
/* Arrival times of the first / last owned row (cols uint32 each; NULL = not wanted).             */
ws_status ws_plan_strip_export_times(ws_plan *plan, uint32_t *d_top, uint32_t *d_bottom);
/* Min-merge the neighbours' rows into the halo rows and re-run the local flood from the tiles that
 * saw a lower value.  *changed = 1 if a halo value got lower.                                    */
ws_status ws_plan_strip_import_times(ws_plan *plan, const uint32_t *d_top, const uint32_t *d_bottom,
                                 int *changed);
/* Parent pointers + pointer jumping inside the strip; halo pixels stay pending.                 */
ws_status ws_plan_strip_labels(ws_plan *plan);
/* Label words of the first / last owned row (bit 31 set = resolved label).                       */
ws_status ws_plan_strip_export_labels(ws_plan *plan, uint32_t *d_top, uint32_t *d_bottom);
/* Take the neighbours' resolved labels for the halo rows, jump again; *pending = owned pixels
 * still unresolved (loop over export / exchange / import until it is 0 on every strip).          */
ws_status ws_plan_strip_import_labels(ws_plan *plan, const uint32_t *d_top, const uint32_t *d_bottom,
                                      size_t *pending);
/* Forest edges of the strip's tiles as (colour a - 1, colour b - 1) uint32 pairs + a level byte
 * each (device pointers owned by the plan), and the number of this strip's colours that are
 * present on the canvas.  Bit 31 of the second word marks a FINAL edge: a certain edge of the
 * global spanning forest (both basins lie inside the strip, contracted in their tile) that only
 * has to be COUNTED at its level; the others are DEFERRED and go through ws_plan_union_edges.
 * lakes(L) = colours present - unions at levels <= L - FINAL edges at levels <= L.              */
ws_status ws_plan_strip_edges(ws_plan *plan, const void **d_ab, const void **d_w, size_t *n,
                              uint32_t *ndistinct);
/* Kruskal over an edge list gathered from all strips: fills ws_plan_lake_counts()[0..255].      */
ws_status ws_plan_union_edges(ws_plan *plan, const void *d_ab, const void *d_w, size_t n,
                              size_t ncolours, uint32_t ndistinct, uint8_t max_water_level);

/* The same steps without any host synchronisation, for drivers that keep several exchange rounds in flight
 * (strips.py: NCCL on the library's stream): everything is enqueued on ws_ctx_stream(), the `changed` /
 * `pending` results are accumulated into device words of the caller (*d_changed |= 1, *d_pending += n), and
 * errors are collected by ws_plan_strip_check(), the only call that waits.                                  */
ws_status ws_plan_strip_begin_async(ws_plan *plan, const ws_config *cfg, const ws_strip *strip,
                                    const uint8_t *d_img, const uint32_t *d_seeds_rc, size_t nseeds);
ws_status ws_plan_strip_export_times_async(ws_plan *plan, uint32_t *d_top, uint32_t *d_bottom);
ws_status ws_plan_strip_import_times_async(ws_plan *plan, const uint32_t *d_top, const uint32_t *d_bottom,
                                           uint32_t *d_changed);
ws_status ws_plan_strip_labels_async(ws_plan *plan);
ws_status ws_plan_strip_export_labels_async(ws_plan *plan, uint32_t *d_top, uint32_t *d_bottom);
ws_status ws_plan_strip_import_labels_async(ws_plan *plan, const uint32_t *d_top, const uint32_t *d_bottom,
                                            uint32_t *d_pending);
/* (between the asynchronous label rounds only the compact rim array is resolved; this finishes the label
 * plane once the rounds are over -- before ws_plan_labels() is read or ws_plan_strip_forest() runs)      */
ws_status ws_plan_strip_labels_finish_async(ws_plan *plan);
ws_status ws_plan_strip_check(ws_plan *plan);
/* Merging over strips without gathering the strips' edge lists ("boundary union-merge").
 * Every strip reduces its own basin graph to (a) the number of certain forest edges per level and (b) the few
 * forest edges between basins that touch its boundary rows -- at most one per such basin -- and writes both
 * into one device-resident packet of ws_strip_packet_bytes(cols) bytes:
 *   uint32 header[260] = { edges, colours present, error bits, rounds, FINAL edges per level [256] },
 *   then cap = (3 * cols + 64 rounded up to 16) pairs (colour a - 1, colour b - 1) of uint32, then cap level bytes.
 * The driver all-gathers the packets (a few hundred KB each) and hands them, back to back, to
 * ws_plan_forest_packets on any plan of the same width, which leaves the lakes per level of the whole field
 * in ws_plan_lake_counts().  Nothing synchronises; ws_plan_strip_check() reports errors.
 * ncolours_total = number of seeds of the whole field (colours are global: colour_base + i + 1).           */
size_t ws_strip_packet_bytes(size_t cols);
ws_status ws_plan_strip_forest(ws_plan *plan, size_t ncolours_total, void *d_packet);
ws_status ws_plan_forest_packets(ws_plan *plan, const void *d_packets, size_t n_packets,
                                 size_t ncolours_total, uint8_t max_water_level);

/* CUDA-event durations (ms) of the phases of the last run, measured on the ctx stream:
 * [0] state fill + seed colouring, [1] flood kernel, [2] parent + pointer jumping,
 * [3] merging (edges, union-find, counts; 0 for segmenting runs).            */
ws_status ws_plan_phase_ms(ws_plan *plan, float out[4]);

/* CUDA-event durations (ms) of the kernels of the last ws_plan_run, in launch order (0 where a kernel did not run):
 * [0] fill_state (+ worklist reset), [1] seed_init, [2] flood, [3] label_tile, [4] rim_jump, [5] label_finish,
 * [6] merge_reduce, [7] forest_init, [8] forest rounds, [9] lake counts.                                   */
#define WS_KERNEL_SLOTS 10
ws_status ws_plan_kernel_ms(ws_plan *plan, float out[WS_KERNEL_SLOTS]);

/* ---- plain device-memory helpers for callers without a CUDA runtime of their own ---- */
ws_status ws_dev_malloc(ws_ctx *ctx, size_t bytes, void **out);
ws_status ws_dev_free(ws_ctx *ctx, void *d_ptr);
/* Copies ordered after the work already enqueued on the ctx stream; both return
 * after the copy has completed.                                               */
ws_status ws_memcpy_h2d(ws_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
ws_status ws_memcpy_d2h(ws_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
ws_status ws_memcpy_d2d(ws_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* WS_B200_H */
