/*
 * ws_oracle.h -- CPU oracle for the watershed hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the algorithm of
 * smups/rustronomy-watershed v0.4.1 (src/lib.rs); it exists so that the CUDA
 * path can be checked against it and so that bench.py has a CPU baseline that
 * does the reference's work pass for pass.  Nothing under
 * rustronomy-watershed_b200/ may include, link or call it.
 *
 * PARITY PIN: the Rust reference cannot be built in this image (no cargo /
 * rustc, dependencies un-vendored), so the oracle is pinned by the reference's
 * own seven in-file unit tests (src/lib.rs:259-291, 308-311, 336-344, 369-377,
 * 447-465, 544-587, 594-626), whose vectors live in tests/golden/ and are
 * replayed by tests/test_oracle_golden.py.  Whole-transform outputs are NOT
 * pinned by any reference test ("parity unpinned" for find_local_minima and
 * the transform drivers).
 *
 * Every function cites the reference lines it follows.  Labels ("colours") are
 * uint64_t like the reference's usize; images are row-major, axis 0 = row.
 */
#ifndef WS_ORACLE_H
#define WS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/lib.rs:138-141 */
#define ORC_UNCOLOURED 0u
#define ORC_NORMAL_MAX 254u
#define ORC_ALWAYS_FILL 0u
#define ORC_NEVER_FILL 255u

/* Tie-break policies for a pixel whose coloured 4-neighbours disagree
 * (src/lib.rs:246-254).  The reference draws with rand::thread_rng(); FIRST is
 * its own `col0` (line 245) and is the canonical deterministic policy.       */
enum { ORC_TIE_FIRST = 0, ORC_TIE_RANDOM = 1, ORC_TIE_LAST = 2 };

enum { ORC_SEGMENTING = 0, ORC_MERGING = 1 };

/* src/lib.rs:1178-1197.  Strict 8-connected local MAXIMA of the interior, in
 * row-major order.  Writes up to `cap` (row,col) pairs, returns the total.   */
size_t orc_find_local_minima(const uint8_t *img, size_t rows, size_t cols,
                             uint64_t *out_rc, size_t cap);

/* src/lib.rs:196-257.  One synchronous flood step at water level `lvl`.
 * Outputs flat pixel index + colour, row-major order.  Returns the count.
 * `rng` is the xorshift64* state used by ORC_TIE_RANDOM (may be NULL else).
 * out_idx / out_col must hold rows*cols entries.                            */
size_t orc_find_flooded_px(const uint8_t *img, const uint64_t *col, size_t rows,
                           size_t cols, uint8_t lvl, int tie_mode, uint64_t *rng,
                           uint64_t *out_idx, uint64_t *out_col);

/* src/lib.rs:299-306, 314-334, 347-367: Merge equality and the two
 * comparators exactly as written (including their inconsistent branches).
 * Return -1 / 0 / +1 for Less / Equal / Greater.                             */
int orc_merge_eq(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1);
int orc_sort_by_small_big(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1);
int orc_sort_by_big_small(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1);

/* src/lib.rs:393-445.  Unordered colour pairs to merge, as a duplicate-free
 * set; emitted normalised (small, big) and sorted, because the reference's
 * order after its two unstable sorts is unspecified.  out_pairs holds up to
 * `cap` pairs ([n][2]); returns the total.                                   */
size_t orc_find_merge(const uint64_t *col, size_t rows, size_t cols,
                      uint64_t *out_pairs, size_t cap);

/* src/lib.rs:467-542, literally (regions as growable vectors, linear scans). */
void orc_make_colour_map(uint64_t *base_map, size_t map_len,
                         const uint64_t *pairs, size_t npairs);

/* src/lib.rs:590-592 */
void orc_recolour(uint64_t *canvas, size_t n, const uint64_t *colour_map);

/* src/lib.rs:629-635: out has n+1 entries and is zeroed here.                */
void orc_find_lake_sizes(const uint64_t *col, size_t n, uint64_t *out);

/* Per-level callback = the reference's HookCtx (src/lib.rs:844-850) minus the
 * seed list.  `col` is the (possibly padded) label image after the level.    */
typedef void (*orc_hook_fn)(void *user, uint8_t water_level,
                            uint8_t max_water_level, const uint8_t *img,
                            const uint64_t *col, size_t rows, size_t cols);

typedef struct {
  uint64_t flood_passes;    /* find_flooded_px calls incl. terminating ones   */
  uint64_t max_passes_lvl;  /* largest number of passes in a single level     */
  uint64_t contested_px;    /* pixels whose coloured neighbours disagreed     */
  uint64_t merge_pairs;     /* total pairs returned by find_merge             */
} orc_stats;

/* src/lib.rs:1328-1522 (kind = ORC_MERGING) and 1638-1808 (ORC_SEGMENTING).
 * The driver loop, pass for pass.  Output shape is rows x cols, or
 * (rows+2) x (cols+2) with edge_correction (seeds are NOT shifted, 1365-1367).
 * Optional outputs (may be NULL), all of the output shape:
 *   out_final  labels after the last level
 *   out_lvl    water level at which each pixel was coloured (255 = never)
 *   out_hop    flood iteration inside that level (seeds: 0)
 * `fast_closure` != 0 replaces the literal make_colour_map by a union-find
 * with the same partition (representative = smallest colour) for big tests.
 * Returns 0, or -1 on an out-of-bounds seed (the reference panics, 1366).    */
int orc_transform_with_hook(int kind, const uint8_t *img, size_t rows,
                            size_t cols, const uint64_t *seeds_rc, size_t nseeds,
                            uint8_t max_water_level, int edge_correction,
                            int tie_mode, uint64_t rng_seed, int fast_closure,
                            orc_hook_fn hook, void *user, uint64_t *out_final,
                            uint8_t *out_lvl, uint32_t *out_hop,
                            orc_stats *stats);

/* src/lib.rs:1524-1536: constant image, interior 123, border 0.              */
void orc_merging_transform_const(size_t rows, size_t cols, uint64_t *out);

/* src/lib.rs:1134-1173 pre_processor_with_max (pre_processor = MAX 254, lib.rs:1086), for
 * f64 / f32 / i64 element types.  min and max are folded from zero over the finite values,
 * compared in the element type (1147-1156); a value is scaled only if its f64 image is_normal()
 * (1161); +inf -> ALWAYS_FILL (1165-1167), everything else -> NEVER_FILL (1168-1170).
 * Returns -1 if MAX is not in 1..=254 (the asserts at 1143-1144), -2 if to_u8() would fail.   */
int orc_pre_processor_f64(const double *in, size_t n, uint8_t max, uint8_t *out);
int orc_pre_processor_f32(const float *in, size_t n, uint8_t max, uint8_t *out);
int orc_pre_processor_i64(const int64_t *in, size_t n, uint8_t max, uint8_t *out);

/* Number of OpenMP threads the sweeps will use (1 if built without OpenMP).  */
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
