/* c_abi_smoke.c -- drives libws_b200.so from plain C11 through include/ws_b200.h, with stack structs and caller
 * buffers exactly as the FFI of the reference's language would (shim/src/ffi.rs declares the same calls).
 * Built and run by tests/test_gpu_c_abi.py:
 *     gcc -std=c11 -Wall -Wextra -pedantic -I include tests/c_abi_smoke.c -L rustronomy-watershed_b200 -lws_b200
 * Checks, on an image small enough to follow by hand and on a random one:
 *   build-time validation, find_local_minima, transform (segmenting + merging), transform_with_hook (order,
 *   contents, padded shapes), transform_history == hook snapshots, transform_to_list == histogram of the
 *   snapshots, compact lake sizes, lake counts, random tie-break, error statuses instead of crashes.
 * Prints "c_abi_smoke: OK" and exits 0, or the first failing check and exits 1.                          */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ws_b200.h"

#define CHECK(cond)                                                           \
  do {                                                                        \
    if (!(cond)) {                                                            \
      fprintf(stderr, "c_abi_smoke: FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                               \
    }                                                                         \
  } while (0)
#define OK(call)                                                              \
  do {                                                                        \
    ws_status st_ = (call);                                                   \
    if (st_ != WS_OK) {                                                       \
      fprintf(stderr, "c_abi_smoke: %s -> %s (%s)\n", #call, ws_status_str(st_), ws_last_error(ctx)); \
      return 1;                                                               \
    }                                                                         \
  } while (0)

enum { R = 37, C = 53, LMAX = 254 };

typedef struct hook_state {
  int calls;
  int order_ok;
  size_t rows, cols;
  uint64_t *snapshots; /* [LMAX+1][rows*cols] */
  uint64_t first_seed[3];
} hook_state;

static void on_level(void *user, const ws_hook_ctx *h) {
  hook_state *s = (hook_state *)user;
  if (h->water_level != (uint8_t)s->calls) s->order_ok = 0; /* level order, once per level */
  if (s->calls == 0) {
    s->rows = h->rows;
    s->cols = h->cols;
    if (h->nseeds) memcpy(s->first_seed, h->seeds, sizeof s->first_seed);
  }
  if (s->snapshots)
    memcpy(s->snapshots + (size_t)s->calls * h->rows * h->cols, h->colours, h->rows * h->cols * sizeof(uint64_t));
  s->calls++;
}

static uint32_t lcg(uint32_t *x) { return *x = *x * 1664525u + 1013904223u; }

int main(void) {
  ws_ctx *ctx = NULL;
  CHECK(ws_abi_version() == WS_ABI_VERSION);
  {
    ws_status st = ws_ctx_create(0, &ctx);
    if (st != WS_OK) {
      fprintf(stderr, "c_abi_smoke: ws_ctx_create: %s\n", ws_status_str(st));
      return 1;
    }
  }

  /* ---- TransformBuilder::build_* validation (lib.rs:998-1004) ---- */
  ws_config seg = {WS_SEGMENTING, LMAX, 0, WS_TIE_FIRST};
  ws_config mrg = {WS_MERGING, LMAX, 0, WS_TIE_FIRST};
  CHECK(ws_config_validate(&seg) == WS_OK);
  ws_config bad = seg;
  bad.max_water_level = 255;
  CHECK(ws_config_validate(&bad) == WS_ERR_MAX_TOO_HIGH);
  bad.max_water_level = 0;
  CHECK(ws_config_validate(&bad) == WS_ERR_MAX_TOO_LOW);

  /* ---- a random field (README example, smaller) ---- */
  static uint8_t img[R * C];
  uint32_t rng = 12345u;
  for (int i = 0; i < R * C; ++i) img[i] = (uint8_t)((lcg(&rng) >> 8) % 254u);
  ws_image view = {img, R, C, C, 1};

  uint64_t *seeds = NULL;
  size_t nseeds = 0;
  OK(ws_find_local_minima(ctx, &view, &seeds, &nseeds));
  CHECK(nseeds > 10);
  for (size_t i = 0; i < nseeds; ++i) { /* strict 8-neighbour maxima of interior pixels, row-major order */
    const uint64_t r = seeds[2 * i], c = seeds[2 * i + 1];
    CHECK(r >= 1 && r <= R - 2 && c >= 1 && c <= C - 2);
    for (int dr = -1; dr <= 1; ++dr)
      for (int dc = -1; dc <= 1; ++dc)
        if (dr || dc) CHECK(img[(r + dr) * C + (c + dc)] < img[r * C + c]);
    if (i) CHECK(seeds[2 * i - 2] < r || (seeds[2 * i - 2] == r && seeds[2 * i - 1] < c));
  }

  /* ---- Watershed::transform ---- */
  static uint64_t labels[R * C], merged[R * C];
  OK(ws_transform(ctx, &seg, &view, seeds, nseeds, labels));
  for (size_t i = 0; i < nseeds; ++i) CHECK(labels[seeds[2 * i] * C + seeds[2 * i + 1]] == i + 1); /* lib.rs:1360-1367 */
  for (int c = 0; c < C; ++c) CHECK(labels[c] == 0 && labels[(R - 1) * C + c] == 0); /* only window centres flood */
  for (int r = 1; r < R - 1; ++r)
    for (int c = 1; c < C - 1; ++c) CHECK(labels[r * C + c] >= 1 && labels[r * C + c] <= nseeds);
  OK(ws_transform(ctx, &mrg, &view, seeds, nseeds, merged)); /* lib.rs:1524-1536: interior 123 */
  CHECK(merged[0] == 0 && merged[1 * C + 1] == 123 && merged[(R - 2) * C + (C - 2)] == 123 && merged[R * C - 1] == 0);

  /* ---- transform_with_hook / transform_history / transform_to_list agree ---- */
  hook_state hs = {0, 1, 0, 0, NULL, {0, 0, 0}};
  hs.snapshots = (uint64_t *)malloc((size_t)(LMAX + 1) * R * C * sizeof(uint64_t));
  CHECK(hs.snapshots != NULL);
  OK(ws_transform_with_hook(ctx, &seg, &view, seeds, nseeds, on_level, &hs));
  CHECK(hs.calls == LMAX + 1 && hs.order_ok && hs.rows == R && hs.cols == C);
  CHECK(hs.first_seed[0] == 1 && hs.first_seed[1] == seeds[0] && hs.first_seed[2] == seeds[1]);
  CHECK(memcmp(hs.snapshots + (size_t)LMAX * R * C, labels, sizeof labels) == 0); /* last level == transform */
  {
    uint64_t *hist = (uint64_t *)malloc((size_t)(LMAX + 1) * R * C * sizeof(uint64_t));
    uint8_t levels[LMAX + 1];
    CHECK(hist != NULL);
    OK(ws_transform_history(ctx, &seg, &view, seeds, nseeds, levels, hist));
    for (int l = 0; l <= LMAX; ++l) CHECK(levels[l] == l);
    CHECK(memcmp(hist, hs.snapshots, (size_t)(LMAX + 1) * R * C * sizeof(uint64_t)) == 0);
    free(hist);
  }
  {
    const size_t row = (size_t)R * C + 1; /* find_lake_sizes: rows*cols + 1 entries (lib.rs:630) */
    uint64_t *sizes = (uint64_t *)malloc((size_t)(LMAX + 1) * row * sizeof(uint64_t));
    uint64_t *compact = (uint64_t *)malloc((size_t)(LMAX + 1) * (nseeds + 1) * sizeof(uint64_t));
    uint64_t counts[LMAX + 1], lakes[LMAX + 1], unc[LMAX + 1];
    uint8_t levels[LMAX + 1];
    CHECK(sizes != NULL && compact != NULL);
    OK(ws_transform_to_list(ctx, &seg, &view, seeds, nseeds, levels, sizes));
    OK(ws_transform_lake_sizes_compact(ctx, &seg, &view, seeds, nseeds, counts, compact));
    OK(ws_transform_lake_counts(ctx, &seg, &view, seeds, nseeds, lakes, unc));
    for (int l = 0; l <= LMAX; l += 17) {
      uint64_t *expect = (uint64_t *)calloc(row, sizeof(uint64_t));
      CHECK(expect != NULL);
      for (int i = 0; i < R * C; ++i) expect[hs.snapshots[(size_t)l * R * C + i]]++;
      CHECK(memcmp(expect, sizes + (size_t)l * row, row * sizeof(uint64_t)) == 0);
      CHECK(memcmp(expect, compact + (size_t)l * (nseeds + 1), (nseeds + 1) * sizeof(uint64_t)) == 0);
      uint64_t n = 0;
      for (size_t c = 1; c <= nseeds; ++c) n += expect[c] != 0;
      CHECK(counts[l] == n && lakes[l] == n && unc[l] == expect[0]);
      free(expect);
    }
    free(sizes);
    free(compact);
  }

  /* ---- merging: lakes never increase, snapshots are coarsenings of the segmenting ones ---- */
  {
    uint64_t lakes[LMAX + 1], unc[LMAX + 1];
    OK(ws_transform_lake_counts(ctx, &mrg, &view, seeds, nseeds, lakes, unc));
    for (int l = 1; l <= LMAX; ++l) CHECK(lakes[l] <= lakes[l - 1] && unc[l] <= unc[l - 1]);
    CHECK(lakes[LMAX] == 1 && lakes[0] <= nseeds);
    hook_state hm = {0, 1, 0, 0, NULL, {0, 0, 0}};
    hm.snapshots = (uint64_t *)malloc((size_t)(LMAX + 1) * R * C * sizeof(uint64_t));
    CHECK(hm.snapshots != NULL);
    OK(ws_transform_with_hook(ctx, &mrg, &view, seeds, nseeds, on_level, &hm));
    CHECK(hm.calls == LMAX + 1 && hm.order_ok);
    for (int l = 0; l <= LMAX; l += 51)
      for (int i = 0; i < R * C; ++i) {
        const uint64_t a = hs.snapshots[(size_t)l * R * C + i], b = hm.snapshots[(size_t)l * R * C + i];
        CHECK((a == 0) == (b == 0) && b <= a); /* representative = smallest colour of the lake */
      }
    free(hm.snapshots);
  }

  /* ---- edge correction: outputs two larger per axis, seeds NOT shifted (lib.rs:1330-1367) ---- */
  {
    ws_config ec = {WS_SEGMENTING, 100, 1, WS_TIE_FIRST};
    size_t orows = 0, ocols = 0;
    OK(ws_output_shape(&ec, R, C, &orows, &ocols));
    CHECK(orows == R + 2 && ocols == C + 2);
    uint64_t *out = (uint64_t *)malloc(orows * ocols * sizeof(uint64_t));
    CHECK(out != NULL);
    OK(ws_transform(ctx, &ec, &view, seeds, nseeds, out));
    CHECK(out[seeds[0] * ocols + seeds[1]] == 1);
    hook_state he = {0, 1, 0, 0, NULL, {0, 0, 0}};
    OK(ws_transform_with_hook(ctx, &ec, &view, seeds, nseeds, on_level, &he));
    CHECK(he.calls == 101 && he.rows == orows && he.cols == ocols);
    free(out);
  }

  /* ---- the reference's own tie-break: reproducible per seed, same coloured set ---- */
  {
    ws_config rnd = {WS_SEGMENTING, LMAX, 0, WS_TIE_RANDOM};
    static uint64_t a[R * C], b[R * C], c2[R * C];
    OK(ws_ctx_set_tie_seed(ctx, 7));
    OK(ws_transform(ctx, &rnd, &view, seeds, nseeds, a));
    OK(ws_transform(ctx, &rnd, &view, seeds, nseeds, b));
    OK(ws_ctx_set_tie_seed(ctx, 8));
    OK(ws_transform(ctx, &rnd, &view, seeds, nseeds, c2));
    CHECK(memcmp(a, b, sizeof a) == 0);
    int differs = 0;
    for (int i = 0; i < R * C; ++i) {
      CHECK((a[i] == 0) == (labels[i] == 0) && (c2[i] == 0) == (labels[i] == 0));
      differs |= a[i] != c2[i];
    }
    CHECK(differs);
  }

  /* ---- strided views: the transposed image through strides == the transposed copy ---- */
  {
    static uint8_t tr[R * C];
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) tr[c * R + r] = img[r * C + c];
    ws_image tview = {img, C, R, 1, C}; /* element (i, j) of the transposed image = img[j][i] */
    ws_image tdense = {tr, C, R, R, 1};
    uint64_t *s1 = NULL, *s2 = NULL;
    size_t n1 = 0, n2 = 0;
    OK(ws_find_local_minima(ctx, &tview, &s1, &n1));
    OK(ws_find_local_minima(ctx, &tdense, &s2, &n2));
    CHECK(n1 == n2 && memcmp(s1, s2, 2 * n1 * sizeof(uint64_t)) == 0);
    static uint64_t o1[R * C], o2[R * C];
    OK(ws_transform(ctx, &seg, &tview, s1, n1, o1));
    OK(ws_transform(ctx, &seg, &tdense, s2, n2, o2));
    CHECK(memcmp(o1, o2, sizeof o1) == 0);
    ws_free(s1);
    ws_free(s2);
  }

  /* ---- failures are statuses, not crashes (the reference panics: lib.rs:1366) ---- */
  {
    uint64_t oob[2] = {R, 3};
    CHECK(ws_transform(ctx, &seg, &view, oob, 1, labels) == WS_ERR_SEED_OOB);
    CHECK(strlen(ws_last_error(ctx)) > 0);
    CHECK(ws_transform(ctx, &seg, NULL, seeds, nseeds, labels) == WS_ERR_INVALID_ARG);
    CHECK(ws_transform(ctx, &bad, &view, seeds, nseeds, labels) == WS_ERR_MAX_TOO_LOW);
    OK(ws_transform(ctx, &seg, &view, seeds, nseeds, labels)); /* the context is still usable */
  }

  /* ---- pre_processor (lib.rs:1081-1173) ---- */
  {
    const double in[6] = {1.0, 2.0, 3.0, 0.0, 1.0 / 0.0, -(1.0 / 0.0)};
    uint8_t out[6];
    OK(ws_pre_processor(ctx, WS_F64, in, 6, WS_NORMAL_MAX, out));
    /* min is folded from 0: (x - 0) / 3 * 254, truncated; 0.0 is not `is_normal` -> NEVER_FILL; +inf -> 0 */
    CHECK(out[0] == 84 && out[1] == 169 && out[2] == 254 && out[3] == 255 && out[4] == 0 && out[5] == 255);
  }

  free(hs.snapshots);
  ws_free(seeds);
  ws_ctx_destroy(ctx);
  printf("c_abi_smoke: OK\n");
  return 0;
}
