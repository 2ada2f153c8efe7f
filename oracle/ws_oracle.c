/*
 * ws_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see ws_oracle.h).
 *
 * Restates rustronomy-watershed v0.4.1 src/lib.rs function by function.  The
 * sweeps are "pass for pass": every flood iteration visits every 3x3 window of
 * the image like the reference's ndarray::Zip over windows does, every merging
 * level runs a full find_merge scan, a sequential closure, a sequential
 * recolour, and the hook.  OpenMP stands in for rayon on exactly the loops the
 * reference parallelises (lib.rs:220-222, 411-412, 1183-1184).
 */
#include "ws_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* small helpers                                                            */
/* ------------------------------------------------------------------------ */

static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* Compact per-row runs (row r holds cnt[r] entries starting at r*stride) into
 * a dense prefix of the arrays; keeps row-major order.                       */
static size_t compact_rows(uint64_t *a, uint64_t *b, const size_t *cnt,
                           size_t rows, size_t stride) {
  size_t n = 0;
  for (size_t r = 0; r < rows; ++r) {
    if (cnt[r] && n != r * stride) {
      memmove(a + n, a + r * stride, cnt[r] * sizeof(uint64_t));
      if (b) memmove(b + n, b + r * stride, cnt[r] * sizeof(uint64_t));
    }
    n += cnt[r];
  }
  return n;
}

/* ------------------------------------------------------------------------ */
/* find_local_minima  (lib.rs:1178-1197, neighbours_8con 170-185)           */
/* ------------------------------------------------------------------------ */

size_t orc_find_local_minima(const uint8_t *img, size_t rows, size_t cols,
                             uint64_t *out_rc, size_t cap) {
  if (rows < 3 || cols < 3) return 0; /* no 3x3 window exists */
  /* 8-neighbour offsets in the order of lib.rs:172-181 (x = row, y = col).  */
  static const int dx[8] = {1, 1, 1, 0, 0, -1, -1, -1};
  static const int dy[8] = {0, 1, -1, 1, -1, 0, 1, -1};
  size_t n = 0;
  /* window index = top-left corner; centre = idx + (1,1)  (lib.rs:1191)     */
  for (size_t x = 1; x + 1 < rows; ++x) {
    for (size_t y = 1; y + 1 < cols; ++y) {
      const uint8_t target = img[x * cols + y];
      int all_lower = 1;
      for (int k = 0; k < 8; ++k) {
        const uint8_t v = img[(x + dx[k]) * cols + (y + dy[k])];
        if (!(v < target)) { /* lib.rs:1190: every neighbour `<` the target   */
          all_lower = 0;
          break;
        }
      }
      if (all_lower) {
        if (n < cap) {
          out_rc[2 * n] = x;
          out_rc[2 * n + 1] = y;
        }
        ++n;
      }
    }
  }
  return n;
}

/* ------------------------------------------------------------------------ */
/* find_flooded_px  (lib.rs:196-257, neighbours_4con 187-194)               */
/* ------------------------------------------------------------------------ */

static size_t flood_pass(const uint8_t *img, const uint64_t *col, size_t rows,
                         size_t cols, uint8_t lvl, int tie_mode,
                         uint64_t rng_key, uint64_t *out_idx, uint64_t *out_col,
                         size_t *row_cnt, uint64_t *contested) {
  if (rows < 3 || cols < 3) return 0;
  uint64_t n_contested = 0;
  row_cnt[0] = 0;
  row_cnt[rows - 1] = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_contested)
  for (size_t x = 1; x < rows - 1; ++x) {
    size_t k = 0;
    uint64_t *oi = out_idx + x * cols;
    uint64_t *oc = out_col + x * cols;
    const uint8_t *irow = img + x * cols;
    const uint64_t *crow = col + x * cols;
    for (size_t y = 1; y + 1 < cols; ++y) {
      /* (1) flooded?  (2) still uncoloured?                                 */
      if (irow[y] > lvl) continue;
      if (crow[y] != ORC_UNCOLOURED) continue;
      /* neighbours in the order of lib.rs:190: (x+1,y) (x,y+1) (x,y-1) (x-1,y) */
      const uint64_t nb[4] = {crow[y + cols], crow[y + 1], crow[y - 1],
                              crow[y - cols]};
      uint64_t cand[4];
      int nc = 0;
      for (int j = 0; j < 4; ++j)
        if (nb[j] != ORC_UNCOLOURED) cand[nc++] = nb[j];
      /* (3) at least one coloured 4-neighbour                               */
      if (nc == 0) continue;
      /* (4) colour decision (lib.rs:245-254)                                */
      uint64_t c = cand[0];
      int same = 1;
      for (int j = 1; j < nc; ++j)
        if (cand[j] != c) same = 0;
      if (!same) {
        ++n_contested;
        if (tie_mode == ORC_TIE_RANDOM)
          c = cand[splitmix64(rng_key ^ (x * cols + y)) % (uint64_t)nc];
        else if (tie_mode == ORC_TIE_LAST)
          c = cand[nc - 1];
      }
      oi[k] = x * cols + y;
      oc[k] = c;
      ++k;
    }
    row_cnt[x] = k;
  }
  if (contested) *contested += n_contested;
  return compact_rows(out_idx, out_col, row_cnt, rows, cols);
}

size_t orc_find_flooded_px(const uint8_t *img, const uint64_t *col, size_t rows,
                           size_t cols, uint8_t lvl, int tie_mode, uint64_t *rng,
                           uint64_t *out_idx, uint64_t *out_col) {
  size_t *row_cnt = (size_t *)calloc(rows ? rows : 1, sizeof(size_t));
  uint64_t key = 0;
  if (rng) {
    *rng = splitmix64(*rng);
    key = *rng;
  }
  size_t n = flood_pass(img, col, rows, cols, lvl, tie_mode, key, out_idx,
                        out_col, row_cnt, NULL);
  free(row_cnt);
  return n;
}

/* ------------------------------------------------------------------------ */
/* Merge, PartialEq and the two comparators  (lib.rs:293-377)               */
/* ------------------------------------------------------------------------ */

int orc_merge_eq(uint64_t x1, uint64_t y1, uint64_t x2, uint64_t y2) {
  return (x1 == x2 && y1 == y2) || (x1 == y2 && y1 == x2); /* lib.rs:304 */
}

/* lib.rs:319-322 / 352-355 as written: `this` is never reordered (both arms
 * of its `if` are the same) and `that` is reordered the wrong way round.     */
static void cmp_operands(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1,
                         uint64_t *ss, uint64_t *sb, uint64_t *os, uint64_t *ob) {
  *ss = a0;
  *sb = a1;
  if (b0 > b1) {
    *os = b0;
    *ob = b1;
  } else {
    *os = b1;
    *ob = b0;
  }
}

int orc_sort_by_small_big(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1) {
  if (orc_merge_eq(a0, a1, b0, b1)) return 0;
  uint64_t ss, sb, os, ob;
  cmp_operands(a0, a1, b0, b1, &ss, &sb, &os, &ob);
  if (ss < os) return -1; /* lib.rs:325-333 */
  if (ss > os) return 1;
  if (sb < ob) return -1;
  return 1;
}

int orc_sort_by_big_small(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1) {
  if (orc_merge_eq(a0, a1, b0, b1)) return 0;
  uint64_t ss, sb, os, ob;
  cmp_operands(a0, a1, b0, b1, &ss, &sb, &os, &ob);
  if (sb < ob) return -1; /* lib.rs:358-366 */
  if (sb > ob) return 1;
  if (ss < os) return -1;
  return 1;
}

/* ------------------------------------------------------------------------ */
/* find_merge  (lib.rs:393-445)                                             */
/* ------------------------------------------------------------------------ */

typedef struct {
  uint64_t *v;
  size_t n, cap;
} u64vec;

static void vec_push2(u64vec *a, uint64_t x, uint64_t y) {
  if (a->n + 2 > a->cap) {
    a->cap = a->cap ? a->cap * 2 : 256;
    a->v = (uint64_t *)realloc(a->v, a->cap * sizeof(uint64_t));
  }
  a->v[a->n++] = x;
  a->v[a->n++] = y;
}

static int pair_cmp(const void *pa, const void *pb) {
  const uint64_t *a = (const uint64_t *)pa, *b = (const uint64_t *)pb;
  if (a[0] != b[0]) return a[0] < b[0] ? -1 : 1;
  if (a[1] != b[1]) return a[1] < b[1] ? -1 : 1;
  return 0;
}

/* Returns a malloc'd array of unique (small,big) pairs, sorted.             */
static uint64_t *find_merge_alloc(const uint64_t *col, size_t rows, size_t cols,
                                  size_t *npairs) {
  *npairs = 0;
  if (rows < 3 || cols < 3) return NULL;
  int nthr = orc_num_threads();
  u64vec *per = (u64vec *)calloc((size_t)nthr, sizeof(u64vec));
#pragma omp parallel
  {
#ifdef _OPENMP
    u64vec *mine = &per[omp_get_thread_num()];
#else
    u64vec *mine = &per[0];
#endif
#pragma omp for schedule(static)
    for (size_t x = 1; x < rows - 1; ++x) {
      const uint64_t *crow = col + x * cols;
      for (size_t y = 1; y + 1 < cols; ++y) {
        const uint64_t own = crow[y];
        if (own == ORC_UNCOLOURED) continue; /* (1) lib.rs:414 */
        const uint64_t nb[4] = {crow[y + cols], crow[y + 1], crow[y - 1],
                                crow[y - cols]};
        for (int j = 0; j < 4; ++j) {
          /* (2)+(3): coloured neighbours of a different colour, lib.rs:421,432 */
          if (nb[j] == ORC_UNCOLOURED || nb[j] == own) continue;
          /* Merge([a,b]) == Merge([b,a]) (lib.rs:299-306): keep (small,big)  */
          if (own < nb[j])
            vec_push2(mine, own, nb[j]);
          else
            vec_push2(mine, nb[j], own);
        }
      }
    }
  }
  size_t total = 0;
  for (int t = 0; t < nthr; ++t) total += per[t].n;
  uint64_t *all = (uint64_t *)malloc((total ? total : 2) * sizeof(uint64_t));
  size_t off = 0;
  for (int t = 0; t < nthr; ++t) {
    if (per[t].n) memcpy(all + off, per[t].v, per[t].n * sizeof(uint64_t));
    off += per[t].n;
    free(per[t].v);
  }
  free(per);
  size_t np = total / 2;
  /* sort + dedup: the reference sorts twice with its comparators and dedups
   * (lib.rs:440-443); only the resulting SET is well defined.                */
  qsort(all, np, 2 * sizeof(uint64_t), pair_cmp);
  size_t u = 0;
  for (size_t i = 0; i < np; ++i) {
    if (u && all[2 * (u - 1)] == all[2 * i] && all[2 * (u - 1) + 1] == all[2 * i + 1])
      continue;
    all[2 * u] = all[2 * i];
    all[2 * u + 1] = all[2 * i + 1];
    ++u;
  }
  *npairs = u;
  return all;
}

size_t orc_find_merge(const uint64_t *col, size_t rows, size_t cols,
                      uint64_t *out_pairs, size_t cap) {
  size_t n = 0;
  uint64_t *p = find_merge_alloc(col, rows, cols, &n);
  size_t m = n < cap ? n : cap;
  if (m) memcpy(out_pairs, p, m * 2 * sizeof(uint64_t));
  free(p);
  return n;
}

/* ------------------------------------------------------------------------ */
/* make_colour_map  (lib.rs:467-542), literal                               */
/* ------------------------------------------------------------------------ */

typedef struct {
  uint64_t *v;
  size_t n, cap;
} region;

static int region_contains(const region *r, uint64_t c) {
  for (size_t i = 0; i < r->n; ++i)
    if (r->v[i] == c) return 1;
  return 0;
}

static void region_reserve(region *r, size_t extra) {
  if (r->n + extra > r->cap) {
    size_t nc = r->cap ? r->cap : 4;
    while (nc < r->n + extra) nc *= 2;
    r->v = (uint64_t *)realloc(r->v, nc * sizeof(uint64_t));
    r->cap = nc;
  }
}

static int u64_cmp(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

void orc_make_colour_map(uint64_t *base_map, size_t map_len,
                         const uint64_t *pairs, size_t npairs) {
  region *full = NULL; /* full_mergers: Vec<Vec<usize>>, lib.rs:481 */
  size_t nfull = 0, capfull = 0;

  for (size_t p = 0; p < npairs; ++p) {
    const uint64_t col1 = pairs[2 * p], col2 = pairs[2 * p + 1];
    long connect[2] = {-1, -1};
    int duplicate = 0;
    for (size_t idx = 0; idx < nfull; ++idx) {
      const int has1 = region_contains(&full[idx], col1);
      const int has2 = region_contains(&full[idx], col2);
      if (has1 && has2) { /* lib.rs:489-492 */
        duplicate = 1;
        break;
      } else if (has1 || has2) { /* lib.rs:493-502 */
        if (connect[0] < 0) {
          connect[0] = (long)idx;
        } else if (connect[1] < 0) {
          connect[1] = (long)idx;
          break;
        } else {
          abort(); /* "Unreachable code path!" */
        }
      }
    }
    if (duplicate) continue;

    if (connect[0] < 0 && connect[1] < 0) { /* lib.rs:505-508 */
      if (nfull == capfull) {
        capfull = capfull ? capfull * 2 : 16;
        full = (region *)realloc(full, capfull * sizeof(region));
      }
      region r = {NULL, 0, 0};
      region_reserve(&r, 2);
      r.v[0] = col1;
      r.v[1] = col2;
      r.n = 2;
      full[nfull++] = r;
    } else if (connect[1] < 0) { /* lib.rs:509-514: extend, sort, dedup */
      region *reg = &full[connect[0]];
      region_reserve(reg, 2);
      reg->v[reg->n++] = col1;
      reg->v[reg->n++] = col2;
      qsort(reg->v, reg->n, sizeof(uint64_t), u64_cmp);
      size_t u = 0;
      for (size_t i = 0; i < reg->n; ++i)
        if (!u || reg->v[u - 1] != reg->v[i]) reg->v[u++] = reg->v[i];
      reg->n = u;
    } else { /* lib.rs:515-532: drain the later region into the earlier one */
      const long smaller = connect[0] < connect[1] ? connect[0] : connect[1];
      const long larger = connect[0] < connect[1] ? connect[1] : connect[0];
      region *reg1 = &full[smaller], *reg2 = &full[larger];
      region_reserve(reg1, reg2->n);
      memcpy(reg1->v + reg1->n, reg2->v, reg2->n * sizeof(uint64_t));
      reg1->n += reg2->n;
      reg2->n = 0;
    }
    /* remove empty regions, keeping order (lib.rs:535) */
    size_t w = 0;
    for (size_t i = 0; i < nfull; ++i) {
      if (full[i].n == 0) {
        free(full[i].v);
        continue;
      }
      full[w++] = full[i];
    }
    nfull = w;
  }

  /* lib.rs:538-541 */
  for (size_t i = 0; i < nfull; ++i) {
    const uint64_t merged_col = full[i].v[0];
    for (size_t k = 0; k < map_len; ++k)
      if (region_contains(&full[i], base_map[k])) base_map[k] = merged_col;
    free(full[i].v);
  }
  free(full);
}

/* Same partition through a union-find; representative = smallest colour.    */
static uint64_t uf_find(uint64_t *parent, uint64_t x) {
  while (parent[x] != x) {
    parent[x] = parent[parent[x]];
    x = parent[x];
  }
  return x;
}

static void fast_colour_map(uint64_t *base_map, uint64_t *parent, size_t map_len,
                            const uint64_t *pairs, size_t npairs) {
  for (size_t p = 0; p < npairs; ++p) {
    uint64_t a = uf_find(parent, pairs[2 * p]);
    uint64_t b = uf_find(parent, pairs[2 * p + 1]);
    if (a == b) continue;
    if (a < b)
      parent[b] = a;
    else
      parent[a] = b;
  }
  if (npairs)
    for (size_t k = 0; k < map_len; ++k) base_map[k] = uf_find(parent, k);
}

/* ------------------------------------------------------------------------ */
/* recolour (lib.rs:590-592) and find_lake_sizes (lib.rs:629-635)           */
/* ------------------------------------------------------------------------ */

void orc_recolour(uint64_t *canvas, size_t n, const uint64_t *colour_map) {
  for (size_t i = 0; i < n; ++i) canvas[i] = colour_map[canvas[i]];
}

void orc_find_lake_sizes(const uint64_t *col, size_t n, uint64_t *out) {
  memset(out, 0, (n + 1) * sizeof(uint64_t));
  for (size_t i = 0; i < n; ++i) out[col[i]] += 1;
}

/* ------------------------------------------------------------------------ */
/* the drivers  (lib.rs:1328-1522 merging, 1638-1808 segmenting)            */
/* ------------------------------------------------------------------------ */

int orc_transform_with_hook(int kind, const uint8_t *img_in, size_t rows_in,
                            size_t cols_in, const uint64_t *seeds_rc,
                            size_t nseeds, uint8_t max_water_level,
                            int edge_correction, int tie_mode, uint64_t rng_seed,
                            int fast_closure, orc_hook_fn hook, void *user,
                            uint64_t *out_final, uint8_t *out_lvl,
                            uint32_t *out_hop, orc_stats *stats) {
  /* (1a) output shape, lib.rs:1330-1337 / 1640-1647 */
  const size_t rows = edge_correction ? rows_in + 2 : rows_in;
  const size_t cols = edge_correction ? cols_in + 2 : cols_in;
  const size_t npx = rows * cols;

  for (size_t i = 0; i < nseeds; ++i) /* output[idx] panics when out of bounds */
    if (seeds_rc[2 * i] >= rows || seeds_rc[2 * i + 1] >= cols) return -1;

  uint64_t *output = (uint64_t *)calloc(npx ? npx : 1, sizeof(uint64_t));
  /* (1b) padded copy of the input, lib.rs:1340-1356 */
  uint8_t *padded = NULL;
  const uint8_t *img = img_in;
  if (edge_correction) {
    padded = (uint8_t *)calloc(npx ? npx : 1, 1);
    for (size_t r = 0; r < rows_in; ++r)
      memcpy(padded + (r + 1) * cols + 1, img_in + r * cols_in, cols_in);
    img = padded;
  }

  /* (2) colours 1..=nseeds; colour the seeds in order (later wins),
   * then insert UNCOLOURED at index 0.  lib.rs:1360-1369                     */
  uint64_t *colours = (uint64_t *)malloc((nseeds + 1) * sizeof(uint64_t));
  uint64_t *uf = (uint64_t *)malloc((nseeds + 1) * sizeof(uint64_t));
  for (size_t i = 0; i <= nseeds; ++i) colours[i] = uf[i] = i;
  if (out_lvl) memset(out_lvl, 255, npx);
  if (out_hop) memset(out_hop, 0, npx * sizeof(uint32_t));
  for (size_t i = 0; i < nseeds; ++i) {
    const size_t p = seeds_rc[2 * i] * cols + seeds_rc[2 * i + 1];
    output[p] = i + 1;
    if (out_lvl) out_lvl[p] = 0;
  }

  uint64_t *px_idx = (uint64_t *)malloc((npx ? npx : 1) * sizeof(uint64_t));
  uint64_t *px_col = (uint64_t *)malloc((npx ? npx : 1) * sizeof(uint64_t));
  size_t *row_cnt = (size_t *)calloc(rows ? rows : 1, sizeof(size_t));
  orc_stats st = {0, 0, 0, 0};
  uint64_t rng = splitmix64(rng_seed);

  /* (4) the water-level loop, lib.rs:1379 / 1689: 0..=max inclusive          */
  for (unsigned water_level = 0; water_level <= max_water_level; ++water_level) {
    uint64_t passes = 0;
    uint32_t hop = 0;
    for (;;) { /* 'colouring_loop */
      rng = splitmix64(rng);
      const size_t n = flood_pass(img, output, rows, cols, (uint8_t)water_level,
                                  tie_mode, rng, px_idx, px_col, row_cnt,
                                  &st.contested_px);
      ++passes;
      if (n == 0) break; /* lib.rs:1423-1425 */
      ++hop;
      for (size_t k = 0; k < n; ++k) { /* sequential write-back, lib.rs:1431 */
        output[px_idx[k]] = px_col[k];
        if (out_lvl) out_lvl[px_idx[k]] = (uint8_t)water_level;
        if (out_hop) out_hop[px_idx[k]] = hop;
      }
    }
    st.flood_passes += passes;
    if (passes > st.max_passes_lvl) st.max_passes_lvl = passes;

    if (kind == ORC_MERGING) { /* (ii) lib.rs:1450-1466 */
      size_t npairs = 0;
      uint64_t *pairs = find_merge_alloc(output, rows, cols, &npairs);
      st.merge_pairs += npairs;
      if (fast_closure)
        fast_colour_map(colours, uf, nseeds + 1, pairs, npairs);
      else
        orc_make_colour_map(colours, nseeds + 1, pairs, npairs);
      if (colours[ORC_UNCOLOURED] != ORC_UNCOLOURED) abort(); /* lib.rs:1461 */
      if (npairs > 0) orc_recolour(output, npx, colours);
      free(pairs);
    }

    /* (vi) hook after every level, lib.rs:1510-1518 / 1796-1804 */
    if (hook)
      hook(user, (uint8_t)water_level, max_water_level, img, output, rows, cols);
  }

  if (out_final) memcpy(out_final, output, npx * sizeof(uint64_t));
  if (stats) *stats = st;
  free(row_cnt);
  free(px_idx);
  free(px_col);
  free(colours);
  free(uf);
  free(padded);
  free(output);
  return 0;
}

void orc_merging_transform_const(size_t rows, size_t cols, uint64_t *out) {
  memset(out, 0, rows * cols * sizeof(uint64_t));
  if (rows < 3 || cols < 3) return;
  for (size_t r = 1; r + 1 < rows; ++r)
    for (size_t c = 1; c + 1 < cols; ++c) out[r * cols + c] = 123;
}

/* ------------------------------------------------------------------------ */
/* pre_processor_with_max  (lib.rs:1134-1173)                               */
/* ------------------------------------------------------------------------ */

static int pp_map(double f, double min, double max, uint8_t maxv, uint8_t *out) {
  if (isnormal(f)) { /* lib.rs:1161 */
    const volatile double normal = (f - min) / (max - min); /* lib.rs:1163 */
    const volatile double scaled = normal * (double)maxv;   /* lib.rs:1164 */
    /* ToPrimitive::to_u8 on f64: Some(truncated) iff -1 < x < 256 */
    if (!(scaled > -1.0 && scaled < 256.0)) return -2;
    *out = (uint8_t)scaled;
  } else if (isinf(f) && !signbit(f)) {
    *out = ORC_ALWAYS_FILL; /* lib.rs:1165-1167 */
  } else {
    *out = ORC_NEVER_FILL; /* lib.rs:1168-1170 */
  }
  return 0;
}

#define ORC_PP_IMPL(NAME, T)                                                     \
  int NAME(const T *in, size_t n, uint8_t maxv, uint8_t *out) {                  \
    if (!(maxv < ORC_NEVER_FILL) || !(maxv > ORC_ALWAYS_FILL)) return -1;        \
    T mn = (T)0, mx = (T)0; /* fold(T::zero(), ..), lib.rs:1147-1156 */          \
    for (size_t i = 0; i < n; ++i) {                                             \
      const double f = (double)in[i];                                            \
      if (in[i] < mn && isfinite(f)) mn = in[i];                                 \
      if (in[i] > mx && isfinite(f)) mx = in[i];                                 \
    }                                                                            \
    for (size_t i = 0; i < n; ++i) {                                             \
      const int rc = pp_map((double)in[i], (double)mn, (double)mx, maxv, &out[i]); \
      if (rc) return rc;                                                         \
    }                                                                            \
    return 0;                                                                    \
  }

ORC_PP_IMPL(orc_pre_processor_f64, double)
ORC_PP_IMPL(orc_pre_processor_f32, float)
ORC_PP_IMPL(orc_pre_processor_i64, int64_t)
