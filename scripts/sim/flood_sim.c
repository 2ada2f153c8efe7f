// flood_sim.c -- CPU model of the tile scheduler of the flood kernel (design exploration only;
// not product code, not the oracle).  Counts tile activations / sweeps under different policies.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define T_INF 0xFF000000u
static inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

typedef struct {
  int R, C, TW, TH, tx, ty;
  const uint8_t* pix;  // 255 = never
  uint32_t* T;
  uint32_t* key;       // per tile pending key (T_INF.. = none)
  long gs_passes, changed_px;
} Sim;

static inline uint32_t Aof(uint8_t p) { return p == 255 ? T_INF : (((uint32_t)p << 24) | 1u); }


// stage: copy tile + halo from the global state into W ((TH+2) x (TW+2), row stride TW+2)
static void stage_tile(Sim* s, int tY, int tX, uint32_t* W) {
  const int r0 = tY * s->TH - 1, c0 = tX * s->TW - 1, ws = s->TW + 2;
  for (int r = 0; r < s->TH + 2; ++r)
    for (int c = 0; c < ws; ++c) {
      const int gr = r0 + r, gc = c0 + c;
      W[r * ws + c] = (gr >= 0 && gr < s->R && gc >= 0 && gc < s->C) ? s->T[(size_t)gr * s->C + gc] : T_INF;
    }
}
// returns bitmask of woken neighbours; keys[4] = min (v+1) per direction U,D,L,R
static int process_tile(Sim* s, int tY, int tX, uint32_t* W, uint32_t keys[4]) {
  const int r0 = tY * s->TH, c0 = tX * s->TW, ws = s->TW + 2;
  const int h = (r0 + s->TH < s->R ? s->TH : s->R - r0), w = (c0 + s->TW < s->C ? s->TW : s->C - c0);
  const int C = s->C;
  int any = 1;
#define WW(r, c) W[((r) + 1) * ws + (c) + 1]
#define PX(r, c) s->pix[(size_t)(r0 + (r)) * C + c0 + (c)]
  while (any) {
    any = 0;
    s->gs_passes++;
    for (int r = 0; r < h; ++r)
      for (int c = 0; c < w; ++c) {
        const uint8_t p = PX(r, c);
        if (p == 255) continue;
        uint32_t m = umin(umin(WW(r - 1, c), WW(r + 1, c)), umin(WW(r, c - 1), WW(r, c + 1)));
        if (m >= T_INF) continue;
        uint32_t v = umax(Aof(p), m + 1u);
        if (v < WW(r, c)) { WW(r, c) = v; any = 1; }
      }
    for (int r = h - 1; r >= 0; --r)
      for (int c = w - 1; c >= 0; --c) {
        const uint8_t p = PX(r, c);
        if (p == 255) continue;
        uint32_t m = umin(umin(WW(r - 1, c), WW(r + 1, c)), umin(WW(r, c - 1), WW(r, c + 1)));
        if (m >= T_INF) continue;
        uint32_t v = umax(Aof(p), m + 1u);
        if (v < WW(r, c)) { WW(r, c) = v; any = 1; }
      }
  }
  int wake = 0;
  keys[0] = keys[1] = keys[2] = keys[3] = 0xFFFFFFFFu;
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < w; ++c) {
      const uint32_t v = WW(r, c);
      uint32_t* g = &s->T[(size_t)(r0 + r) * C + c0 + c];
      if (v == *g) continue;
      *g = v;
      s->changed_px++;
      if (r == 0 && r0 > 0 && v + 1u < WW(-1, c)) { wake |= 1; keys[0] = umin(keys[0], v + 1u); }
      if (r == h - 1 && r0 + h < s->R && v + 1u < WW(h, c)) { wake |= 2; keys[1] = umin(keys[1], v + 1u); }
      if (c == 0 && c0 > 0 && v + 1u < WW(r, -1)) { wake |= 4; keys[2] = umin(keys[2], v + 1u); }
      if (c == w - 1 && c0 + w < s->C && v + 1u < WW(r, w)) { wake |= 8; keys[3] = umin(keys[3], v + 1u); }
    }
  return wake;
}

// policy 0: FIFO sweeps (every pending tile runs each sweep)
// policy 1: priority: only tiles with key <= ((minlevel + delta) << 24 | 0xFFFFFF) run in a sweep
// stats out: [0] sweeps [1] activations [2] gs passes [3] rounds (sum ceil(n/ncta)) [4] changed px [5] sweeps with n < ncta
int flood_sim(const uint8_t* img, int R, int C, const int32_t* seeds, int nseeds, int TW, int TH, int policy, int delta,
              int ncta, uint32_t* Tout, long* stats, uint32_t* sweep_hist, int hist_cap) {
  Sim s; memset(&s, 0, sizeof s);
  s.R = R; s.C = C; s.TW = TW; s.TH = TH; s.tx = (C + TW - 1) / TW; s.ty = (R + TH - 1) / TH;
  uint8_t* pix = malloc((size_t)R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) {
    uint8_t v = img[(size_t)r * C + c];
    pix[(size_t)r * C + c] = (r >= 1 && r <= R - 2 && c >= 1 && c <= C - 2 && v <= 254) ? v : 255;
  }
  s.pix = pix; s.T = Tout;
  for (size_t i = 0; i < (size_t)R * C; ++i) Tout[i] = T_INF;
  const int nt = s.tx * s.ty;
  s.key = malloc(4 * (size_t)nt);
  for (int i = 0; i < nt; ++i) s.key[i] = 0xFFFFFFFFu;
  for (int i = 0; i < nseeds; ++i) {
    int r = seeds[2 * i], c = seeds[2 * i + 1];
    Tout[(size_t)r * C + c] = 0;
    int t = (r / TH) * s.tx + c / TW;
    s.key[t] = 0;
    if (r % TH == 0 && r > 0) s.key[t - s.tx] = 0;
    if (r % TH == TH - 1 && r + 1 < R) s.key[t + s.tx] = 0;
    if (c % TW == 0 && c > 0) s.key[t - 1] = 0;
    if (c % TW == TW - 1 && c + 1 < C) s.key[t + 1] = 0;
  }

  if (policy == 3) {
    // async bucketed worklist, modelled in lock-step rounds of ncta tiles taken lowest bucket first.
    // bucket = level >> delta; per (tile, bucket) at most one entry (bit mask); DIRTY flag per tile.
    const int shift = delta, NB = (256 >> shift) ? (256 >> shift) : 1;
    uint32_t** bl = malloc(sizeof(uint32_t*) * NB); long* bh = calloc(NB, sizeof(long)); long* bt = calloc(NB, sizeof(long));
    long* bcap = calloc(NB, sizeof(long));
    for (int b = 0; b < NB; ++b) { bcap[b] = 1024; bl[b] = malloc(4 * 1024); }
    uint8_t* inb = calloc((size_t)nt * NB, 1);  // entry present bit
    uint8_t* dirty = calloc(nt, 1);
#define PUSH(t, k) do { int b_ = (int)((k) >> 24) >> shift; if (b_ >= NB) b_ = NB - 1; dirty[t] = 1; \
      int lower_ = 0; for (int q_ = 0; q_ <= b_; ++q_) lower_ |= inb[(size_t)(t) * NB + q_]; \
      if (!lower_) { inb[(size_t)(t) * NB + b_] = 1; if (bt[b_] == bcap[b_]) { bcap[b_] *= 2; bl[b_] = realloc(bl[b_], 4 * bcap[b_]); } bl[b_][bt[b_]++] = (uint32_t)(t); } } while (0)
    for (int i = 0; i < nt; ++i) if (s.key[i] != 0xFFFFFFFFu) {
      const int par = ((i / s.tx) + (i % s.tx)) & 1;
      const uint32_t k = (par && NB > 1) ? ((1u << shift) << 24) : 0u;  // red-black: odd tiles one bucket later
      PUSH(i, k);
    }
    long rounds = 0, acts = 0, skips = 0, small = 0;
    uint32_t* Wbuf = malloc(4 * (size_t)(TW + 2) * (TH + 2) * (size_t)ncta);
    int* cur = malloc(4 * (size_t)ncta);
    const size_t wsz = (size_t)(TW + 2) * (TH + 2);
    for (;;) {
      int n = 0;
      for (int b = 0; b < NB && n < ncta; ++b)
        while (bh[b] < bt[b] && n < ncta) {
          const int t = (int)bl[b][bh[b]++];
          inb[(size_t)t * NB + b] = 0;
          if (!dirty[t]) { skips++; continue; }
          dirty[t] = 0;
          cur[n++] = t;
        }
      if (n == 0) break;
      for (int k = 0; k < n; ++k) stage_tile(&s, cur[k] / s.tx, cur[k] % s.tx, Wbuf + wsz * (size_t)k);
      for (int k = 0; k < n; ++k) {
        const int t = cur[k];
        uint32_t keys[4];
        const int w = process_tile(&s, t / s.tx, t % s.tx, Wbuf + wsz * (size_t)k, keys);
        if (w & 1) PUSH(t - s.tx, keys[0]);
        if (w & 2) PUSH(t + s.tx, keys[1]);
        if (w & 4) PUSH(t - 1, keys[2]);
        if (w & 8) PUSH(t + 1, keys[3]);
      }
      rounds++; acts += n; small += (n < ncta);
    }
    stats[0] = rounds; stats[1] = acts; stats[2] = s.gs_passes; stats[3] = rounds; stats[4] = s.changed_px; stats[5] = small; stats[6] = skips;
    return 0;
  }
  int* list = malloc(4 * (size_t)nt);
  long sweeps = 0, acts = 0, rounds = 0, small = 0;
  for (;;) {
    uint32_t mn = 0xFFFFFFFFu;
    for (int i = 0; i < nt; ++i) mn = umin(mn, s.key[i]);
    if (mn == 0xFFFFFFFFu) break;
    uint32_t thr = 0xFFFFFFFEu;
    if (policy == 1) { uint32_t lv = (mn >> 24) + (uint32_t)delta; if (lv > 254) lv = 254; thr = (lv << 24) | 0xFFFFFFu; }
    if (policy == 2) { thr = mn + (uint32_t)delta; }  // full-T window
    int n = 0;
    // red-black: even tiles first in the list, then odd
    for (int par = 0; par < 2; ++par)
      for (int i = 0; i < nt; ++i) if (s.key[i] <= thr && (((i / s.tx) + (i % s.tx)) & 1) == par) list[n++] = i;
    for (int k = 0; k < n; ++k) s.key[list[k]] = 0xFFFFFFFFu;
    // tiles woken during the sweep run in a LATER sweep (as on the GPU), but see all data written so far
    static uint32_t* newkey = NULL; static int nkcap = 0;
    if (nt > nkcap) { newkey = realloc(newkey, 4 * (size_t)nt); nkcap = nt; }
    for (int i = 0; i < nt; ++i) newkey[i] = 0xFFFFFFFFu;
    static uint32_t* Wbuf = NULL;
    const size_t wsz = (size_t)(TW + 2) * (TH + 2);
    if (!Wbuf) Wbuf = malloc(4 * wsz * (size_t)ncta);
    for (int k0 = 0; k0 < n; k0 += ncta) {
      const int k1 = k0 + ncta < n ? k0 + ncta : n;
      for (int k = k0; k < k1; ++k) stage_tile(&s, list[k] / s.tx, list[k] % s.tx, Wbuf + wsz * (size_t)(k - k0));
      for (int k = k0; k < k1; ++k) {
        const int t = list[k];
        uint32_t keys[4];
        const int w = process_tile(&s, t / s.tx, t % s.tx, Wbuf + wsz * (size_t)(k - k0), keys);
        if (w & 1) newkey[t - s.tx] = umin(newkey[t - s.tx], keys[0]);
        if (w & 2) newkey[t + s.tx] = umin(newkey[t + s.tx], keys[1]);
        if (w & 4) newkey[t - 1] = umin(newkey[t - 1], keys[2]);
        if (w & 8) newkey[t + 1] = umin(newkey[t + 1], keys[3]);
      }
    }
    for (int i = 0; i < nt; ++i) s.key[i] = umin(s.key[i], newkey[i]);
    if (sweeps < hist_cap) sweep_hist[sweeps] = (uint32_t)n;
    sweeps++; acts += n; rounds += (n + ncta - 1) / ncta; small += (n < ncta);
  }
  stats[0] = sweeps; stats[1] = acts; stats[2] = s.gs_passes; stats[3] = rounds; stats[4] = s.changed_px; stats[5] = small;
  free(pix); free(s.key); free(list);
  return 0;
}
