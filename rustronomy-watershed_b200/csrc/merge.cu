// merge.cu -- K4a: basin-adjacency edges of the merging transform, contracted and reduced per tile.
//
// find_merge (lib.rs:393-445) looks from every coloured window centre at its coloured
// 4-neighbours of a different colour; make_colour_map (lib.rs:467-542) closes those pairs
// transitively at every water level.  In arrival-time terms: two adjacent coloured pixels p, q
// (at least one of them a window centre, lib.rs:411-414) with different segmenting labels a != b
// put an edge (a, b) of weight w = max(level(p), level(q)) into the basin graph G, and the lakes at
// level L are the components of the edges with w <= L, so
//     lakes(L) = colours on the canvas - (edges of a minimum spanning forest of G with w <= L).
// Every CTA takes ONE 64x32 tile (plus the pixels right of / below it) and does two things in shared
// memory, Boruvka-style, under the total edge order (w, tile, edge id in the tile):
//   1. CONTRACTION.  A basin none of whose pixels touches the tile's rim is *closed*: all its edges
//      are in this tile.  The lightest edge leaving a component made of closed basins only is the
//      lightest edge leaving it in the WHOLE graph, hence a forest edge for good (cut property).  It is
//      emitted as FINAL and only counted; the component is merged into its neighbour.  Open basins
//      never pick, so a contracted component holds at most one open basin, which stays its root: the
//      component's identity in every other tile is simply that basin's colour.
//   2. REDUCTION.  Among the contracted components (all open now) only a spanning forest of the
//      tile's remaining edges can matter (cycle property); those edges are emitted as DEFERRED, between
//      the components' open basins, and go through the global level-ordered union-find.
// On a noise field ~3/4 of all basins are closed in their tile, so the global pass -- random accesses
// over tens of millions of colours -- sees a quarter of the graph.  Final edges are kept (flagged) in the
// same list: the merge tree that per-level representatives need is built from all of them on demand.
#include "kernels.cuh"

namespace ws {

constexpr int MR_NW = TILE_W + 1;              // node grid = tile pixels + right / bottom neighbours
constexpr int MR_NH = TILE_H + 1;
constexpr int MR_NODES = MR_NW * MR_NH;        // 2145
constexpr int MR_THREADS = 256;
constexpr int MR_PER_THREAD = (MR_NODES + MR_THREADS - 1) / MR_THREADS;  // 9
constexpr int MR_HASH = 4096;                  // label -> dense id table (load factor <= 0.53)
constexpr uint32_t MR_NONE = 0xFFFFFFFFu;
constexpr uint16_t MR_NOLAB = 0xFFFFu;

struct MergeSmem {
  uint16_t lid[MR_NODES];        // node -> dense id of its basin in this tile (MR_NOLAB: uncoloured)
  uint8_t lvl[MR_NODES + 3];
  uint32_t label_of[MR_NODES];   // dense id -> colour
  uint16_t parent[MR_NODES];     // union-find over dense ids; only a root's own thread re-parents it
  uint16_t comp[MR_NODES];       // root of every id at the start of the round
  uint16_t rep[MR_NODES];        // root after the contraction stage (the component's open basin)
  uint8_t open[MR_NODES];        // dense id: the basin has pixels on the tile's rim
  union {
    uint32_t table[MR_HASH];       // while dense ids are handed out: 0 = free, else id + 1 (open addressing)
    struct {
      uint32_t best[MR_NODES];     // per root: smallest (level << 16 | edge id) offered this round
      uint16_t out_edge[MR_NODES]; // forest edges found (bit 15: FINAL) ...
      uint8_t out_lvl[MR_NODES];   // ... and their levels
    } r;
  } u;
  uint32_t nlab, nout, gpos, first;
};

// `contract` = 0 (row strips: a strip's rim is not only its tiles' rims) skips stage 1.
__global__ void __launch_bounds__(MR_THREADS) merge_reduce_kernel(const uint32_t* __restrict__ lab,
                                                                  const uint8_t* __restrict__ lvl, ImageDims d,
                                                                  const uint32_t* __restrict__ seed_off, int contract,
                                                                  uint2* __restrict__ red_ab, uint8_t* __restrict__ red_w,
                                                                  uint32_t* __restrict__ red_count) {
  __shared__ MergeSmem sm;
  const int tid = threadIdx.x;
  const int tpi = d.tiles_per_img();
  const int img = blockIdx.x / tpi;
  const int trem = blockIdx.x - img * tpi;
  const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
  const int r0 = ty * TILE_H, c0 = tx * TILE_W;
  const size_t base = (size_t)img * d.px_per_img();

  // (a) labels and levels of the tile and of its right / bottom neighbours
  uint32_t L[MR_PER_THREAD];
  if (tid == 0) { sm.nlab = 0; sm.nout = 0; sm.first = 0; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < MR_PER_THREAD; ++k) {
    const int i = tid + k * MR_THREADS;
    L[k] = 0;
    if (i < MR_NODES) {
      const int r = i / MR_NW, c = i - r * MR_NW;
      const int gr = r0 + r, gc = c0 + c;
      uint32_t v = 255;
      if (gr < d.rows && gc < d.cols) {
        const size_t p = base + (size_t)gr * d.cols + gc;
        L[k] = lab[p] & LAB_MASK;
        v = lvl[p];
      }
      sm.lvl[i] = (uint8_t)v;
      sm.open[i] = 0;
      sm.parent[i] = (uint16_t)i;
      if (L[k] != 0u && sm.first == 0u) sm.first = L[k];  // any coloured label (benign race)
    }
  }
  for (int i = tid; i < MR_HASH; i += MR_THREADS) sm.u.table[i] = 0u;
  __syncthreads();
  {  // a tile inside one basin (most tiles of a smooth field) has no edge at all
    const uint32_t f = sm.first;
    bool differs = false;
#pragma unroll
    for (int k = 0; k < MR_PER_THREAD; ++k) differs |= (L[k] != 0u && L[k] != f);
    if (!__syncthreads_or(differs)) return;
  }

  // (b) dense ids, one per distinct label: insert the labels into an open-addressing table, number the
  // occupied slots, then replace every slot's label by its id
  uint16_t slot[MR_PER_THREAD];
#pragma unroll
  for (int k = 0; k < MR_PER_THREAD; ++k) {
    const uint32_t l = L[k];
    uint32_t h = (l * 2654435761u) >> 20;
    if (l != 0u) {
      for (;;) {
        uint32_t cur = ((volatile uint32_t*)sm.u.table)[h];
        if (cur == 0u) cur = atomicCAS(&sm.u.table[h], 0u, l);
        if (cur == 0u || cur == l) break;
        h = (h + 1u) & (MR_HASH - 1);
      }
    }
    slot[k] = (uint16_t)h;
  }
  __syncthreads();
  {
    uint32_t key[MR_HASH / MR_THREADS];
#pragma unroll
    for (int k = 0; k < MR_HASH / MR_THREADS; ++k) key[k] = sm.u.table[tid + k * MR_THREADS];
#pragma unroll
    for (int k = 0; k < MR_HASH / MR_THREADS; ++k) {
      if (key[k] == 0u) continue;
      const uint32_t id = atomicAdd(&sm.nlab, 1u);
      sm.label_of[id] = key[k];
      sm.u.table[tid + k * MR_THREADS] = id;  // (only this thread touches the slot in this phase)
    }
  }
  __syncthreads();
  const int nlab = (int)sm.nlab;
#pragma unroll
  for (int k = 0; k < MR_PER_THREAD; ++k) {
    const int i = tid + k * MR_THREADS;
    if (i >= MR_NODES) continue;
    uint32_t id = MR_NOLAB;
    if (L[k] != 0u) {
      id = sm.u.table[slot[k]];
      const int r = i / MR_NW, c = i - r * MR_NW;
      // rim: pixels whose up / left neighbour lies outside the tile, and the neighbour row / column itself
      if (!contract || r == TILE_H || c == TILE_W || (r == 0 && r0 > 0) || (c == 0 && c0 > 0)) sm.open[id] = 1;
    }
    sm.lid[i] = (uint16_t)id;
  }
  __syncthreads();  // last use of the table; u.r may be written from here on

  // (c) candidate edges of my 8 pixels: endpoints as dense ids in registers, levels (0xFF = no edge)
  const int lc = tid % TILE_W, g = tid / TILE_W;
  uint32_t ew[ROWS_PER_THREAD];   // level of the right edge | level of the down edge << 8
  uint32_t ea[ROWS_PER_THREAD];   // my pixel's id | the right neighbour's id << 16
  uint16_t ed[ROWS_PER_THREAD];   // the lower neighbour's id
  bool have = false;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const int r = g * ROWS_PER_THREAD + i;
    const int n = r * MR_NW + lc;
    const int gr = r0 + r, gc = c0 + lc;
    const uint32_t a = sm.lid[n];
    const uint32_t br = sm.lid[n + 1], bd = sm.lid[n + MR_NW];
    uint32_t wr = 0xFFu, wd = 0xFFu;
    // An edge belongs to the strip that owns its upper / left pixel; a halo row's own edges are the
    // neighbouring strip's.  Plain plans own every row.
    if (a != MR_NOLAB && gr < d.rows && !(d.halo_top && gr == 0) && !(d.halo_bottom && gr == d.rows - 1)) {
      const bool pin = d.is_centre(gr, gc);
      if (br != MR_NOLAB && br != a && (pin || d.is_centre(gr, gc + 1)))
        wr = max((uint32_t)sm.lvl[n], (uint32_t)sm.lvl[n + 1]);
      if (bd != MR_NOLAB && bd != a && (pin || d.is_centre(gr + 1, gc)))
        wd = max((uint32_t)sm.lvl[n], (uint32_t)sm.lvl[n + MR_NW]);
    }
    ew[i] = wr | (wd << 8);
    ea[i] = a | (br << 16);
    ed[i] = (uint16_t)bd;
    have |= (ew[i] != 0xFFFFu);
  }
  if (!__syncthreads_or(have)) return;  // no edge between different basins in this tile

  // (d) stage 0: contraction (only components of closed basins pick), stage 1: spanning forest of the rest.
  // Keys (level << 16 | edge id) are distinct, so the picks of a round form a forest apart from mutual
  // picks of one edge, where the larger root goes under the smaller.  comp[] = root | open << 15.
  constexpr uint32_t OPEN = 0x8000u, ROOT = 0x7FFFu;
  // Per thread: bit 2i / 2i+1 of `live` = right / down edge of pixel i still has to be looked at in this
  // stage; `asleep` = edges between two open components, which only stage 1 can use.
  uint32_t live = 0, asleep = 0;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    if ((ew[i] & 0xFFu) != 0xFFu) live |= 1u << (2 * i);
    if ((ew[i] >> 8) != 0xFFu) live |= 2u << (2 * i);
  }
  for (int stage = contract ? 0 : 1; stage < 2; ++stage) {
    const uint32_t picks = stage == 1 ? OPEN : 0u;  // a component picks when (comp & OPEN) <= picks
    if (stage == 1) live |= asleep;
    for (;;) {
#ifdef WS_MERGE_STATS
      if (tid == 0) atomicAdd(&red_count[4 + stage], 1u);
      atomicAdd(&red_count[6 + stage], (uint32_t)__popc(live));
#endif
      // roots (chains are short; reads racing with the flattening stores still see an ancestor)
      for (int i = tid; i < nlab; i += MR_THREADS) {
        uint32_t x = (uint32_t)i;
        for (uint32_t p = sm.parent[x]; p != x; p = sm.parent[x]) x = p;
        sm.parent[i] = (uint16_t)x;
        sm.comp[i] = (uint16_t)(x | (sm.open[x] ? OPEN : 0u));
        sm.u.r.best[i] = MR_NONE;
      }
      __syncthreads();
      bool any = false;
      if (live) {
        uint32_t still = 0;
#pragma unroll
        for (int i = 0; i < ROWS_PER_THREAD; ++i) {
          if (!(live & (3u << (2 * i)))) continue;
          const uint32_t n = (g * ROWS_PER_THREAD + i) * MR_NW + lc;
          const uint32_t cu = sm.comp[ea[i] & 0xFFFFu];
          const bool pu = (cu & OPEN) <= picks;
          if (live & (1u << (2 * i))) {  // right edge
            const uint32_t cv = sm.comp[ea[i] >> 16];
            if (cu != cv) {
              const bool pv = (cv & OPEN) <= picks;
              if (pu | pv) {
                const uint32_t key = ((ew[i] & 0xFFu) << 16) | (n * 2u);
                if (pu) atomicMin(&sm.u.r.best[cu & ROOT], key);
                if (pv) atomicMin(&sm.u.r.best[cv & ROOT], key);
                any = true;
                still |= 1u << (2 * i);
              } else {
                asleep |= 1u << (2 * i);  // between two open components: nothing to do before stage 1
              }
            }  // else: both ends already in one component -- dead for good
          }
          if (live & (2u << (2 * i))) {  // down edge
            const uint32_t cv = sm.comp[ed[i]];
            if (cu != cv) {
              const bool pv = (cv & OPEN) <= picks;
              if (pu | pv) {
                const uint32_t key = ((ew[i] >> 8) << 16) | (n * 2u + 1u);
                if (pu) atomicMin(&sm.u.r.best[cu & ROOT], key);
                if (pv) atomicMin(&sm.u.r.best[cv & ROOT], key);
                any = true;
                still |= 2u << (2 * i);
              } else {
                asleep |= 2u << (2 * i);
              }
            }
          }
        }
        live = still;
      }
      if (!__syncthreads_or(any)) break;
      // every picking root goes under the component at the other end of its edge
      for (int i = tid; i < nlab; i += MR_THREADS) {
        const uint32_t key = sm.u.r.best[i];
        if (key == MR_NONE) continue;  // (only roots ever receive offers)
        const uint32_t e = key & 0xFFFFu;
        const uint32_t n = e >> 1;
        const uint32_t q = n + ((e & 1u) ? MR_NW : 1);
        const uint32_t cu = sm.comp[sm.lid[n]] & ROOT, cv = sm.comp[sm.lid[q]] & ROOT;
        const uint32_t other = cu == (uint32_t)i ? cv : cu;
        if (sm.u.r.best[other] == key && (uint32_t)i < other) continue;  // mutual pick: the other one moves
        sm.parent[i] = (uint16_t)other;
        const uint32_t o = atomicAdd(&sm.nout, 1u);
        sm.u.r.out_edge[o] = (uint16_t)(e | (stage == 0 ? 0x8000u : 0u));
        sm.u.r.out_lvl[o] = (uint8_t)(key >> 16);
      }
      __syncthreads();
    }
    if (stage == 0) {  // comp[] is current (no hook since the last flatten): the contracted components
      for (int i = tid; i < nlab; i += MR_THREADS) sm.rep[i] = sm.comp[i] & ROOT;
      __syncthreads();
    }
  }

  // (e) append the edges to the global list as global colour ids: FINAL edges (bit 31 of .y) between the
  // two basins themselves, DEFERRED ones between the open basins of the contracted components
  const uint32_t nout = sm.nout;
#ifdef WS_MERGE_STATS
  if (tid == 0) { atomicAdd(&red_count[8], 1u); atomicAdd(&red_count[9], (uint32_t)nlab); atomicAdd(&red_count[10], nout); }
#endif
  if (nout == 0u) return;
  if (tid == 0) sm.gpos = atomicAdd(red_count, nout);
  __syncthreads();
  const uint32_t gpos = sm.gpos;
  const uint32_t gbase = __ldg(seed_off + img) - 1u;  // global colour id = seed_off[img] + colour - 1
  for (uint32_t k = tid; k < nout; k += MR_THREADS) {
    const uint32_t oe = sm.u.r.out_edge[k];
    const uint32_t e = oe & 0x7FFFu, fin = oe >> 15;
    const uint32_t n = e >> 1;
    const uint32_t q = n + ((e & 1u) ? MR_NW : 1);
    uint32_t ia = sm.lid[n], ib = sm.lid[q];
    if (!fin && contract) { ia = sm.rep[ia]; ib = sm.rep[ib]; }
    red_ab[gpos + k] = make_uint2(gbase + sm.label_of[ia], (gbase + sm.label_of[ib]) | (fin << 31));
    red_w[gpos + k] = sm.u.r.out_lvl[k];
  }
}

size_t merge_reduce_capacity(const ImageDims& d) { return (size_t)d.tiles_total() * (MR_NODES - 1); }

cudaError_t launch_merge_reduce(const uint32_t* lab, const uint8_t* lvl, ImageDims d, const uint32_t* seed_off,
                                int contract, uint2* red_ab, uint8_t* red_w, uint32_t* red_count, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(red_count, 0, 16 * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  merge_reduce_kernel<<<d.tiles_total(), MR_THREADS, 0, s>>>(lab, lvl, d, seed_off, contract, red_ab, red_w, red_count);
  return cudaGetLastError();
}

// ---- counting sort of the edges by level (sizes live on the device: no host round trip) ----
// `all` = 0: DEFERRED edges are bucketed for the union-find, FINAL ones only counted per slice and
// level (fin_hist); `all` = 1: every edge is bucketed (the merge tree is built from all of them).

__device__ __forceinline__ int slice_of(const uint32_t* __restrict__ seed_off, int n_img, uint32_t colour) {
  int s0 = 0, s1 = n_img;
  while (s1 - s0 > 1) {
    const int mid = (s0 + s1) >> 1;
    if (__ldg(seed_off + mid) <= colour) s0 = mid; else s1 = mid;
  }
  return s0;
}

__global__ void __launch_bounds__(256) red_hist_kernel(const uint2* __restrict__ red_ab,
                                                       const uint8_t* __restrict__ red_w,
                                                       const uint32_t* __restrict__ red_count, int all,
                                                       const uint32_t* __restrict__ seed_off, int n_img,
                                                       uint32_t* __restrict__ level_hist,
                                                       uint32_t* __restrict__ fin_hist) {
  __shared__ uint32_t s_hist[256], s_fin[256];
  s_hist[threadIdx.x] = 0;
  s_fin[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t n = *red_count;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t w = red_w[i];
    const uint2 e = red_ab[i];
    if (all || !(e.y >> 31)) atomicAdd(&s_hist[w], 1u);
    else if (n_img == 1) atomicAdd(&s_fin[w], 1u);
    else atomicAdd(&fin_hist[(size_t)slice_of(seed_off, n_img, e.x) * 256 + w], 1u);
  }
  __syncthreads();
  if (s_hist[threadIdx.x]) atomicAdd(&level_hist[threadIdx.x], s_hist[threadIdx.x]);
  if (s_fin[threadIdx.x]) atomicAdd(&fin_hist[threadIdx.x], s_fin[threadIdx.x]);
}

__global__ void __launch_bounds__(256) edge_scan_kernel(uint32_t* level_hist, uint32_t* level_cursor) {
  __shared__ uint32_t s[256];
  s[threadIdx.x] = level_hist[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 256; ++i) {
      const uint32_t v = s[i];
      s[i] = run;
      run += v;
    }
    level_hist[256] = run;
  }
  __syncthreads();
  level_hist[threadIdx.x] = s[threadIdx.x];
  level_cursor[threadIdx.x] = s[threadIdx.x];
}

// each CTA takes a contiguous chunk, counts it per level, reserves its ranges, then places the edges
constexpr int RS_CHUNK = 4096;
__global__ void __launch_bounds__(256) red_scatter_kernel(const uint2* __restrict__ red_ab,
                                                          const uint8_t* __restrict__ red_w,
                                                          const uint32_t* __restrict__ red_count, int all,
                                                          uint32_t* __restrict__ level_cursor,
                                                          uint2* __restrict__ edges) {
  __shared__ uint32_t s_cnt[256], s_base[256];
  const uint32_t n = *red_count;
  for (uint32_t c0 = blockIdx.x * RS_CHUNK; c0 < n; c0 += gridDim.x * RS_CHUNK) {
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t slot[RS_CHUNK / 256], w[RS_CHUNK / 256];
    uint2 ab[RS_CHUNK / 256];
#pragma unroll
    for (int k = 0; k < RS_CHUNK / 256; ++k) {
      const uint32_t i = c0 + k * 256 + threadIdx.x;
      w[k] = 0xFFFFu;
      if (i < n) {
        ab[k] = red_ab[i];
        if (all || !(ab[k].y >> 31)) {
          w[k] = red_w[i];
          slot[k] = atomicAdd(&s_cnt[w[k]], 1u);
        }
      }
    }
    __syncthreads();
    if (s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&level_cursor[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_CHUNK / 256; ++k)
      if (w[k] != 0xFFFFu) edges[s_base[w[k]] + slot[k]] = make_uint2(ab[k].x, ab[k].y & LAB_MASK);
    __syncthreads();
  }
}

cudaError_t launch_red_sort(const uint2* red_ab, const uint8_t* red_w, const uint32_t* red_count, int all,
                            const uint32_t* seed_off, int n_img, uint32_t* level_hist, uint32_t* level_cursor,
                            uint32_t* fin_hist, uint2* edges, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(level_hist, 0, 257 * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  if (!all) {
    e = cudaMemsetAsync(fin_hist, 0, sizeof(uint32_t) * 256 * (size_t)n_img, s);
    if (e != cudaSuccess) return e;
  }
  red_hist_kernel<<<148 * 8, 256, 0, s>>>(red_ab, red_w, red_count, all, seed_off, n_img, level_hist, fin_hist);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  edge_scan_kernel<<<1, 256, 0, s>>>(level_hist, level_cursor);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  red_scatter_kernel<<<148 * 8, 256, 0, s>>>(red_ab, red_w, red_count, all, level_cursor, edges);
  return cudaGetLastError();
}

}  // namespace ws
