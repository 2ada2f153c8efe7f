// merge.cu -- K4a: basin-adjacency edges of the merging transform, reduced per tile.
//
// find_merge (lib.rs:393-445) looks from every coloured window centre at its coloured
// 4-neighbours of a different colour; make_colour_map (lib.rs:467-542) closes those pairs
// transitively at every water level.  In arrival-time terms: two adjacent coloured pixels p, q
// (at least one of them a window centre, lib.rs:411-414) with different segmenting labels a != b
// put an edge (a, b) of weight w = max(level(p), level(q)) into the basin graph, and the lakes at
// level L are the components of the edges with w <= L.  Only a minimum spanning forest of that graph
// matters (the number of components lost at each level is the same for every MSF), and the MSF of a
// union of edge sets is contained in the union of the sets' MSFs.  So every CTA computes the exact
// minimum spanning forest of the edges of ONE 64x32 tile in shared memory (Boruvka rounds) and emits
// only those forest edges: on noise fields ~5.7x fewer edges reach the global union-find, whose
// random accesses are what the merging costs.
#include "kernels.cuh"

namespace ws {

constexpr int MR_NW = TILE_W + 1;              // node grid = tile pixels + right / bottom neighbours
constexpr int MR_NH = TILE_H + 1;
constexpr int MR_NODES = MR_NW * MR_NH;        // 2145
constexpr int MR_THREADS = 256;
constexpr uint32_t MR_NONE = 0xFFFFFFFFu;

struct MergeSmem {
  uint32_t lab[MR_NODES];
  uint32_t parent[MR_NODES];
  uint32_t best[MR_NODES];       // per component: smallest (level << 16 | edge id) leaving it this round
  uint16_t out_edge[MR_NODES];   // forest edges found (ids) ...
  uint8_t out_lvl[MR_NODES];     // ... and their levels
  uint8_t lvl[MR_NODES + 3];
  uint32_t nout, gpos;
};

__device__ __forceinline__ uint32_t sm_find(volatile uint32_t* parent, uint32_t x) {
  uint32_t p = parent[x];
  while (p != x) {
    const uint32_t gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving; only ever points at an ancestor
    x = p;
    p = gp;
  }
  return x;
}

// true when this call merged two components
__device__ __forceinline__ bool sm_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = sm_find(parent, a);
    b = sm_find(parent, b);
    if (a == b) return false;
    if (a < b) { const uint32_t t = a; a = b; b = t; }
    if (atomicCAS(parent + a, a, b) == a) return true;
  }
}

// Exact minimum spanning forest of the tile's edges by Boruvka rounds in shared memory.  Keys
// (level << 16 | edge id) are distinct, so the edges picked in a round form a forest and every
// successful union is an MSF edge; a 255-step level loop with a CTA barrier per level (the first
// version: 170 us per tile, latency bound) becomes ~log2(components) fully parallel rounds.
__global__ void __launch_bounds__(MR_THREADS) merge_reduce_kernel(const uint32_t* __restrict__ lab,
                                                                  const uint8_t* __restrict__ lvl, ImageDims d,
                                                                  const uint32_t* __restrict__ seed_off, uint32_t lmax,
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
  (void)lmax;

  // (a) nodes: labels and levels of the tile and of its right / bottom neighbours
  for (int i = tid; i < MR_NODES; i += MR_THREADS) {
    const int r = i / MR_NW, c = i - r * MR_NW;
    const int gr = r0 + r, gc = c0 + c;
    uint32_t l = 0, v = 255;
    if (gr < d.rows && gc < d.cols) {
      const size_t p = base + (size_t)gr * d.cols + gc;
      l = lab[p] & LAB_MASK;
      v = lvl[p];
    }
    sm.lab[i] = l;
    sm.lvl[i] = (uint8_t)v;
    sm.parent[i] = i;
  }
  if (tid == 0) sm.nout = 0;
  __syncthreads();

  // (b) candidate edges of my 8 pixels: same label -> one node (union now); different labels ->
  //     remember the edge's level (0xFF = no edge)
  const int lc = tid % TILE_W, g = tid / TILE_W;
  uint32_t ew[ROWS_PER_THREAD];  // level of the right edge | level of the down edge << 8
  bool have = false;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const int r = g * ROWS_PER_THREAD + i;
    const int n = r * MR_NW + lc;
    const int gr = r0 + r, gc = c0 + lc;
    const uint32_t a = sm.lab[n];
    uint32_t wr = 0xFFu, wd = 0xFFu;
    // An edge belongs to the strip that owns its upper / left pixel; a halo row's own edges are the
    // neighbouring strip's.  Plain plans own every row.
    if (a != 0u && gr < d.rows && !(d.halo_top && gr == 0) && !(d.halo_bottom && gr == d.rows - 1)) {
      const bool pin = d.is_centre(gr, gc);
      const uint32_t br = sm.lab[n + 1], bd = sm.lab[n + MR_NW];
      if (br != 0u) {
        if (br == a) sm_union(sm.parent, n, n + 1);
        else if (pin || d.is_centre(gr, gc + 1)) wr = max((uint32_t)sm.lvl[n], (uint32_t)sm.lvl[n + 1]);
      }
      if (bd != 0u) {
        if (bd == a) sm_union(sm.parent, n, n + MR_NW);
        else if (pin || d.is_centre(gr + 1, gc)) wd = max((uint32_t)sm.lvl[n], (uint32_t)sm.lvl[n + MR_NW]);
      }
    }
    ew[i] = wr | (wd << 8);
    have |= (ew[i] != 0xFFFFu);
  }
  if (!__syncthreads_or(have)) return;  // no edge between different basins in this tile

  for (;;) {
    // flatten the forest (read-only walks first, then the writes: a path-halving store racing with
    // another thread's flatten store would leave a non-root behind and break the round's invariant
    // that parent[n] IS the component), clear the per-component minima
    uint32_t root[(MR_NODES + MR_THREADS - 1) / MR_THREADS];
#pragma unroll
    for (int k = 0; k < (MR_NODES + MR_THREADS - 1) / MR_THREADS; ++k) {
      const int i = tid + k * MR_THREADS;
      uint32_t x = i < MR_NODES ? (uint32_t)i : 0u;
      for (uint32_t p = sm.parent[x]; p != x; p = sm.parent[x]) x = p;
      root[k] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < (MR_NODES + MR_THREADS - 1) / MR_THREADS; ++k) {
      const int i = tid + k * MR_THREADS;
      if (i < MR_NODES) {
        sm.parent[i] = root[k];
        sm.best[i] = MR_NONE;
      }
    }
    __syncthreads();
    // smallest edge leaving each component
    bool any = false;
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      if (ew[i] == 0xFFFFu) continue;
      const uint32_t n = (g * ROWS_PER_THREAD + i) * MR_NW + lc;
      const uint32_t ru = sm.parent[n];
      uint32_t wr = ew[i] & 0xFFu, wd = ew[i] >> 8;
      if (wr != 0xFFu) {
        const uint32_t rv = sm.parent[n + 1];
        if (ru != rv) {
          const uint32_t key = (wr << 16) | (n * 2u);
          atomicMin(&sm.best[ru], key);
          atomicMin(&sm.best[rv], key);
          any = true;
        } else wr = 0xFFu;  // both ends already in one component: dead for good
      }
      if (wd != 0xFFu) {
        const uint32_t rv = sm.parent[n + MR_NW];
        if (ru != rv) {
          const uint32_t key = (wd << 16) | (n * 2u + 1u);
          atomicMin(&sm.best[ru], key);
          atomicMin(&sm.best[rv], key);
          any = true;
        } else wd = 0xFFu;
      }
      ew[i] = wr | (wd << 8);
    }
    if (!__syncthreads_or(any)) break;
    // hook every component along its smallest edge
    for (int i = tid; i < MR_NODES; i += MR_THREADS) {
      const uint32_t key = sm.best[i];
      if (key == MR_NONE) continue;
      const uint32_t e = key & 0xFFFFu;
      const uint32_t n = e >> 1;
      const uint32_t q = n + ((e & 1u) ? MR_NW : 1);
      if (sm_union(sm.parent, n, q)) {
        const uint32_t o = atomicAdd(&sm.nout, 1u);
        sm.out_edge[o] = (uint16_t)e;
        sm.out_lvl[o] = (uint8_t)(key >> 16);
      }
    }
    __syncthreads();
  }

  // (f) append the forest edges to the global list, as global colour ids
  const uint32_t nout = sm.nout;
  if (nout == 0u) return;
  if (tid == 0) sm.gpos = atomicAdd(red_count, nout);
  __syncthreads();
  const uint32_t gpos = sm.gpos;
  const uint32_t gbase = __ldg(seed_off + img) - 1u;  // global colour id = seed_off[img] + colour - 1
  for (uint32_t k = tid; k < nout; k += MR_THREADS) {
    const uint32_t e = sm.out_edge[k];
    const uint32_t n = e >> 1;
    const uint32_t q = n + ((e & 1u) ? MR_NW : 1);
    red_ab[gpos + k] = make_uint2(gbase + sm.lab[n], gbase + sm.lab[q]);
    red_w[gpos + k] = sm.out_lvl[k];
  }
}

size_t merge_reduce_capacity(const ImageDims& d) { return (size_t)d.tiles_total() * (MR_NODES - 1); }

cudaError_t launch_merge_reduce(const uint32_t* lab, const uint8_t* lvl, ImageDims d, const uint32_t* seed_off,
                                uint32_t lmax, uint2* red_ab, uint8_t* red_w, uint32_t* red_count, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(red_count, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  merge_reduce_kernel<<<d.tiles_total(), MR_THREADS, 0, s>>>(lab, lvl, d, seed_off, lmax, red_ab, red_w, red_count);
  return cudaGetLastError();
}

// ---- counting sort of the reduced edges by level (sizes live on the device: no host round trip) ----

__global__ void __launch_bounds__(256) red_hist_kernel(const uint8_t* __restrict__ red_w,
                                                       const uint32_t* __restrict__ red_count,
                                                       uint32_t* __restrict__ level_hist) {
  __shared__ uint32_t s_hist[256];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t n = *red_count;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&s_hist[red_w[i]], 1u);
  __syncthreads();
  if (s_hist[threadIdx.x]) atomicAdd(&level_hist[threadIdx.x], s_hist[threadIdx.x]);
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
                                                          const uint32_t* __restrict__ red_count,
                                                          uint32_t* __restrict__ level_cursor,
                                                          uint2* __restrict__ edges) {
  __shared__ uint32_t s_cnt[256], s_base[256];
  const uint32_t n = *red_count;
  for (uint32_t c0 = blockIdx.x * RS_CHUNK; c0 < n; c0 += gridDim.x * RS_CHUNK) {
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t slot[RS_CHUNK / 256], w[RS_CHUNK / 256];
#pragma unroll
    for (int k = 0; k < RS_CHUNK / 256; ++k) {
      const uint32_t i = c0 + k * 256 + threadIdx.x;
      w[k] = 0xFFFFu;
      if (i < n) {
        w[k] = red_w[i];
        slot[k] = atomicAdd(&s_cnt[w[k]], 1u);
      }
    }
    __syncthreads();
    if (s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&level_cursor[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_CHUNK / 256; ++k) {
      const uint32_t i = c0 + k * 256 + threadIdx.x;
      if (w[k] != 0xFFFFu) edges[s_base[w[k]] + slot[k]] = red_ab[i];
    }
    __syncthreads();
  }
}

cudaError_t launch_red_sort(const uint2* red_ab, const uint8_t* red_w, const uint32_t* red_count,
                            uint32_t* level_hist, uint32_t* level_cursor, uint2* edges, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(level_hist, 0, 257 * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  red_hist_kernel<<<148 * 8, 256, 0, s>>>(red_w, red_count, level_hist);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  edge_scan_kernel<<<1, 256, 0, s>>>(level_hist, level_cursor);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  red_scatter_kernel<<<148 * 8, 256, 0, s>>>(red_ab, red_w, red_count, level_cursor, edges);
  return cudaGetLastError();
}

}  // namespace ws
