// kernels.cu -- hand-written sm_100a kernels of the watershed hot path.
//
// Every kernel is integer, HBM/L2/shared-memory bound work; there is no dense
// contraction anywhere on this path, so no tensor-core code.  What matters
// here: coalesced row-major access, shared-memory staging of tiles with their
// halo, persistent cooperative grids sized to the co-resident CTA count
// (multiples of the 148 SMs), and as few passes over HBM as possible.
#include "kernels.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>

namespace cg = cooperative_groups;

namespace ws {

static int g_num_sms = 148;
int num_sms() { return g_num_sms; }
void set_num_sms(int n) { if (n > 0) g_num_sms = n; }

// ===========================================================================
// K1  find_local_minima  (lib.rs:1178-1197)
//
// One CTA = one row segment of MINIMA_CHUNK = 4096 columns; a thread handles MINIMA_GROUPS groups of 4
// pixels, 1024 columns apart, and issues the loads of all groups before it uses any (a chunk is a chain
// load -> stencil -> reduction -> store, and with one group per thread the kernel was bound by that
// latency, not by bandwidth).  When the rows are 4-byte aligned a group is one word per row and the pixels
// left / right of it come from the neighbouring lanes by shuffle.
// Chunks are numbered row-major (slice, row, segment), so an exclusive scan of the per-chunk counts followed
// by an in-chunk rank reproduces the row-major order of the reference's `collect()`.
// ===========================================================================

// mask[g] bit k set <=> pixel (r, c0 + 1024 g + k) is an interior pixel strictly greater than its 8
// neighbours (lib.rs:1190: all(|val| val < target_val)).  Warp-collective.
template <bool kWords>
__device__ __forceinline__ void minima_masks(const uint8_t* __restrict__ img, const ImageDims& d, int b, int r, int c0,
                                             uint32_t mask[MINIMA_GROUPS]) {
#pragma unroll
  for (int g = 0; g < MINIMA_GROUPS; ++g) mask[g] = 0u;
  if (r < 1 || r > d.rows - 2) return;  // (uniform: a CTA is one row segment)
  const uint8_t* base = img + (size_t)b * d.px_per_img();
  const uint8_t* rowp[3] = {base + (size_t)(r - 1) * d.cols, base + (size_t)r * d.cols, base + (size_t)(r + 1) * d.cols};
  uint32_t px[MINIMA_GROUPS][3][6];  // [group][row][c0-1 .. c0+4]
  if (kWords) {
    const int lane = threadIdx.x & 31;
    uint32_t w[MINIMA_GROUPS][3];
#pragma unroll
    for (int g = 0; g < MINIMA_GROUPS; ++g) {
      const int c = c0 + g * 1024;
#pragma unroll
      for (int k = 0; k < 3; ++k) w[g][k] = c < d.cols ? __ldg(reinterpret_cast<const uint32_t*>(rowp[k] + c)) : 0u;
    }
#pragma unroll
    for (int g = 0; g < MINIMA_GROUPS; ++g) {
      const int c = c0 + g * 1024;
      const bool in = c < d.cols;  // then c + 3 < cols as well (cols % 4 == 0)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint32_t x = w[g][k];
        uint32_t left = __shfl_up_sync(0xffffffffu, x, 1) >> 24;
        uint32_t right = __shfl_down_sync(0xffffffffu, x, 1) & 0xFFu;
        if (lane == 0) left = (in && c > 0) ? __ldg(rowp[k] + c - 1) : 0u;
        if (lane == 31) right = (in && c + 4 < d.cols) ? __ldg(rowp[k] + c + 4) : 0u;
        px[g][k][0] = left;
        px[g][k][1] = x & 0xFFu;
        px[g][k][2] = (x >> 8) & 0xFFu;
        px[g][k][3] = (x >> 16) & 0xFFu;
        px[g][k][4] = x >> 24;
        px[g][k][5] = right;
      }
    }
  } else {
#pragma unroll
    for (int g = 0; g < MINIMA_GROUPS; ++g)
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int c = c0 + g * 1024 - 1 + j;
        const bool in = (c >= 0 && c < d.cols);
#pragma unroll
        for (int k = 0; k < 3; ++k) px[g][k][j] = in ? __ldg(rowp[k] + c) : 0u;
      }
  }
#pragma unroll
  for (int g = 0; g < MINIMA_GROUPS; ++g)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + g * 1024 + k;
      if (c < 1 || c > d.cols - 2) continue;
      const uint32_t t = px[g][1][k + 1];
      uint32_t nb = max(max(px[g][0][k], px[g][0][k + 1]), px[g][0][k + 2]);
      nb = max(nb, max(px[g][1][k], px[g][1][k + 2]));
      nb = max(nb, max(max(px[g][2][k], px[g][2][k + 1]), px[g][2][k + 2]));
      if (nb < t) mask[g] |= 1u << k;
    }
}

__device__ __forceinline__ void minima_chunk_coords(const ImageDims& d, size_t chunk, int& b, int& r, int& c0) {
  const int segs = (d.cols + MINIMA_CHUNK - 1) / MINIMA_CHUNK;
  const size_t per_img = (size_t)d.rows * segs;
  b = (int)(chunk / per_img);
  const size_t rem = chunk - (size_t)b * per_img;
  r = (int)(rem / segs);
  c0 = (int)(rem - (size_t)r * segs) * MINIMA_CHUNK + threadIdx.x * 4;
}

template <bool kWords>
__global__ void __launch_bounds__(256) minima_count_kernel(const uint8_t* __restrict__ img, ImageDims d,
                                                           uint32_t* __restrict__ chunk_counts) {
  __shared__ uint32_t s_warp[8];
  int b, r, c0;
  minima_chunk_coords(d, blockIdx.x, b, r, c0);
  uint32_t mask[MINIMA_GROUPS];
  minima_masks<kWords>(img, d, b, r, c0, mask);
  uint32_t n = 0;
#pragma unroll
  for (int g = 0; g < MINIMA_GROUPS; ++g) n += __popc(mask[g]);
#pragma unroll
  for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_warp[w];
    chunk_counts[blockIdx.x] = t;
  }
}

// Single-CTA exclusive scan (n_chunks is at most a few 100k).  Every warp owns a contiguous block of the
// array and walks it 32 values at a time -- coalesced loads, a shuffle scan per step -- instead of every
// thread walking its own 1 KB-strided run (the first version: 0.45 ms for 262144 counts, now 0.10 ms).
__global__ void __launch_bounds__(1024) minima_scan_kernel(uint32_t* __restrict__ v, size_t n, ImageDims d,
                                                           uint32_t* __restrict__ seed_off,
                                                           uint32_t* __restrict__ total) {
  __shared__ uint32_t s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t per = ((n + 31) / 32 + 31) / 32 * 32;  // values per warp, a multiple of 32
  const size_t lo = min(n, (size_t)warp * per), hi = min(n, lo + per);
  // pass 1: the warp's total
  uint32_t sum = 0;
  for (size_t i = lo + lane; i < hi; i += 32) sum += v[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_warp[warp] = sum;
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the 32 warp totals
    const uint32_t mine = s_warp[lane];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    s_warp[lane] = incl - mine;
    if (lane == 31) total[0] = incl;
  }
  __syncthreads();
  // pass 2: exclusive scan inside the warp's block
  uint32_t run = s_warp[warp];
  for (size_t i0 = lo; i0 < hi; i0 += 32) {
    const size_t i = i0 + lane;
    const uint32_t c = i < hi ? v[i] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (i < hi) v[i] = run + incl - c;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncthreads();
  // offsets of the slices = offset of each slice's first chunk
  const int segs = (d.cols + MINIMA_CHUNK - 1) / MINIMA_CHUNK;
  const size_t per_img = (size_t)d.rows * segs;
  for (int b = threadIdx.x; b <= d.n_img; b += 1024)
    seed_off[b] = (b == d.n_img) ? total[0] : v[(size_t)b * per_img];
}

template <bool kWords>
__global__ void __launch_bounds__(256) minima_write_kernel(const uint8_t* __restrict__ img, ImageDims d,
                                                           const uint32_t* __restrict__ chunk_offsets,
                                                           uint32_t* __restrict__ out_rc, uint32_t cap) {
  __shared__ uint32_t s_warp[MINIMA_GROUPS][8];
  int b, r, c0;
  minima_chunk_coords(d, blockIdx.x, b, r, c0);
  uint32_t mask[MINIMA_GROUPS];
  minima_masks<kWords>(img, d, b, r, c0, mask);
  // exclusive rank inside the chunk, in column order: group by group, inside a group thread by thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl[MINIMA_GROUPS];
#pragma unroll
  for (int g = 0; g < MINIMA_GROUPS; ++g) {
    incl[g] = __popc(mask[g]);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl[g], o);
      if (lane >= o) incl[g] += t;
    }
    if (lane == 31) s_warp[g][warp] = incl[g];
  }
  __syncthreads();
  uint32_t group_base = chunk_offsets[blockIdx.x];
#pragma unroll
  for (int g = 0; g < MINIMA_GROUPS; ++g) {
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t x = s_warp[g][w];
      if (w < warp) pre += x;
      tot += x;
    }
    uint32_t pos = group_base + pre + incl[g] - __popc(mask[g]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (mask[g] & (1u << k)) {
        if (pos < cap) *reinterpret_cast<uint2*>(out_rc + 2 * (size_t)pos) = make_uint2((uint32_t)r, (uint32_t)(c0 + g * 1024 + k));
        ++pos;
      }
    }
    group_base += tot;
  }
}

size_t minima_num_chunks(const ImageDims& d) {
  const size_t segs = (d.cols + MINIMA_CHUNK - 1) / MINIMA_CHUNK;
  return (size_t)d.n_img * d.rows * segs;
}

cudaError_t launch_minima_count(const uint8_t* img, ImageDims d, uint32_t* chunk_counts, cudaStream_t s) {
  const bool words = (d.cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 3u) == 0);
  if (words) minima_count_kernel<true><<<(unsigned)minima_num_chunks(d), 256, 0, s>>>(img, d, chunk_counts);
  else minima_count_kernel<false><<<(unsigned)minima_num_chunks(d), 256, 0, s>>>(img, d, chunk_counts);
  return cudaGetLastError();
}
cudaError_t launch_minima_scan(uint32_t* chunk_counts, size_t n_chunks, ImageDims d, uint32_t* seed_off,
                               uint32_t* total, cudaStream_t s) {
  minima_scan_kernel<<<1, 1024, 0, s>>>(chunk_counts, n_chunks, d, seed_off, total);
  return cudaGetLastError();
}
cudaError_t launch_minima_write(const uint8_t* img, ImageDims d, const uint32_t* chunk_offsets, uint32_t* out_rc,
                                uint32_t cap, cudaStream_t s) {
  const bool words = (d.cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 3u) == 0);
  if (words) minima_write_kernel<true><<<(unsigned)minima_num_chunks(d), 256, 0, s>>>(img, d, chunk_offsets, out_rc, cap);
  else minima_write_kernel<false><<<(unsigned)minima_num_chunks(d), 256, 0, s>>>(img, d, chunk_offsets, out_rc, cap);
  return cudaGetLastError();
}

// K2 (flood) and the parent-pointer kernel of K3 live in flood.cu.

static int coop_max_grid(const void* fn, int threads, int device) {
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

// K3 (labels) lives in labels.cu.

// ===========================================================================
// K4  merging: the edge reduction lives in merge.cu; union-find over the reduced edges below
//     (find_merge / make_colour_map / recolour, lib.rs:393-542, 590-592)
// ===========================================================================

// union-find initialisation (the number of colours present on the canvas comes from label_tile)
__global__ void __launch_bounds__(256) uf_init_kernel(MergeBuffers m, uint32_t nseeds) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseeds) return;
  m.parent[i] = i;
  m.hook_to[i] = i;
  m.hook_lvl[i] = 255;
}

// strips: union-find over GLOBAL colours, reset only (the number of colours present comes from the host)
__global__ void __launch_bounds__(256) uf_reset_kernel(MergeBuffers m, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  m.parent[i] = i;
  m.hook_to[i] = i;
  m.hook_lvl[i] = 255;
}
cudaError_t launch_uf_reset(MergeBuffers m, uint32_t n, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(m.unions, 0, sizeof(uint32_t) * 256, s);
  if (e != cudaSuccess || n == 0) return e;
  uf_reset_kernel<<<(n + 255) / 256, 256, 0, s>>>(m, n);
  return cudaGetLastError();
}

// strips: seeds of this strip whose pixel still carries their colour (not overwritten by a later duplicate)
__global__ void __launch_bounds__(256) count_present_kernel(const uint32_t* __restrict__ lab, ImageDims d,
                                                            const uint32_t* __restrict__ seeds_rc, uint32_t nseeds,
                                                            uint32_t colour_base, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool present = false;
  if (i < nseeds) {
    const uint32_t r = seeds_rc[2 * (size_t)i], c = seeds_rc[2 * (size_t)i + 1];
    if (r < (uint32_t)d.rows && c < (uint32_t)d.cols)
      present = (lab[(size_t)r * d.cols + c] & LAB_MASK) == colour_base + i + 1u;
  }
  const int n = __syncthreads_count(present);
  if (threadIdx.x == 0 && n) atomicAdd(out, (uint32_t)n);
}
cudaError_t launch_count_present(const uint32_t* lab, ImageDims d, const uint32_t* seeds_rc, uint32_t nseeds,
                                 uint32_t colour_base, uint32_t* out, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess || nseeds == 0) return e;
  count_present_kernel<<<(nseeds + 255) / 256, 256, 0, s>>>(lab, d, seeds_rc, nseeds, colour_base, out);
  return cudaGetLastError();
}

cudaError_t launch_uf_init(MergeBuffers m, ImageDims d, uint32_t nseeds, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(m.unions, 0, sizeof(uint32_t) * 256 * (size_t)d.n_img, s);
  if (e != cudaSuccess || nseeds == 0) return e;
  uf_init_kernel<<<(nseeds + 255) / 256, 256, 0, s>>>(m, nseeds);
  return cudaGetLastError();
}


__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  uint32_t p = ld_cg(parent + x);
  while (p != x) {
    const uint32_t gp = ld_cg(parent + p);
    if (gp != p) st_cg(parent + x, gp);  // path halving; only ever points at an ancestor
    x = p;
    p = gp;
  }
  return x;
}

// Both roots at once: the two walks are independent chains of dependent loads (L2 or DRAM round trips),
// and a level of union_levels is latency bound -- issue the hops of both walks together.
__device__ __forceinline__ void uf_find2(uint32_t* parent, uint32_t& a, uint32_t& b) {
  uint32_t pa = ld_cg(parent + a), pb = ld_cg(parent + b);
  while (pa != a || pb != b) {
    const uint32_t ga = ld_cg(parent + pa), gb = ld_cg(parent + pb);  // (a root's parent is itself)
    if (pa != a && ga != pa) st_cg(parent + a, ga);  // path halving; only ever points at an ancestor
    if (pb != b && gb != pb) st_cg(parent + b, gb);
    a = pa; pa = ga;
    b = pb; pb = gb;
  }
}

constexpr uint32_t UNION_SMALL_LIMIT = 1u << 18;  // edge lists up to this size go through ONE CTA

// Small edge lists (e.g. a 512x512 field: ~4e4 forest edges): one CTA, a CTA barrier per level
// (~0.1 us) instead of a grid barrier over a thousand CTAs (~5 us) -- 255 levels make the difference.
__global__ void __launch_bounds__(1024) union_levels_small_kernel(MergeBuffers m,
                                                                  const uint32_t* __restrict__ seed_off, int n_img,
                                                                  uint32_t lmax) {
  if (m.level_hist[256] > UNION_SMALL_LIMIT) return;  // the cooperative kernel takes it
  for (uint32_t l = 0; l <= lmax; ++l) {
    const uint32_t lo = m.level_hist[l], hi = m.level_hist[l + 1];
    if (lo == hi) continue;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const uint2 e = m.edges[i];
      uint32_t a = e.x, b = e.y;
      for (;;) {
        a = uf_find(m.parent, a);
        b = uf_find(m.parent, b);
        if (a == b) break;
        if (a < b) { const uint32_t t = a; a = b; b = t; }
        if (atomicCAS(m.parent + a, a, b) == a) {
          m.hook_to[a] = b;
          m.hook_lvl[a] = (uint8_t)l;
          int s0 = 0, s1 = n_img;
          while (s1 - s0 > 1) {
            const int mid = (s0 + s1) >> 1;
            if (__ldg(seed_off + mid) <= a) s0 = mid; else s1 = mid;
          }
          atomicAdd(&m.unions[(size_t)s0 * 256 + l], 1u);
          break;
        }
      }
    }
    __threadfence();
    __syncthreads();
  }
}

// All levels in one persistent cooperative kernel; a grid barrier separates the levels.
__global__ void __launch_bounds__(256) union_levels_kernel(MergeBuffers m, const uint32_t* __restrict__ seed_off,
                                                           int n_img, uint32_t lmax) {
  if (m.level_hist[256] <= UNION_SMALL_LIMIT) return;  // handled by union_levels_small_kernel
  cg::grid_group grid = cg::this_grid();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  __shared__ uint32_t s_ok;  // successes of this CTA at the current level (single-slice runs)
  for (uint32_t l = 0; l <= lmax; ++l) {
    const uint32_t lo = __ldg(m.level_hist + l), hi = __ldg(m.level_hist + l + 1);
    if (lo == hi) continue;  // uniform across the grid
    if (threadIdx.x == 0) s_ok = 0;
    __syncthreads();
    for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
      const uint2 e = m.edges[i];
      uint32_t a = e.x, b = e.y;
      for (;;) {
        uf_find2(m.parent, a, b);
        if (a == b) break;
        if (a < b) { const uint32_t t = a; a = b; b = t; }   // hook the larger root under the smaller
        if (atomicCAS(m.parent + a, a, b) == a) {
          m.hook_to[a] = b;
          m.hook_lvl[a] = (uint8_t)l;
          if (n_img == 1) {
            atomicAdd(&s_ok, 1u);  // one global atomic per CTA and level instead of one per union
          } else {
            int s0 = 0, s1 = n_img;
            while (s1 - s0 > 1) {
              const int mid = (s0 + s1) >> 1;
              if (__ldg(seed_off + mid) <= a) s0 = mid; else s1 = mid;
            }
            atomicAdd(&m.unions[(size_t)s0 * 256 + l], 1u);
          }
          break;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_ok) atomicAdd(&m.unions[l], s_ok);
    grid.sync();
  }
}

// The kernel is bound by its 255 grid barriers and by dependent L2 round trips, not by thread count: after
// the per-tile contraction a level holds ~3e4 edges, and a barrier over 2 CTAs per SM is much cheaper
// than one over 8 (merge phase, uniform 16384^2: 1 CTA per SM 10.2 ms, 2 per SM 9.6 ms, 8 per SM 9.8 ms).
int union_max_grid(int device) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int full = coop_max_grid((const void*)union_levels_kernel, 256, device);
  return full < 2 * sms ? full : 2 * sms;
}

cudaError_t launch_union_levels(MergeBuffers m, const uint32_t* seed_off, int n_img, uint32_t lmax, int grid,
                                cudaStream_t s) {
  union_levels_small_kernel<<<1, 1024, 0, s>>>(m, seed_off, n_img, lmax);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  void* args[] = {&m, &seed_off, &n_img, &lmax};
  return cudaLaunchCooperativeKernel((const void*)union_levels_kernel, dim3(grid), dim3(256), args, 0, s);
}

// counts[img][l] = colours present - forest edges at levels <= l (unions of the global pass + FINAL edges)
__global__ void lake_counts_kernel(MergeBuffers m, int n_img, uint32_t lmax) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n_img) return;
  uint32_t n = m.ndistinct[img];
  for (uint32_t l = 0; l < 256; ++l) {
    if (l <= lmax) n -= m.unions[(size_t)img * 256 + l] + m.fin_hist[(size_t)img * 256 + l];
    m.counts[(size_t)img * 256 + l] = (l <= lmax) ? n : 0u;
  }
}

cudaError_t launch_lake_counts(MergeBuffers m, int n_img, uint32_t lmax, cudaStream_t s) {
  lake_counts_kernel<<<(n_img + 63) / 64, 64, 0, s>>>(m, n_img, lmax);
  return cudaGetLastError();
}

// Representative of every colour at `level`: follow the hook links whose level is <= level.
// Link levels never decrease along a chain (levels are processed in order), so the first
// link above `level` ends the walk.
__global__ void __launch_bounds__(256) rep_table_kernel(const uint32_t* __restrict__ hook_to,
                                                        const uint8_t* __restrict__ hook_lvl, uint32_t nseeds,
                                                        uint32_t level, int incremental, uint32_t* rep) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseeds) return;
  uint32_t x = incremental ? rep[i] : i;
  while (hook_lvl[x] <= level) x = hook_to[x];
  rep[i] = x;
}

cudaError_t launch_rep_table(const uint32_t* hook_to, const uint8_t* hook_lvl, uint32_t nseeds, uint32_t level,
                             int incremental, uint32_t* rep, cudaStream_t s) {
  if (nseeds == 0) return cudaSuccess;
  rep_table_kernel<<<(nseeds + 255) / 256, 256, 0, s>>>(hook_to, hook_lvl, nseeds, level, incremental, rep);
  return cudaGetLastError();
}

// ===========================================================================
// K5-K7  per-level outputs
// ===========================================================================

template <typename O>
__global__ void __launch_bounds__(256) snapshot_kernel(const uint32_t* __restrict__ lab,
                                                       const uint8_t* __restrict__ lvl, size_t n, uint32_t level,
                                                       const uint32_t* __restrict__ rep, uint32_t colour_base,
                                                       O* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t v = 0;
    if (lvl[i] <= level) {
      v = lab[i] & LAB_MASK;
      if (rep && v) v = rep[colour_base + v - 1u] - colour_base + 1u;
    }
    out[i] = v;
  }
}

static unsigned stream_grid(size_t n) {
  const size_t want = (n + 255) / 256;
  const size_t cap = (size_t)num_sms() * 16;
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

cudaError_t launch_snapshot(const uint32_t* lab, const uint8_t* lvl, size_t n_px, uint32_t level,
                            const uint32_t* rep, uint32_t colour_base, uint64_t* out, cudaStream_t s) {
  snapshot_kernel<uint64_t><<<stream_grid(n_px), 256, 0, s>>>(lab, lvl, n_px, level, rep, colour_base, out);
  return cudaGetLastError();
}
// the same as 32-bit words (pageable destinations: half the bytes on the link, widened by the host's workers)
cudaError_t launch_snapshot32(const uint32_t* lab, const uint8_t* lvl, size_t n_px, uint32_t level,
                              const uint32_t* rep, uint32_t colour_base, uint32_t* out, cudaStream_t s) {
  snapshot_kernel<uint32_t><<<stream_grid(n_px), 256, 0, s>>>(lab, lvl, n_px, level, rep, colour_base, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) widen_kernel(const uint32_t* __restrict__ lab, size_t n,
                                                    uint64_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = lab[i] & LAB_MASK;
}
cudaError_t launch_widen_labels(const uint32_t* lab, size_t n_px, uint64_t* out, cudaStream_t s) {
  widen_kernel<<<stream_grid(n_px), 256, 0, s>>>(lab, n_px, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) strip_kernel(const uint32_t* __restrict__ lab, size_t n,
                                                    uint32_t* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = lab[i] & LAB_MASK;
}
cudaError_t launch_strip_labels(const uint32_t* lab, size_t n_px, uint32_t* out, cudaStream_t s) {
  strip_kernel<<<stream_grid(n_px), 256, 0, s>>>(lab, n_px, out);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) level_hist_kernel(const uint8_t* __restrict__ lvl, ImageDims d,
                                                         uint32_t* __restrict__ lvl_hist) {
  __shared__ uint32_t s_hist[256];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const size_t n = d.px_per_img();
  const uint8_t* base = lvl + (size_t)blockIdx.y * n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&s_hist[base[i]], 1u);
  __syncthreads();
  if (s_hist[threadIdx.x]) atomicAdd(&lvl_hist[(size_t)blockIdx.y * 256 + threadIdx.x], s_hist[threadIdx.x]);
}
cudaError_t launch_level_hist(const uint8_t* lvl, ImageDims d, uint32_t* lvl_hist, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(lvl_hist, 0, sizeof(uint32_t) * 256 * (size_t)d.n_img, s);
  if (e != cudaSuccess) return e;
  const size_t n = d.px_per_img();
  const unsigned gx = (unsigned)min((size_t)592, (n + 255) / 256);
  level_hist_kernel<<<dim3(gx ? gx : 1, d.n_img), 256, 0, s>>>(lvl, d, lvl_hist);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) colour_level_count_kernel(const uint32_t* __restrict__ lab,
                                                                 const uint8_t* __restrict__ lvl, size_t n,
                                                                 uint32_t ncol, uint32_t* __restrict__ cnt) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t l = lvl[i];
    if (l == 255u) continue;
    atomicAdd(&cnt[(size_t)l * ncol + (lab[i] & LAB_MASK)], 1u);
  }
}
cudaError_t launch_colour_level_count(const uint32_t* lab, const uint8_t* lvl, size_t n_px, uint32_t ncol,
                                      uint32_t* cnt, cudaStream_t s) {
  colour_level_count_kernel<<<stream_grid(n_px), 256, 0, s>>>(lab, lvl, n_px, ncol, cnt);
  return cudaGetLastError();
}

// one thread per colour walks the levels; column 0 (uncoloured) is filled by colour-0's thread
// from the per-level totals accumulated with one atomic per (colour, level) pair.
__global__ void __launch_bounds__(256) sizes_cumulate_kernel(const uint32_t* __restrict__ cnt, uint32_t ncol,
                                                             uint32_t nlevels, uint64_t* __restrict__ sizes) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol || c == 0u) return;
  uint64_t run = 0;
  for (uint32_t l = 0; l < nlevels; ++l) {
    run += cnt[(size_t)l * ncol + c];
    sizes[(size_t)l * ncol + c] = run;
  }
}
__global__ void __launch_bounds__(256) sizes_uncoloured_kernel(const uint32_t* __restrict__ cnt, uint32_t ncol,
                                                               uint32_t nlevels, size_t n_px,
                                                               uint64_t* __restrict__ sizes) {
  // block l sums row l of cnt; thread 0 of block 0 then turns the totals into the running column 0
  __shared__ unsigned long long s_sum;
  __shared__ unsigned long long s_tot[256];
  for (uint32_t l = 0; l < nlevels; ++l) {
    if (threadIdx.x == 0) s_sum = 0ull;
    __syncthreads();
    unsigned long long acc = 0;
    for (uint32_t c = 1 + threadIdx.x; c < ncol; c += blockDim.x) acc += cnt[(size_t)l * ncol + c];
    atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) s_tot[l] = s_sum;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (uint32_t l = 0; l < nlevels; ++l) {
      run += s_tot[l];
      sizes[(size_t)l * ncol] = (uint64_t)n_px - run;
    }
  }
}
cudaError_t launch_sizes_cumulate(const uint32_t* cnt, uint32_t ncol, uint32_t nlevels, size_t n_px,
                                  uint64_t* sizes, cudaStream_t s) {
  sizes_cumulate_kernel<<<(ncol + 255) / 256, 256, 0, s>>>(cnt, ncol, nlevels, sizes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sizes_uncoloured_kernel<<<1, 256, 0, s>>>(cnt, ncol, nlevels, n_px, sizes);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sizes_fold_kernel(const uint64_t* __restrict__ row,
                                                         const uint32_t* __restrict__ rep, uint32_t ncol,
                                                         uint64_t* __restrict__ scratch) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol || c == 0u) return;
  const uint64_t v = row[c];
  if (v) atomicAdd((unsigned long long*)&scratch[rep[c - 1u] + 1u], (unsigned long long)v);
}
cudaError_t launch_sizes_fold(uint64_t* sizes_row, const uint32_t* rep, uint32_t ncol, uint64_t* scratch_row,
                              cudaStream_t s) {
  // scratch <- 0 ; scratch[rep(c)] += row[c] ; row[1..] <- scratch[1..]   (column 0 is kept)
  cudaError_t e = cudaMemsetAsync(scratch_row, 0, sizeof(uint64_t) * ncol, s);
  if (e != cudaSuccess) return e;
  sizes_fold_kernel<<<(ncol + 255) / 256, 256, 0, s>>>(sizes_row, rep, ncol, scratch_row);
  e = cudaGetLastError();
  if (e != cudaSuccess || ncol <= 1) return e;
  return cudaMemcpyAsync(sizes_row + 1, scratch_row + 1, sizeof(uint64_t) * (ncol - 1), cudaMemcpyDeviceToDevice, s);
}

__global__ void __launch_bounds__(256) fill_const123_kernel(uint64_t* __restrict__ out, int rows, int cols) {
  const size_t n = (size_t)rows * cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int r = (int)(i / cols), c = (int)(i - (size_t)r * cols);
    out[i] = (r >= 1 && r <= rows - 2 && c >= 1 && c <= cols - 2) ? 123ull : 0ull;
  }
}
cudaError_t launch_fill_const123(uint64_t* out, int rows, int cols, cudaStream_t s) {
  fill_const123_kernel<<<stream_grid((size_t)rows * cols), 256, 0, s>>>(out, rows, cols);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) pad_image_kernel(const uint8_t* __restrict__ src, int rows, int cols,
                                                        uint8_t* __restrict__ dst) {
  const int pc = cols + 2;
  const size_t n = (size_t)(rows + 2) * pc;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int r = (int)(i / pc), c = (int)(i - (size_t)r * pc);
    uint8_t v = 0;
    if (r >= 1 && r <= rows && c >= 1 && c <= cols) v = src[(size_t)(r - 1) * cols + (c - 1)];
    dst[i] = v;
  }
}
cudaError_t launch_pad_image(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s) {
  pad_image_kernel<<<stream_grid((size_t)(rows + 2) * (cols + 2)), 256, 0, s>>>(src, rows, cols, dst);
  return cudaGetLastError();
}

}  // namespace ws

// ===========================================================================
// K8  pre_processor / pre_processor_with_max  (lib.rs:1081-1173)
//
// Any numeric array -> u8, with the reference's exact behaviour:
//   min / max are folded from ZERO over the finite values (lib.rs:1147-1156), compared in T;
//   a value maps through ((x - min) / (max - min)) * MAX, truncated (lib.rs:1163-1164), only if
//   its f64 image `is_normal()` (lib.rs:1161) -- so 0.0 and subnormals take the else branches;
//   +inf -> ALWAYS_FILL (lib.rs:1165-1167; the comment there says "negative infinity", the test
//   `!is_sign_negative()` selects the positive one); everything else (NaN, -inf, 0, subnormal)
//   -> NEVER_FILL (lib.rs:1168-1170).
// f64 arithmetic with explicit round-to-nearest operations (no FMA contraction) so the bytes
// equal the CPU's.
// ===========================================================================
namespace ws {

template <typename T>
__device__ __forceinline__ double pp_to_f64(T v) { return (double)v; }

// Element loads.  kSwap: the array holds big-endian values (the byte order of a FITS file), swapped on the way in,
// so a cube's raw data unit goes to the device as it is on disk.
template <typename T, bool kSwap>
__device__ __forceinline__ T pp_load(const T* __restrict__ in, size_t i) {
  if (!kSwap) return in[i];
  if (sizeof(T) == 2) {
    const uint16_t u = reinterpret_cast<const uint16_t*>(in)[i];
    const uint16_t v = (uint16_t)((u >> 8) | (u << 8));
    T out;
    memcpy(&out, &v, 2);
    return out;
  } else if (sizeof(T) == 4) {
    const uint32_t v = __byte_perm(reinterpret_cast<const uint32_t*>(in)[i], 0u, 0x0123);
    T out;
    memcpy(&out, &v, 4);
    return out;
  } else if (sizeof(T) == 8) {
    const uint2 u = reinterpret_cast<const uint2*>(in)[i];
    const uint2 v = make_uint2(__byte_perm(u.y, 0u, 0x0123), __byte_perm(u.x, 0u, 0x0123));
    T out;
    memcpy(&out, &v, 8);
    return out;
  }
  return in[i];
}

template <typename T, bool kSwap>
__global__ void __launch_bounds__(256) pp_minmax_kernel(const T* __restrict__ in, size_t n, T* __restrict__ part_min,
                                                        T* __restrict__ part_max) {
  __shared__ T s_min[256], s_max[256];
  T mn = (T)0, mx = (T)0;  // the fold starts at T::zero()
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T x = pp_load<T, kSwap>(in, i);
    const double f = pp_to_f64(x);
    const bool fin = !(isinf(f) || isnan(f));
    if (x < mn && fin) mn = x;
    if (x > mx && fin) mx = x;
  }
  s_min[threadIdx.x] = mn;
  s_max[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      if (s_min[threadIdx.x + o] < s_min[threadIdx.x]) s_min[threadIdx.x] = s_min[threadIdx.x + o];
      if (s_max[threadIdx.x + o] > s_max[threadIdx.x]) s_max[threadIdx.x] = s_max[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part_min[blockIdx.x] = s_min[0];
    part_max[blockIdx.x] = s_max[0];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pp_finish_kernel(const T* __restrict__ part_min, const T* __restrict__ part_max,
                                                        int nparts, double* __restrict__ minmax) {
  __shared__ T s_min[256], s_max[256];
  T mn = (T)0, mx = (T)0;
  for (int i = threadIdx.x; i < nparts; i += 256) {
    if (part_min[i] < mn) mn = part_min[i];
    if (part_max[i] > mx) mx = part_max[i];
  }
  s_min[threadIdx.x] = mn;
  s_max[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      if (s_min[threadIdx.x + o] < s_min[threadIdx.x]) s_min[threadIdx.x] = s_min[threadIdx.x + o];
      if (s_max[threadIdx.x + o] > s_max[threadIdx.x]) s_max[threadIdx.x] = s_max[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    minmax[0] = pp_to_f64(s_min[0]);
    minmax[1] = pp_to_f64(s_max[0]);
  }
}

template <typename T, bool kSwap>
__global__ void __launch_bounds__(256) pp_map_kernel(const T* __restrict__ in, size_t n, const double* __restrict__ minmax,
                                                     double maxv, uint8_t* __restrict__ out) {
  const double mn = minmax[0], range = __dsub_rn(minmax[1], minmax[0]);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double f = pp_to_f64(pp_load<T, kSwap>(in, i));
    uint8_t v;
    const double a = fabs(f);
    const bool normal = a >= 2.2250738585072014e-308 && !isinf(f) && !isnan(f);  // f64::is_normal
    if (normal) {
      const double nv = __dmul_rn(__ddiv_rn(__dsub_rn(f, mn), range), maxv);
      v = (uint8_t)(int)nv;  // truncation towards zero; nv is in [0, MAX]
    } else if (isinf(f) && f > 0.0) {
      v = 0;    // ALWAYS_FILL
    } else {
      v = 255;  // NEVER_FILL
    }
    out[i] = v;
  }
}

template <typename T, bool kSwap>
static cudaError_t pp_run(const void* in, size_t n, uint32_t maxv, void* scratch, double* minmax, uint8_t* out,
                          cudaStream_t s) {
  // enough CTAs for a cube, few enough that a 2048^2 slice is not all launch overhead
  const int parts = (int)std::min<size_t>(1024, std::max<size_t>(1, (n + 4095) / 4096));
  T* pmin = (T*)scratch;
  T* pmax = pmin + 1024;
  pp_minmax_kernel<T, kSwap><<<parts, 256, 0, s>>>((const T*)in, n, pmin, pmax);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  pp_finish_kernel<T><<<1, 256, 0, s>>>(pmin, pmax, parts, minmax);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int maps = (int)std::min<size_t>((size_t)num_sms() * 8, std::max<size_t>(1, (n + 1023) / 1024));
  pp_map_kernel<T, kSwap><<<maps, 256, 0, s>>>((const T*)in, n, minmax, (double)maxv, out);
  return cudaGetLastError();
}

size_t pre_processor_scratch_bytes() { return 2 * 1024 * 8; }

cudaError_t launch_pre_processor(int dtype, const void* in, size_t n, uint32_t maxv, void* scratch, double* minmax,
                                 uint8_t* out, cudaStream_t s) {
  switch (dtype) {
    case 0: return pp_run<float, false>(in, n, maxv, scratch, minmax, out, s);
    case 1: return pp_run<double, false>(in, n, maxv, scratch, minmax, out, s);
    case 2: return pp_run<int32_t, false>(in, n, maxv, scratch, minmax, out, s);
    case 3: return pp_run<uint16_t, false>(in, n, maxv, scratch, minmax, out, s);
    case 4: return pp_run<int16_t, false>(in, n, maxv, scratch, minmax, out, s);
    case 5: return pp_run<uint8_t, false>(in, n, maxv, scratch, minmax, out, s);
    case 6: return pp_run<long long, false>(in, n, maxv, scratch, minmax, out, s);
    case 7: return pp_run<float, true>(in, n, maxv, scratch, minmax, out, s);      // big-endian (FITS BITPIX -32)
    case 8: return pp_run<double, true>(in, n, maxv, scratch, minmax, out, s);     // BITPIX -64
    case 9: return pp_run<int16_t, true>(in, n, maxv, scratch, minmax, out, s);    // BITPIX 16
    case 10: return pp_run<int32_t, true>(in, n, maxv, scratch, minmax, out, s);   // BITPIX 32
    case 11: return pp_run<long long, true>(in, n, maxv, scratch, minmax, out, s); // BITPIX 64
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ws
