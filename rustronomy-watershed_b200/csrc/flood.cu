// flood.cu -- K2: arrival times of all water levels in one persistent kernel.
//
// Replaces the reference's level loop x 'colouring_loop x find_flooded_px x write-back
// (lib.rs:1379-1438 / 1689-1748 and 196-257).  See common.cuh for the arrival-time
// encoding and DESIGN.md section 2 for why any relaxation order gives the reference's result.
//
// Kernel structure (sm_100a):
//   * persistent grid (one CTA slot per co-resident CTA), NO grid barrier: an asynchronous worklist of
//     active tiles, bucketed by the water level of the wake-up.  Every CTA always takes the tile with
//     the lowest bucket that is available, so tiles run roughly in the order of the reference's level
//     loop and rarely iterate on values a lower level overwrites later (a level-blind FIFO of sweeps
//     re-ran every tile ~9-12 times on smooth fields; this order needs ~4).  The flood ends when the
//     count of queued + in-flight entries reaches zero;
//   * every CTA = 8 consumer warps + 1 producer warp.  The producer claims tiles, stages each tile's
//     34 x 72-word box of arrival times and its 32 x 64 image bytes with two 2-D tensor copies
//     (cp.async.bulk.tensor.2d -> SASS UTMALDG, completion on an mbarrier) into a two-stage ring, and
//     publishes a finished tile's neighbours to the worklist after the stage has been refilled.  The
//     consumers never wait on a global round trip: they copy the stage into an odd-stride working tile
//     and iterate on it;
//   * the in-tile iteration alternates column and row ownership (8 pixels of Gauss-Seidel
//     along the phase's axis per step) until a phase changes nothing; a warp skips a phase when
//     neither its cells nor the cells next to them changed in the previous one;
//   * results go back with atomicMin: two CTAs may (rarely) hold the same tile at once, and arrival
//     times must never go up.
#include "kernels.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime, no libcuda link)

#include <cstring>

#ifndef WS_FLOOD_MINCTAS
#define WS_FLOOD_MINCTAS 3   // co-resident CTAs per SM the register budget is set for
#endif
#ifndef WS_FLOOD_BULK
#define WS_FLOOD_BULK 3
#endif

namespace ws {

// (mbarrier / bulk-copy helpers: common.cuh)

__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(FLOOD_CONSUMERS) : "memory");
}
__device__ __forceinline__ bool consumer_sync_or(bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "bar.red.or.pred p, 1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"((uint32_t)pred), "n"(FLOOD_CONSUMERS)
      : "memory");
  return out != 0u;
}

// ---------------------------------------------------------------------------
// state initialisation
// ---------------------------------------------------------------------------

// T = "never" and the image re-encoded for the flood: a pixel that can never flood -- not a window centre
// (lib.rs:220), above the last water level (filter (1), lib.rs:224, over levels 0..=max), or padding -- is
// stored as 255.  The label plane is NOT cleared: label_tile writes every word of it, and the only words read
// before that are the seeds' own (written by seed_init).
__global__ void __launch_bounds__(256) fill_state_kernel(FloodBuffers b, ImageDims d, const uint8_t* __restrict__ img,
                                                         uint32_t lmax) {
  if (ld_cg(&b.ctrl[FC_SEED_UNSORTED]) == 0u) return;  // fill_rows_kernel has done it
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // arrival times: the padded plane is a multiple of 8 words
  const size_t nt4 = d.t_plane() * d.n_img / 4;
  const uint4 inf4 = make_uint4(T_INF, T_INF, T_INF, T_INF);
  uint4* T4 = reinterpret_cast<uint4*>(b.T);
  for (size_t i = tid; i < nt4; i += stride) __stcg(T4 + i, inf4);
  // image bytes, four per thread (the padded rows are multiples of 64 bytes)
  const size_t pp = d.pix_plane();
  const int ppitch = d.pix_pitch();
  const size_t np4 = pp * d.n_img / 4;
  uchar4* P4 = reinterpret_cast<uchar4*>(b.pix);
  for (size_t i = tid; i < np4; i += stride) {
    const size_t e = i * 4;
    const int im = (int)(e / pp);
    const size_t rem = e - (size_t)im * pp;
    const int r = (int)(rem / ppitch), c0 = (int)(rem - (size_t)r * ppitch);
    uint32_t v[4] = {255u, 255u, 255u, 255u};
    if (r >= 1 && r <= d.rows - 2) {
      const uint8_t* row = img + (size_t)im * d.px_per_img() + (size_t)r * d.cols;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + k;
        if (c >= 1 && c <= d.cols - 2) {
          const uint32_t x = __ldg(row + c);
          v[k] = x > lmax ? 255u : x;
        }
      }
    }
    P4[i] = make_uchar4((unsigned char)v[0], (unsigned char)v[1], (unsigned char)v[2], (unsigned char)v[3]);
  }
}

constexpr uint32_t TILE_NONE_U = 0xFFFFFFFFu;
__device__ __forceinline__ unsigned long long ld_cg64(const unsigned long long* p) { return __ldcg(p); }

__device__ __forceinline__ uint32_t flood_bucket(uint32_t level, int shift) {
  const uint32_t b = level >> shift;
  return b < (uint32_t)FLOOD_BUCKETS ? b : (uint32_t)FLOOD_BUCKETS - 1u;
}

// Wake a tile, part 1: mark it dirty; returns true when the caller must append an entry to ring
// `bucket` (no entry of the same or a better priority is queued).  The caller has made the data the
// tile must see visible (fence) before this call; whoever takes an entry clears its bucket bit and
// the dirty bit in one atomic and loads the tile afterwards.  kQuiescent: no flood kernel is running
// (seeding, strip import), so a plain load may decide that there is nothing to do; inside the flood
// the decision must come from an atomic on the mask word -- a read-modify-write is ordered against the
// taker's atomicAnd, a load is not.
template <bool kQuiescent>
__device__ __forceinline__ bool push_mark(const FloodBuffers& b, uint32_t tile, uint32_t bucket) {
  const unsigned long long bit = 1ull << bucket, not_worse = (bit << 1) - 1ull;
  unsigned long long cur = 0ull;  // optimistic: the tile is idle (neither queued nor dirty)
  if (kQuiescent) {
    cur = ld_cg64(&b.qmask[tile]);
    if ((cur & not_worse) && (cur & Q_DIRTY)) return false;  // the common case when seeds are dense
  }
  for (;;) {
    const bool queued = (cur & not_worse) != 0ull;
    const unsigned long long want = cur | Q_DIRTY | (queued ? 0ull : bit);
    if (want == cur) return false;
    const unsigned long long old = atomicCAS(&b.qmask[tile], cur, want);
    if (old == cur) return !queued;
    cur = old;
  }
}
// part 2: reserve a slot, write it, raise the semaphore.  No fence between the two: a taker that wins a
// slot before its tile id has landed simply waits for it (the slot holds Q_EMPTY until then).
// The slot must be free: the ring holds tiles + FLOOD_QSLACK entries, at most one live entry per tile and
// bucket plus the claims in flight -- but a claimer stalled between taking its head index and reading the
// slot for a whole lap of the ring would be overwritten silently.  The exchange makes that loud (error bit 5)
// instead of losing a tile; its result is not waited for before the semaphore is raised.
__device__ __forceinline__ void push_append(const FloodBuffers& b, uint32_t tile, uint32_t bucket) {
  const uint32_t t = atomicAdd(&b.ctrl[FC_QTAIL0 + bucket], 1u);
  const uint32_t old = atomicExch(&b.qslots[(size_t)bucket * b.qcap + t % b.qcap], tile);
  atomicAdd(&b.ctrl[FC_QAVAIL0 + bucket], 1u);
  if (old != Q_EMPTY) atomicOr(&b.ctrl[FC_ERROR], 32u);
}
template <bool kQuiescent>
__device__ __forceinline__ void push_tile(const FloodBuffers& b, uint32_t tile, uint32_t bucket) {
  if (push_mark<kQuiescent>(b, tile, bucket)) {
    atomicAdd(&b.ctrl[FC_OUTSTANDING], 1u);
    push_append(b, tile, bucket);
  }
}

// ---- sorted seed lists -----------------------------------------------------------------------------
// find_local_minima returns its seeds in row-major order without repeats.  For such a list nothing has to be
// scattered: a CTA per image row writes the row of T once (INF, then 0 at the row's seeds while the lines are
// still in L2), the colour of a seed is its position in the list -- label_tile derives it from rowbase[] and the
// seeds to its left in the tile row -- and the label plane is not touched at all.  (The general path below
// rewrites nearly every 32-byte sector of T and of the label plane a second time: 3.1 GB of DRAM traffic for
// 29 M seeds.)  A list that is not strictly ascending, or has a seed outside the image, takes the general path.

__device__ __forceinline__ int seed_slice_of(const uint32_t* __restrict__ seed_off, int n_img, uint32_t i) {
  int lo = 0, hi = n_img;  // last b with seed_off[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seed_off + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// sets FC_SEED_UNSORTED on a violation; writes row_start[] (meaningful only when the list is sorted)
__global__ void __launch_bounds__(256) seeds_scan_kernel(FloodBuffers b, ImageDims d,
                                                         const uint32_t* __restrict__ seeds_rc,
                                                         const uint32_t* __restrict__ seed_off, uint32_t nseeds) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nseeds; i += stride) {
    const int img = seed_slice_of(seed_off, d.n_img, i);
    const uint2 rc = __ldg(reinterpret_cast<const uint2*>(seeds_rc) + i);
    const uint32_t first = __ldg(seed_off + img), end = __ldg(seed_off + img + 1);
    bool bad = rc.x >= (uint32_t)d.rows || rc.y >= (uint32_t)d.cols;
    int prev_row = -1;
    if (i > first) {
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(seeds_rc) + i - 1);
      bad |= !(q.x < rc.x || (q.x == rc.x && q.y < rc.y));
      prev_row = (int)min(q.x, (uint32_t)d.rows - 1u);
    }
    if (bad) {
      st_cg(&b.ctrl[FC_SEED_UNSORTED], 1u);
      continue;
    }
    uint32_t* rs = b.row_start + (size_t)img * d.rows;
    for (int r = prev_row + 1; r <= (int)rc.x; ++r) rs[r] = i;        // rows without seeds point at the next seed
    if (i + 1 == end)
      for (int r = (int)rc.x + 1; r < d.rows; ++r) rs[r] = end;
  }
}
// slices without a single seed, and the sentinel behind the last row
__global__ void __launch_bounds__(256) seeds_scan_empty_kernel(FloodBuffers b, ImageDims d,
                                                               const uint32_t* __restrict__ seed_off, uint32_t nseeds) {
  const int img = blockIdx.x;
  if (img == 0 && threadIdx.x == 0) b.row_start[(size_t)d.n_img * d.rows] = nseeds;
  const uint32_t first = __ldg(seed_off + img);
  if (first != __ldg(seed_off + img + 1)) return;
  for (int r = threadIdx.x; r < d.rows; r += blockDim.x) b.row_start[(size_t)img * d.rows + r] = first;
}

// One CTA per row of the padded arrival-time plane.  Everything the row needs comes from a bitmap of its seeds'
// columns in shared memory: the words of T (written once: INF, or 0 under a set bit), rowbase (prefix popcounts
// of the bitmap words: the index of the first seed at or right of every 32-column block), and the tiles to queue
// (the two words of a tile are not both zero; a seed in a tile's first / last column or row also concerns the
// neighbour).  Rows wider than FR_COLS columns take several passes.
__global__ void __launch_bounds__(256) fill_rows_kernel(FloodBuffers b, ImageDims d, const uint8_t* __restrict__ img,
                                                        uint32_t lmax, const uint32_t* __restrict__ seeds_rc) {
  if (ld_cg(&b.ctrl[FC_SEED_UNSORTED]) != 0u) return;
  constexpr int FR_COLS = 32768;                    // image columns per pass: 4 KB of bitmap
  constexpr int FR_WORDS = FR_COLS / 32;
  __shared__ uint32_t s_bits[FR_WORDS];
  __shared__ uint32_t s_warp[8], s_total;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t_rows = d.t_rows(), pitch = d.t_pitch();
  const int im = blockIdx.x / t_rows, prow = blockIdx.x - im * t_rows;
  const int r = prow - 1;  // image row of this padded row
  uint4* T4 = reinterpret_cast<uint4*>(b.T + (size_t)im * d.t_plane() + (size_t)prow * pitch);
  const uint4 inf4 = make_uint4(T_INF, T_INF, T_INF, T_INF);
  const bool image_row = r >= 0 && r < d.rows;
  uint32_t lo = 0, hi = 0;
  if (image_row) {
    lo = b.row_start[(size_t)im * d.rows + r];
    hi = b.row_start[(size_t)im * d.rows + r + 1];
  }
  const int wcols = d.tiles_x * TILE_W;             // image columns the plane has room for (uint4 1 .. wcols / 4)
  if (tid == 0) {                                   // the pad words left and right of them
    __stcg(T4, inf4);
    __stcg(T4 + wcols / 4 + 1, inf4);
  }
  if (!image_row) {
    for (int j = tid; j < wcols / 4; j += blockDim.x) __stcg(T4 + 1 + j, inf4);
  } else {
    const int ty = r / TILE_H;
    const uint32_t tile0 = (uint32_t)im * d.tiles_per_img() + (uint32_t)ty * d.tiles_x;
    uint32_t* rb = b.rowbase + ((size_t)im * d.rows + r) * (2 * d.tiles_x);
    uint32_t passbase = lo;                         // seeds of the row left of this pass
    for (int cb = 0; cb < wcols; cb += FR_COLS) {
      const int np = min(FR_COLS, wcols - cb);      // a multiple of 64
      const int nw = np / 32;
      for (int i = tid; i < nw; i += blockDim.x) s_bits[i] = 0u;
      __syncthreads();
      for (uint32_t j = lo + tid; j < hi; j += blockDim.x) {
        const int c = (int)__ldg(seeds_rc + 2 * (size_t)j + 1) - cb;
        if (c >= 0 && c < np) atomicOr(&s_bits[c >> 5], 1u << (c & 31));
      }
      __syncthreads();
      // arrival times
      for (int j = tid; j < np / 4; j += blockDim.x) {
        const uint32_t m = (s_bits[j >> 3] >> ((j & 7) * 4)) & 0xFu;
        __stcg(T4 + 1 + cb / 4 + j, m == 0u ? inf4
                                             : make_uint4((m & 1u) ? 0u : T_INF, (m & 2u) ? 0u : T_INF,
                                                          (m & 4u) ? 0u : T_INF, (m & 8u) ? 0u : T_INF));
      }
      // rowbase: exclusive prefix popcount over the bitmap words (4 consecutive words per thread)
      {
        uint32_t w[4], mine = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = tid * 4 + k;
          w[k] = i < nw ? (uint32_t)__popc(s_bits[i]) : 0u;
          mine += w[k];
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (int k = 0; k < warp; ++k) before += s_warp[k];
        if (tid == 255) s_total = before + incl;
        uint32_t run = passbase + before + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = tid * 4 + k;
          if (i < nw) rb[cb / 32 + i] = run;
          run += w[k];
        }
      }
      // tiles to queue (red-black order as in seed_init; the flood is a later launch)
      for (int t = tid; t < np / TILE_W; t += blockDim.x) {
        const uint32_t w0 = s_bits[2 * t], w1 = s_bits[2 * t + 1];
        if ((w0 | w1) == 0u) continue;
        const int tx = cb / TILE_W + t;
        const uint32_t tile = tile0 + (uint32_t)tx;
        const uint32_t par = (uint32_t)(tx + ty) & 1u;
        push_tile<true>(b, tile, par);
        if (r % TILE_H == 0 && ty > 0) push_tile<true>(b, tile - d.tiles_x, par ^ 1u);
        if (r % TILE_H == TILE_H - 1 && ty + 1 < d.tiles_y) push_tile<true>(b, tile + d.tiles_x, par ^ 1u);
        if ((w0 & 1u) && tx > 0) push_tile<true>(b, tile - 1, par ^ 1u);
        if ((w1 >> 31) && tx + 1 < d.tiles_x) push_tile<true>(b, tile + 1, par ^ 1u);
      }
      __syncthreads();
      passbase += s_total;
      __syncthreads();
    }
  }
  if (r < 0 || r >= d.pix_rows()) return;
  {  // the image row re-encoded for the flood: 255 = never floods (border, above the last level, padding)
    const int ppitch = d.pix_pitch();
    uchar4* P4 = reinterpret_cast<uchar4*>(b.pix + (size_t)im * d.pix_plane() + (size_t)r * ppitch);
    const bool inner = r >= 1 && r <= d.rows - 2;
    const uint8_t* row = img + (size_t)im * d.px_per_img() + (size_t)(inner ? r : 0) * d.cols;
    const bool words = inner && (d.cols & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 3u) == 0;
    for (int i = tid; i < ppitch / 4; i += blockDim.x) {
      uint32_t v[4] = {255u, 255u, 255u, 255u};
      if (inner) {
        if (words && 4 * i + 3 < d.cols) {
          const uint32_t x = __ldg(reinterpret_cast<const uint32_t*>(row) + i);
          v[0] = x & 0xFFu; v[1] = (x >> 8) & 0xFFu; v[2] = (x >> 16) & 0xFFu; v[3] = x >> 24;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (4 * i + k < d.cols) v[k] = __ldg(row + 4 * i + k);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 4 * i + k;
          v[k] = (c >= 1 && c <= d.cols - 2 && v[k] <= lmax) ? v[k] : 255u;
        }
      }
      P4[i] = make_uchar4((unsigned char)v[0], (unsigned char)v[1], (unsigned char)v[2], (unsigned char)v[3]);
    }
  }
}

cudaError_t launch_fill_state(FloodBuffers b, ImageDims d, const uint8_t* img, uint32_t lmax, const uint32_t* seeds_rc,
                              const uint32_t* seed_off, uint32_t nseeds, int sms, cudaStream_t s) {
  // empty worklist: all ring slots unwritten, no tile queued, counters zero
  cudaError_t e = cudaMemsetAsync(b.qslots, 0xFF, sizeof(uint32_t) * (size_t)FLOOD_BUCKETS * b.qcap, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(b.qmask, 0, sizeof(unsigned long long) * (size_t)d.tiles_total(), s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(b.ctrl, 0, sizeof(uint32_t) * FC_WORDS, s);
  if (e != cudaSuccess) return e;
  if (nseeds) {
    const uint32_t want = (nseeds + 255) / 256, cap = (uint32_t)sms * 32u;
    seeds_scan_kernel<<<want < cap ? want : cap, 256, 0, s>>>(b, d, seeds_rc, seed_off, nseeds);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  seeds_scan_empty_kernel<<<d.n_img, 256, 0, s>>>(b, d, seed_off, nseeds);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fill_rows_kernel<<<d.n_img * d.t_rows(), 256, 0, s>>>(b, d, img, lmax, seeds_rc);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  fill_state_kernel<<<sms * 16, 256, 0, s>>>(b, d, img, lmax);   // (returns at once unless the list is unsorted)
  return cudaGetLastError();
}

// Colour the starting pixels (lib.rs:1365-1367): T = 0, colour = index + 1, a later duplicate overwrites an
// earlier one (sequential loop) == the largest index wins.  The label plane is not cleared beforehand, so the
// colour is a plain store; a position that occurs twice shows up as an exchange on T that returns 0, and
// only then seed_dup_kernel settles the largest index with atomicMax (every word it touches already holds
// the colour of one of the duplicates).
__global__ void __launch_bounds__(256) seed_init_kernel(FloodBuffers b, ImageDims d,
                                                        const uint32_t* __restrict__ seeds_rc,
                                                        const uint32_t* __restrict__ seed_off, uint32_t nseeds,
                                                        uint32_t colour_base) {
  if (ld_cg(&b.ctrl[FC_SEED_UNSORTED]) == 0u) return;  // sorted list: fill_rows_kernel has placed the seeds
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t n_pad = (nseeds + 31u) & ~31u;         // whole warps stay together for the match below
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
    uint32_t tile = TILE_NONE_U, par = 0;
    uint32_t r = 0, c = 0;
    int ty = 0, tx = 0;
    if (i < nseeds) {
      const int img = seed_slice_of(seed_off, d.n_img, i);
      const uint2 rc = __ldg(reinterpret_cast<const uint2*>(seeds_rc) + i);
      r = rc.x;
      c = rc.y;
      if (r >= (uint32_t)d.rows || c >= (uint32_t)d.cols) {
        atomicOr(&b.ctrl[FC_ERROR], 1u);
      } else {
        const uint32_t old = atomicExch(&b.T[(size_t)img * d.t_plane() + d.t_index((int)r, (int)c)], 0u);
        st_cg(&b.lab[(size_t)img * d.px_per_img() + (size_t)r * d.cols + c],
              LAB_RESOLVED | (colour_base + i - __ldg(seed_off + img) + 1u));
        if (old == 0u) st_cg(&b.ctrl[FC_SEED_DUP], 1u);
        ty = r / TILE_H;
        tx = c / TILE_W;
        tile = (uint32_t)img * d.tiles_per_img() + ty * d.tiles_x + tx;
        // Red-black order over the tiles: the even tiles (tx + ty even) start in bucket 0, the odd ones in
        // bucket 1 -- they then already see their neighbours' results, a Gauss-Seidel step at tile level that
        // saves re-activations.
        par = (uint32_t)(tx + ty) & 1u;
      }
    }
    // Lists from find_local_minima are row-major, so the 32 seeds of a warp share a handful of tiles: one lane
    // per distinct tile queues it.  (the flood is a later launch: no fence needed here)
    const uint32_t peers = __match_any_sync(0xffffffffu, tile);
    if (tile != TILE_NONE_U) {
      if ((int)(__ffs((int)peers) - 1) == (int)(threadIdx.x & 31)) push_tile<true>(b, tile, par);
      // a seed on the tile's edge is in the halo of the neighbouring tile: that tile must look as well
      if (r % TILE_H == 0 && ty > 0) push_tile<true>(b, tile - d.tiles_x, par ^ 1u);
      if (r % TILE_H == TILE_H - 1 && ty + 1 < d.tiles_y) push_tile<true>(b, tile + d.tiles_x, par ^ 1u);
      if (c % TILE_W == 0 && tx > 0) push_tile<true>(b, tile - 1, par ^ 1u);
      if (c % TILE_W == TILE_W - 1 && tx + 1 < d.tiles_x) push_tile<true>(b, tile + 1, par ^ 1u);
    }
  }
}

// Only when seed_init saw a position twice: the largest index wins (lib.rs:1365-1367 is a sequential loop).
__global__ void __launch_bounds__(256) seed_dup_kernel(FloodBuffers b, ImageDims d,
                                                       const uint32_t* __restrict__ seeds_rc,
                                                       const uint32_t* __restrict__ seed_off, uint32_t nseeds,
                                                       uint32_t colour_base) {
  if (ld_cg(&b.ctrl[FC_SEED_UNSORTED]) == 0u || ld_cg(&b.ctrl[FC_SEED_DUP]) == 0u) return;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nseeds; i += stride) {
    int lo = 0, hi = d.n_img;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(seed_off + mid) <= i) lo = mid; else hi = mid;
    }
    const uint32_t r = seeds_rc[2 * (size_t)i], c = seeds_rc[2 * (size_t)i + 1];
    if (r >= (uint32_t)d.rows || c >= (uint32_t)d.cols) continue;
    atomicMax(&b.lab[(size_t)lo * d.px_per_img() + (size_t)r * d.cols + c],
              LAB_RESOLVED | (colour_base + i - __ldg(seed_off + lo) + 1u));
  }
}

cudaError_t launch_seed_init(FloodBuffers b, ImageDims d, const uint32_t* seeds_rc, const uint32_t* seed_off,
                             uint32_t nseeds, uint32_t colour_base, cudaStream_t s) {
  if (nseeds == 0) return cudaSuccess;
  const uint32_t want = (nseeds + 255) / 256, cap = (uint32_t)num_sms() * 32u;
  seed_init_kernel<<<want < cap ? want : cap, 256, 0, s>>>(b, d, seeds_rc, seed_off, nseeds, colour_base);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  seed_dup_kernel<<<want < cap ? want : cap, 256, 0, s>>>(b, d, seeds_rc, seed_off, nseeds, colour_base);
  return cudaGetLastError();
}

// (usize, usize) pairs -> u32 pairs; anything outside the image becomes 0xFFFFFFFF so that
// seed_init flags it (the reference panics on an out-of-bounds seed, lib.rs:1366 / 1676).
__global__ void __launch_bounds__(256) seeds_convert_kernel(const uint64_t* __restrict__ in,
                                                            uint32_t* __restrict__ out, size_t n2, uint64_t rows,
                                                            uint64_t cols) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const uint64_t v = in[i];
    const uint64_t lim = (i & 1) ? cols : rows;
    out[i] = v < lim ? (uint32_t)v : 0xFFFFFFFFu;
  }
}

cudaError_t launch_seeds_convert(const uint64_t* in, uint32_t* out, size_t nseeds, size_t rows, size_t cols,
                                 cudaStream_t s) {
  if (nseeds == 0) return cudaSuccess;
  const size_t n2 = 2 * nseeds;
  const size_t want = (n2 + 255) / 256;
  const size_t gcap = (size_t)num_sms() * 8;
  const unsigned grid = (unsigned)(want < gcap ? want : gcap);
  seeds_convert_kernel<<<grid, 256, 0, s>>>(in, out, n2, rows, cols);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// the flood kernel
// ---------------------------------------------------------------------------

struct FloodArgs {
  alignas(64) CUtensorMap tmT;  // arrival times: [n_img * t_rows][t_pitch] u32, box 72 x 34
  alignas(64) CUtensorMap tmP;  // image bytes:   [n_img * pix_rows][pix_pitch] u8, box 64 x 32
  FloodBuffers b;
  ImageDims d;
  int check_overflow;
  int bucket_shift;  // worklist bucket = wake-up level >> bucket_shift
};

// one instruction per box: 2D tiled tensor copy global -> shared, completes on `bar`
__device__ __forceinline__ void tensor_g2s(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

enum { DIR_UP = 0, DIR_DOWN = 1, DIR_LEFT = 2, DIR_RIGHT = 3 };
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;
constexpr uint32_t TILE_NONE = 0xFFFFFFFFu;

// Stages of the producer -> consumer ring.  (Measured: a third stage does not shorten the consumers'
// waits -- they wait on the producer's throughput, not on its latency -- and holding more claimed tiles
// per CTA costs re-activations on smooth fields.)
constexpr int FLOOD_STAGES = 2;

struct FloodStage {  // (tensor copies want 128-byte aligned destinations)
  alignas(128) uint32_t T[STG_H * STG_W];      // staged arrival times (also the "before" image of the tile)
  alignas(128) uint8_t pix[TILE_H * TILE_W];   // staged image bytes
};

struct __align__(128) FloodSmem {
  FloodStage st[FLOOD_STAGES];
  uint32_t W[SM_H * SM_W];          // working tile incl. halo
  size_t tbase[FLOOD_STAGES];       // per stage: word index of the tile's pixel (0, 0) in the padded arrival times
  uint64_t full[FLOOD_STAGES], empty[FLOOD_STAGES];  // mbarriers of the ring
  uint32_t tile[FLOOD_STAGES];
  uint32_t dirty[3];                // cells changed in a phase (rotating: written, read, cleared)
  uint32_t key[FLOOD_STAGES][4];               // per stage and direction: smallest value a changed edge pixel offers
                                    // the pixel facing it (KEY_NONE: that neighbour need not re-run)
};

// A(p) = (img[p] << 24) | 1.  A pixel that can never flood is stored as 255 (fill_state): its A = 0xFF000001 lies
// above T_INF = 0xFF000000, the value its arrival time starts with, so no relaxation ever lowers it -- no special
// case needed.
__device__ __forceinline__ uint32_t flood_A(uint32_t pix) { return (pix << 24) | 1u; }

// one Gauss-Seidel step of a sweep: T(p) <- min(T(p), max(A(p), 1 + min(m, prev, next)))   (3 instructions)
__device__ __forceinline__ uint32_t flood_relax(uint32_t t, uint32_t m, uint32_t prev, uint32_t nb, uint32_t A) {
  return min(t, __viaddmax_u32(__vimin3_u32(m, prev, nb), 1u, A));
}

// One phase of a thread: two Gauss-Seidel sweeps (forwards, then backwards) over its 8 pixels along one axis
// (ALONG = word stride between them, ACROSS = stride to the two neighbours off the axis).  Returns whether
// any of the 8 changed.
template <int ALONG, int ACROSS>
__device__ __forceinline__ bool flood_phase(uint32_t* base, const uint32_t (&A)[ROWS_PER_THREAD]) {
  uint32_t t[ROWS_PER_THREAD], o[ROWS_PER_THREAD], m[ROWS_PER_THREAD];
  const uint32_t lo = base[-ALONG], hi = base[ROWS_PER_THREAD * ALONG];
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    t[i] = o[i] = base[i * ALONG];
    m[i] = min(base[i * ALONG - ACROSS], base[i * ALONG + ACROSS]);
  }
  uint32_t prev = lo;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    t[i] = flood_relax(t[i], m[i], prev, (i < ROWS_PER_THREAD - 1) ? t[i + 1] : hi, A[i]);
    prev = t[i];
  }
  prev = hi;
#pragma unroll
  for (int i = ROWS_PER_THREAD - 1; i >= 0; --i) {
    t[i] = flood_relax(t[i], m[i], prev, (i > 0) ? t[i - 1] : lo, A[i]);
    prev = t[i];
  }
  uint32_t x = 0u;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) x |= t[i] ^ o[i];
  if (x == 0u) return false;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) base[i * ALONG] = t[i];
  return true;
}

// cells (bit = cell row * 8 + cell column, 4 x 8 cells) and their four neighbours
__device__ __forceinline__ uint32_t cells_dilate(uint32_t m) {
  return m | ((m & ~0x80808080u) << 1) | ((m & ~0x01010101u) >> 1) | (m << 8) | (m >> 8);
}

// ---- consumer side: one tile to its local fixed point ------------------------------------
__device__ __forceinline__ void flood_consume(const FloodArgs& a, FloodSmem& sm, int s) {
  const ImageDims& d = a.d;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const FloodStage& st = sm.st[s];

  // stage -> working tile (odd stride).  Staged row = image columns c0-4 .. c0+67; the working
  // tile keeps c0-1 .. c0+64.
  for (int lr = warp; lr < SM_H; lr += FLOOD_CONSUMERS / 32) {   // a warp per row: no index arithmetic per word
    const uint32_t* src = st.T + lr * STG_W + (T_PAD_L - 1);
    uint32_t* dst = sm.W + lr * SM_W;
    dst[lane] = src[lane];
    dst[lane + 32] = src[lane + 32];
    if (lane < TILE_W + 2 - 64) dst[lane + 64] = src[lane + 64];
  }

  // column ownership: column cl, rows cgp*8 .. cgp*8+7;  row ownership: row rl, columns rgp*8 .. +7
  const int cl = tid % TILE_W, cgp = tid / TILE_W;
  uint32_t* colp = sm.W + (cgp * ROWS_PER_THREAD + 1) * SM_W + cl + 1;
  const int rl = lane, rgp = warp;
  uint32_t* rowp = sm.W + (rl + 1) * SM_W + rgp * ROWS_PER_THREAD + 1;
  uint32_t Ac[ROWS_PER_THREAD], Ar[ROWS_PER_THREAD];
  {
    // straight from the staged image bytes (dense rows of TILE_W): a thread's eight row-phase pixels are two
    // aligned words, one byte permute each
    const uint32_t* prow = reinterpret_cast<const uint32_t*>(st.pix + rl * TILE_W + rgp * ROWS_PER_THREAD);
    const uint32_t w0 = prow[0], w1 = prow[1];
    Ar[0] = __byte_perm(w0, 1u, 0x0554); Ar[1] = __byte_perm(w0, 1u, 0x1554);
    Ar[2] = __byte_perm(w0, 1u, 0x2554); Ar[3] = __byte_perm(w0, 1u, 0x3554);
    Ar[4] = __byte_perm(w1, 1u, 0x0554); Ar[5] = __byte_perm(w1, 1u, 0x1554);
    Ar[6] = __byte_perm(w1, 1u, 0x2554); Ar[7] = __byte_perm(w1, 1u, 0x3554);
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) Ac[i] = flood_A(st.pix[(cgp * ROWS_PER_THREAD + i) * TILE_W + cl]);
  }

  uint32_t nphase = 0;
  // Dirty cells.  The tile is 4 x 8 cells of 8 x 8 pixels (bit = cell row * 8 + cell column).  A warp
  // iterates in a phase only if one of its four cells, or a cell next to them, changed in the previous
  // phase -- otherwise every pixel it owns was verified against unchanged neighbours the last time it ran.
  // After the first few phases of a tile, and in re-activations that only bring in one edge, most warps skip.
  const uint32_t col_cells = 0xFu << (cgp * 8 + (warp & 1) * 4);  // column phase: rows cgp*8.., 32 columns
  const uint32_t col_bit = 1u << (cgp * 8 + cl / 8);
  const uint32_t row_cells = 0x01010101u << warp;                  // row phase: 32 rows, columns warp*8..
  const uint32_t row_bit = 1u << ((rl / 8) * 8 + rgp);
  // "one of my cells or a cell next to them changed" == the changed cells meet my cells dilated by one cell
  const uint32_t col_near = cells_dilate(col_cells), row_near = cells_dilate(row_cells);
  uint32_t dm = 0xFFFFFFFFu;  // cells that changed in the previous phase
  uint32_t ever = 0u;         // cells that changed in any phase
  if (tid == 0) sm.dirty[0] = sm.dirty[1] = sm.dirty[2] = 0u;
  consumer_sync();
  // dirty words rotate: written in this phase (read after its barrier) / read in the previous phase / cleared now
  for (int k = 0;; k = (k == 2) ? 0 : k + 1) {
    uint32_t* dw = &sm.dirty[k];
    // ---- column phase: down then up along the thread's 8 rows ----
    {
      ++nphase;
      bool changed = false;
      if (dm & col_near) changed = flood_phase<SM_W, 1>(colp, Ac);
      const uint32_t wm = __reduce_or_sync(0xffffffffu, changed ? col_bit : 0u);
      if (lane == 0 && wm) atomicOr(dw, wm);
      if (tid == 0) sm.dirty[k == 2 ? 0 : k + 1] = 0u;  // last read after the barrier of the previous phase
      if (!consumer_sync_or(changed)) break;
      dm = *dw;
      ever |= dm;
    }
    k = (k == 2) ? 0 : k + 1;
    dw = &sm.dirty[k];
    // ---- row phase: right then left along its 8 columns ----
    {
      ++nphase;
      bool changed = false;
      if (dm & row_near) changed = flood_phase<1, SM_W>(rowp, Ar);
      const uint32_t wm = __reduce_or_sync(0xffffffffu, changed ? row_bit : 0u);
      if (lane == 0 && wm) atomicOr(dw, wm);
      if (tid == 0) sm.dirty[k == 2 ? 0 : k + 1] = 0u;
      if (!consumer_sync_or(changed)) break;
      dm = *dw;
      ever |= dm;
    }
  }

  // write back what changed against the staged copy (column ownership: coalesced along rows)
  uint32_t* Tg = a.b.T + sm.tbase[s] + (size_t)(cgp * ROWS_PER_THREAD) * d.t_pitch() + cl;
  const int tp = d.t_pitch();
  uint32_t v[ROWS_PER_THREAD], chm = 0u;
  if (ever & col_bit) {  // else: nothing in the 8 x 8 cell of my pixels ever changed
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      v[i] = colp[i * SM_W];
      if (v[i] != st.T[(cgp * ROWS_PER_THREAD + i + 1) * STG_W + cl + T_PAD_L]) chm |= 1u << i;
    }
  }
  if (chm) {
    // A hop counter that ran past 2^24 - 1 carries into the level and leaves hop == 0 behind: the pixel
    // where that happens keeps this value (everything after it is built on it), so looking at the values
    // that are written back is enough -- no test inside the relaxation.
    bool ovf = false;
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i)
      if (chm & (1u << i)) {
        atomicMin(Tg + (size_t)i * tp, v[i]);  // result unused: a fire-and-forget RED.MIN (measured: same time as st)
        ovf |= (v[i] & HOP_MASK) == 0u;
      }
    if (a.check_overflow && ovf) atomicOr(&a.b.ctrl[FC_ERROR], 2u);
    // A neighbour tile re-runs only if a changed edge pixel can still lower the pixel facing it:
    // T(edge) + 1 < T(facing pixel), the latter as staged in our halo (never newer than the truth, so the
    // test never drops a needed wake-up).  Without it every tile woke all four neighbours, including the
    // one its values came from, and most activations were such echoes.  The smallest such offer per
    // direction is the priority of the wake-up.  Only the threads that own edge pixels look.
    if (cl == 0 || cl == TILE_W - 1) {
      const int side = cl == 0 ? -1 : 1;
      uint32_t k = KEY_NONE;
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i)
        if ((chm & (1u << i)) && v[i] + 1u < colp[i * SM_W + side]) k = min(k, v[i] + 1u);
      if (k != KEY_NONE) atomicMin(&sm.key[s][cl == 0 ? DIR_LEFT : DIR_RIGHT], k);
    }
    if (cgp == 0 && (chm & 1u) && v[0] + 1u < colp[-SM_W]) atomicMin(&sm.key[s][DIR_UP], v[0] + 1u);
    if (cgp == TILE_H / ROWS_PER_THREAD - 1 && (chm & (1u << (ROWS_PER_THREAD - 1))) &&
        v[ROWS_PER_THREAD - 1] + 1u < colp[ROWS_PER_THREAD * SM_W])
      atomicMin(&sm.key[s][DIR_DOWN], v[ROWS_PER_THREAD - 1] + 1u);
  }
  if (tid == 0) atomicAdd(&a.b.ctrl[FC_PHASES], nphase);
}

// ---- producer side: the worklist -----------------------------------------------------------

constexpr int POP_MAX = 4;  // tiles one claim may take when the worklist is long


// Claim tiles from the best (lowest) non-empty bucket (warp-collective): up to POP_MAX when that
// bucket holds far more entries than there are CTAs -- the claim is a chain of five dependent L2 round
// trips, and on noise fields, where every tile is queued at once, that chain, not the iteration, would
// bound the kernel -- otherwise one.  Returns the number claimed (0: nothing available right now);
// lane j < n holds the j-th tile in `mine`.
__device__ __forceinline__ int flood_pop(const FloodBuffers& b, int lane, uint32_t& mine) {
  for (;;) {
    const int a0 = (int)ld_poll(&b.ctrl[FC_QAVAIL0 + lane]);
    const int a1 = (int)ld_poll(&b.ctrl[FC_QAVAIL0 + 32 + lane]);
    const uint32_t m0 = __ballot_sync(0xffffffffu, a0 > 0), m1 = __ballot_sync(0xffffffffu, a1 > 0);
    if ((m0 | m1) == 0u) return 0;
    const int bk = m0 ? __ffs((int)m0) - 1 : 31 + __ffs((int)m1);
    const int av = __shfl_sync(0xffffffffu, bk < 32 ? a0 : a1, bk & 31);
    int want = av / (int)(2u * gridDim.x);
    want = want < 1 ? 1 : (want > POP_MAX ? POP_MAX : want);
    int got = 0;
    uint32_t h0 = 0;
    if (lane == 0) {
      // the semaphore guarantees a reserved slot for every unit of a successful decrement
      const int old = (int)atomicSub(&b.ctrl[FC_QAVAIL0 + bk], (uint32_t)want);
      got = old <= 0 ? 0 : (old < want ? old : want);
      if (got < want) atomicAdd(&b.ctrl[FC_QAVAIL0 + bk], (uint32_t)(want - got));  // lost (part of) the race
      if (got) h0 = atomicAdd(&b.ctrl[FC_QHEAD0 + bk], (uint32_t)got);
    }
    got = __shfl_sync(0xffffffffu, got, 0);
    h0 = __shfl_sync(0xffffffffu, h0, 0);
    if (got == 0) continue;
    uint32_t v = Q_EMPTY;
    bool valid = false;
    if (lane < got) {
      uint32_t* slot = &b.qslots[(size_t)bk * b.qcap + (h0 + (uint32_t)lane) % b.qcap];
      v = ld_poll(slot);
      // the slot may belong to a push that has reserved it and is about to write it
      for (uint32_t spin = 0; v == Q_EMPTY && spin < (1u << 22); ++spin) v = ld_poll(slot);
      if (v == Q_EMPTY) {
        atomicOr(&b.ctrl[FC_ERROR], 8u);
      } else {
        st_cg(slot, Q_EMPTY);
        const unsigned long long old = atomicAnd(&b.qmask[v], ~((1ull << bk) | Q_DIRTY));
        valid = (old & Q_DIRTY) != 0ull;  // else: already served through a better bucket since it was queued
      }
    }
    const uint32_t vm = __ballot_sync(0xffffffffu, valid);
    const int nvalid = __popc(vm);
    if (lane == 0 && nvalid < got) {
      atomicSub(&b.ctrl[FC_OUTSTANDING], (uint32_t)(got - nvalid));
      atomicAdd(&b.ctrl[FC_STALE], (uint32_t)(got - nvalid));
    }
    if (nvalid == 0) continue;
    mine = TILE_NONE;
#pragma unroll
    for (int j = 0; j < POP_MAX; ++j) {
      const uint32_t src = __fns(vm, 0, j + 1);  // lane of the j-th valid claim
      const uint32_t t = __shfl_sync(0xffffffffu, v, (int)(src & 31u));
      if (lane == j && src != 0xffffffffu) mine = t;
    }
    return nvalid;
  }
}

// Persistent kernel, no grid barrier.  See the head of this file.
__global__ void __launch_bounds__(FLOOD_THREADS, WS_FLOOD_MINCTAS) flood_kernel(const __grid_constant__ FloodArgs a) {
  __shared__ FloodSmem sm;
  const ImageDims& d = a.d;
  const bool producer = threadIdx.x >= FLOOD_CONSUMERS;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FLOOD_STAGES; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (producer) {
    // =========================== producer warp ============================================
    bool pending[FLOOD_STAGES];
    uint32_t uses[FLOOD_STAGES];  // tiles staged into each stage so far
    for (int i = 0; i < FLOOD_STAGES; ++i) { pending[i] = false; uses[i] = 0u; }
    uint32_t q_tile = TILE_NONE;  // lane j: j-th tile of the last claim
    int q_n = 0, q_i = 0;         // tiles claimed / already handed to a stage
    // Publish the neighbours of a finished tile (k: lanes 0..3 hold the offers per direction).
    auto retire = [&](uint32_t tile, uint32_t k) {
      const uint32_t km = __ballot_sync(0xffffffffu, k != KEY_NONE);
      if (km) {
        const int trem = tile % d.tiles_per_img();
        const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
        // Count the entries we may add BEFORE any of them can be taken (the count must never read 0
        // while work exists); the fence orders that, and the consumers' results (ordered before us by
        // the mbarrier), before the marks and appends below.  The surplus is returned afterwards.
        if (lane == 0) atomicAdd(&a.b.ctrl[FC_OUTSTANDING], 3u);  // + 4 possible entries - this tile
        // Sequentially consistent fence (MEMBAR.SC.GPU), not just acquire-release: when the mark below finds
        // the neighbour already dirty and queued it only READS the mask word, and "my results are visible
        // to whoever takes that entry afterwards" is then a store-buffering shape that needs SC on this side.
        __threadfence();
        const uint32_t bk = flood_bucket(k >> 24, a.bucket_shift);
        uint32_t nb = TILE_NONE;
        if (k != KEY_NONE) {
          if (lane == DIR_UP && ty > 0) nb = tile - d.tiles_x;
          if (lane == DIR_DOWN && ty + 1 < d.tiles_y) nb = tile + d.tiles_x;
          if (lane == DIR_LEFT && tx > 0) nb = tile - 1;
          if (lane == DIR_RIGHT && tx + 1 < d.tiles_x) nb = tile + 1;
        }
        const bool app = nb != TILE_NONE && push_mark<false>(a.b, nb, bk);
        if (app) push_append(a.b, nb, bk);
        const int n = __popc(__ballot_sync(0xffffffffu, app));
        if (lane == 0 && n < 4) atomicSub(&a.b.ctrl[FC_OUTSTANDING], (uint32_t)(4 - n));
      } else {
        if (lane == 0) atomicSub(&a.b.ctrl[FC_OUTSTANDING], 1u);
      }
    };
    auto retire_stage = [&](int s) {
      retire(sm.tile[s], lane < 4 ? sm.key[s][lane] : KEY_NONE);
      pending[s] = false;
    };
    for (uint32_t slot = 0;; ++slot) {
      const int s = (int)(slot % FLOOD_STAGES);
      // The stage is reused: wait for its consumers.  Its tile is published AFTER the next tile's loads
      // have been issued (when one is already claimed), so the stage refills while we talk to the worklist.
      bool saved = false;
      uint32_t saved_tile = TILE_NONE, saved_k = KEY_NONE;
      if (pending[s]) {
        mbar_wait(&sm.empty[s], (uses[s] - 1u) & 1u);
        saved_tile = sm.tile[s];
        saved_k = lane < 4 ? sm.key[s][lane] : KEY_NONE;
        saved = true;
        pending[s] = false;
      }
      // claim a tile; while there is none, finish the other stage (its pushes may be the next work)
      uint32_t tile = TILE_NONE;
      uint32_t idle = 0, stalled = 0, seen_activations = 0xFFFFFFFFu;
      long long idle_since = 0;  // clock of the first claim that found the worklist empty
      for (;;) {
        if (q_i < q_n) {
          tile = __shfl_sync(0xffffffffu, q_tile, q_i);
          ++q_i;
          break;
        }
        if (saved) {  // before asking the worklist: what we publish may be exactly what we get
          retire(saved_tile, saved_k);
          saved = false;
        }
        q_n = flood_pop(a.b, lane, q_tile);
        q_i = 0;
        if (q_n) continue;
        if (idle_since == 0) idle_since = clock64();
        bool busy = false, retired = false;
        for (int k = 1; k < FLOOD_STAGES && !retired; ++k) {  // the other stages, oldest first
          const int o = (s + k) % FLOOD_STAGES;
          if (!pending[o]) continue;
          if (mbar_test(&sm.empty[o], (uses[o] - 1u) & 1u)) {
            retire_stage(o);
            retired = true;
          } else {
            busy = true;
          }
        }
        if (retired) continue;
        if (busy) {
          // our own consumers are still iterating
        } else if (ld_poll(&a.b.ctrl[FC_OUTSTANDING]) == 0u || (ld_poll(&a.b.ctrl[FC_ERROR]) & 56u)) {
          break;  // nothing queued, nothing in flight anywhere: the fixed point is reached
        }
        __nanosleep(100);
        // watchdog: never hang the device -- but only when NOTHING moves any more (a flood that is one long
        // dependency chain keeps most CTAs idle for as long as it takes)
        if ((++idle & 0xFFFFu) == 0u) {
          const uint32_t act = ld_poll(&a.b.ctrl[FC_ACTIVATIONS]);
          if (act != seen_activations) {
            seen_activations = act;
            stalled = 0;
          } else if (++stalled > 256u) {  // ~256 x 65536 polls (tens of seconds) without a single tile claimed anywhere
            atomicOr(&a.b.ctrl[FC_ERROR], 16u);
          }
        }
      }
      if (lane == 0) {
        if (idle_since) atomicAdd(&a.b.ctrl[FC_IDLE], (uint32_t)((clock64() - idle_since) >> 10));
        sm.tile[s] = tile;
        if (tile != TILE_NONE) {
          const int tpi = d.tiles_per_img();
          const int img = tile / tpi;
          const int trem = tile - img * tpi;
          const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
          sm.tbase[s] = (size_t)img * d.t_plane() + d.t_index(ty * TILE_H, tx * TILE_W);
        }
        sm.key[s][0] = sm.key[s][1] = sm.key[s][2] = sm.key[s][3] = KEY_NONE;
      }
      if (tile == TILE_NONE) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.full[s]);
        break;
      }
      if (lane == 0) atomicAdd(&a.b.ctrl[FC_ACTIVATIONS], 1u);
      {
        // bisect switch: bit 0 = arrival times by bulk copy, bit 1 = image bytes by bulk copy
        constexpr bool BULK_T = (WS_FLOOD_BULK & 1) != 0, BULK_P = (WS_FLOOD_BULK & 2) != 0;
        constexpr uint32_t TX = (BULK_T ? STG_H * STG_W * 4 : 0) + (BULK_P ? TILE_H * TILE_W : 0);
        const int tpi = d.tiles_per_img();
        const int img = tile / tpi;
        const int trem = tile - img * tpi;
        const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
        // box rows r0-1 .. r0+32 (padded row index r0 .. r0+33), columns c0-4 .. c0+67 (padded c0 .. c0+71)
        const uint32_t* tsrc = a.b.T + (size_t)img * d.t_plane() + (size_t)(ty * TILE_H) * d.t_pitch() + tx * TILE_W;
        const uint8_t* psrc = a.b.pix + (size_t)img * d.pix_plane() + (size_t)(ty * TILE_H) * d.pix_pitch() + tx * TILE_W;
        // (the claim's atomic has returned, and whoever woke the tile fenced its results before marking
        // it: the copies below, issued after the claim, read those results from L2)
        if (!BULK_T) {
          for (int i = lane; i < STG_H * STG_W; i += 32) {
            const int r = i / STG_W, c = i - r * STG_W;
            sm.st[s].T[i] = ld_cg(tsrc + (size_t)r * d.t_pitch() + c);
          }
        }
        if (!BULK_P) {
          for (int i = lane; i < TILE_H * TILE_W; i += 32) {
            const int r = i / TILE_W, c = i - r * TILE_W;
            sm.st[s].pix[i] = psrc[(size_t)r * d.pix_pitch() + c];
          }
        }
        __syncwarp();
        if (lane == 0) {
          if (TX) mbar_arrive_expect_tx(&sm.full[s], TX); else mbar_arrive(&sm.full[s]);
        }
        __syncwarp();
        // The tile's pixels were last written through the generic proxy (atomics / st.global by consumer
        // threads, possibly of other CTAs, ordered before us by the worklist atomics); the bulk copies
        // read them through the async proxy.  Every issuing lane needs the cross-proxy fence.
        asm volatile("fence.proxy.async.global;" ::: "memory");
        // cp.async.bulk is issued from the warp's UNIFORM datapath (SASS: ELECT + R2UR + UBLKCP in a
        // waterfall loop over the lanes).  Letting lanes issue different rows made two such loops run
        // on divergent halves of the warp at once, and they clobbered each other's uniform registers
        // (observed: rows landing late / in the wrong place).  One lane issues every row.
        if (lane == 0) {
          // box rows r0-1 .. r0+32 = padded rows r0 .. r0+33, columns c0-4 .. c0+67 = padded c0 .. c0+71
          if (BULK_T) tensor_g2s(sm.st[s].T, &a.tmT, tx * TILE_W, img * d.t_rows() + ty * TILE_H, &sm.full[s]);
          if (BULK_P) tensor_g2s(sm.st[s].pix, &a.tmP, tx * TILE_W, img * d.pix_rows() + ty * TILE_H, &sm.full[s]);
        }
        __syncwarp();
      }
      pending[s] = true;
      ++uses[s];
      if (saved) retire(saved_tile, saved_k);
      if (q_i == q_n && q_n > 1) {
        // The worklist is long (the last claim took several tiles): claim ahead while the consumers are
        // busy.  When it is short, a tile claimed early only misses what its neighbours are about to write.
        q_n = flood_pop(a.b, lane, q_tile);
        q_i = 0;
      }
    }
  } else {
    // =========================== consumer warps ===========================================
    long long waited = 0, busy = 0;  // thread 0: cycles spent waiting for a staged tile / iterating
    for (uint32_t slot = 0;; ++slot) {
      const int s = (int)(slot % FLOOD_STAGES);
      const long long t0 = clock64();
      mbar_wait(&sm.full[s], (slot / FLOOD_STAGES) & 1u);
      const long long t1 = clock64();
      const uint32_t tile = sm.tile[s];
      if (tile == TILE_NONE) break;
      flood_consume(a, sm, s);
      consumer_sync();  // all results of the tile issued, all reads of the stage done
      if (threadIdx.x == 0) mbar_arrive(&sm.empty[s]);
      waited += t1 - t0;
      busy += clock64() - t1;
    }
    if (threadIdx.x == 0) {  // in units of 1024 cycles, summed over the CTAs
      atomicAdd(&a.b.ctrl[FC_WAIT_KCYC], (uint32_t)(waited >> 10));
      atomicAdd(&a.b.ctrl[FC_BUSY_KCYC], (uint32_t)(busy >> 10));
    }
  }
}

static int coop_max_grid_flood(int device) {
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)flood_kernel, FLOOD_THREADS, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

int flood_max_grid(int device) { return coop_max_grid_flood(device); }

// Granularity of the worklist's priority.  Fields whose structures are much larger than a tile (few
// seeds per tile) have long-range dependencies: fine buckets (4 levels) keep the tiles in the order of
// the reference's level loop and save most re-activations.  On noise-like fields (many seeds per tile)
// everything is local, a tile's wake-ups come from all sides at unrelated levels, and it pays to let
// them accumulate: coarse buckets (32 levels).  Either choice gives the same result, only the number of
// tile activations differs.
int flood_bucket_shift(size_t nseeds, const ImageDims& d) {
  const size_t tiles = (size_t)d.tiles_total();
  return nseeds >= 4 * tiles ? 5 : 2;
}

// The grid never exceeds the co-resident CTA count; no CTA waits for a particular other CTA (only for
// the worklist to drain), so a plain launch is enough.
// Tensor maps of a plan's arrival-time and image planes (host; the driver's encoder through the runtime).
cudaError_t flood_make_tensor_maps(const FloodBuffers& b, const ImageDims& d, void* out_maps) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess) return e;
  if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
  CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(out_maps);
  const cuuint32_t ones[2] = {1, 1};
  {
    const cuuint64_t dim[2] = {(cuuint64_t)d.t_pitch(), (cuuint64_t)d.t_rows() * (cuuint64_t)d.n_img};
    const cuuint64_t stride[1] = {(cuuint64_t)d.t_pitch() * 4};
    const cuuint32_t box[2] = {STG_W, STG_H};
    if (((encode_fn)fn)(&maps[0], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, b.T, dim, stride, box, ones,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t dim[2] = {(cuuint64_t)d.pix_pitch(), (cuuint64_t)d.pix_rows() * (cuuint64_t)d.n_img};
    const cuuint64_t stride[1] = {(cuuint64_t)d.pix_pitch()};
    const cuuint32_t box[2] = {TILE_W, TILE_H};
    if (((encode_fn)fn)(&maps[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, b.pix, dim, stride, box, ones,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

// The grid never exceeds the co-resident CTA count; no CTA waits for a particular other CTA (only for
// the worklist to drain), so a plain launch is enough.
cudaError_t launch_flood(FloodBuffers b, ImageDims d, int check_overflow, int bucket_shift, int grid,
                         const void* tensor_maps, cudaStream_t s) {
  FloodArgs a;
  memcpy(&a.tmT, tensor_maps, sizeof(CUtensorMap));
  memcpy(&a.tmP, (const char*)tensor_maps + sizeof(CUtensorMap), sizeof(CUtensorMap));
  a.b = b;
  a.d = d;
  a.check_overflow = check_overflow;
  a.bucket_shift = bucket_shift;
  const int want = d.tiles_total();
  const int g = want < grid ? (want > 0 ? want : 1) : grid;
  flood_kernel<<<g, FLOOD_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// row-strip decomposition: boundary rows of arrival times and of labels
// ---------------------------------------------------------------------------

// rows `ra` / `rb` of the padded arrival times -> dense rows (negative row = skip)
__global__ void __launch_bounds__(256) strip_export_T_kernel(const uint32_t* __restrict__ T, ImageDims d, int ra,
                                                             int rb, uint32_t* __restrict__ out_a,
                                                             uint32_t* __restrict__ out_b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  if (ra >= 0) out_a[c] = ld_cg(T + d.t_index(ra, c));
  if (rb >= 0) out_b[c] = ld_cg(T + d.t_index(rb, c));
}

// min-merge a neighbour's row into halo row `row`; where it got lower, wake the tile that holds the
// adjacent owned row `nb_row`
__global__ void __launch_bounds__(256) strip_import_T_kernel(FloodBuffers b, ImageDims d, int row, int nb_row,
                                                             const uint32_t* __restrict__ in, int bucket_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  const uint32_t v = in[c];
  uint32_t* t = b.T + d.t_index(row, c);
  if (v < ld_cg(t)) {
    st_cg(t, v);
    atomicOr(&b.ctrl[FC_STRIP_CHANGED], 1u);
    __threadfence();
    push_tile<true>(b, (uint32_t)((nb_row / TILE_H) * d.tiles_x + c / TILE_W), flood_bucket((v + 1u) >> 24, bucket_shift));
  }
}

__device__ __forceinline__ uint32_t lab_through_rim(const uint32_t* lab, const uint32_t* rim, size_t p) {
  const uint32_t v = ld_cg(lab + p);
  if (v & LAB_RESOLVED) return v;
  const uint32_t t = ld_cg(rim + v);   // the label plane may be unfinished: what the reference points at decides
  return (t & LAB_RESOLVED) ? t : v;
}
__global__ void __launch_bounds__(256) strip_export_lab_kernel(const uint32_t* __restrict__ lab,
                                                               const uint32_t* __restrict__ rim, ImageDims d, int ra,
                                                               int rb, uint32_t* __restrict__ out_a,
                                                               uint32_t* __restrict__ out_b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  if (ra >= 0) out_a[c] = lab_through_rim(lab, rim, (size_t)ra * d.cols + c);
  if (rb >= 0) out_b[c] = lab_through_rim(lab, rim, (size_t)rb * d.cols + c);
}

__global__ void __launch_bounds__(256) strip_count_pending_rim_kernel(const uint32_t* __restrict__ rim, size_t n,
                                                                      uint32_t* __restrict__ ctrl) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int cnt = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    cnt += !(ld_cg(rim + i) & LAB_RESOLVED);
  cnt = __syncthreads_count(cnt);  // (the number of THREADS with pending entries is enough for a zero test)
  if (threadIdx.x == 0 && cnt) atomicAdd(&ctrl[FC_STRIP_PENDING], (uint32_t)cnt);
}
cudaError_t launch_strip_count_pending_rim(const uint32_t* rim, size_t n, uint32_t* ctrl, int sms, cudaStream_t s) {
  const size_t want = (n + 255) / 256;
  const size_t cap = (size_t)sms * 4;
  strip_count_pending_rim_kernel<<<(unsigned)(want < cap ? (want ? want : 1) : cap), 256, 0, s>>>(rim, n, ctrl);
  return cudaGetLastError();
}

// resolved labels of the neighbour's boundary row replace the pending words of halo row `row`, in the
// label plane and in the row's pending slots behind the rim entries (labels.cu)
__global__ void __launch_bounds__(256) strip_import_lab_kernel(FloodBuffers b, ImageDims d, int row,
                                                               const uint32_t* __restrict__ in, size_t slot0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  const uint32_t v = in[c];
  uint32_t* l = b.lab + (size_t)row * d.cols + c;
  if ((v & LAB_RESOLVED) && !(ld_cg(l) & LAB_RESOLVED)) {
    st_cg(l, v);
    st_cg(b.rim + slot0 + c, v);
  }
}

// pixels of the owned rows [r0, r1] whose label is still a pointer
__global__ void __launch_bounds__(256) strip_count_pending_kernel(const uint32_t* __restrict__ lab, ImageDims d,
                                                                  int r0, int r1, uint32_t* __restrict__ ctrl) {
  const size_t lo = (size_t)r0 * d.cols, hi = (size_t)(r1 + 1) * d.cols;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  int n = 0;
  for (size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride)
    n += !(ld_cg(lab + i) & LAB_RESOLVED);
  n = __syncthreads_count(n);  // number of THREADS with pending pixels is enough for a zero test...
  if (threadIdx.x == 0 && n) atomicAdd(&ctrl[FC_STRIP_PENDING], (uint32_t)n);
}

cudaError_t launch_strip_export_T(const uint32_t* T, ImageDims d, int ra, int rb, uint32_t* out_a, uint32_t* out_b,
                                  cudaStream_t s) {
  strip_export_T_kernel<<<(d.cols + 255) / 256, 256, 0, s>>>(T, d, ra, rb, out_a, out_b);
  return cudaGetLastError();
}
cudaError_t launch_strip_import_T(FloodBuffers b, ImageDims d, int row, int nb_row, const uint32_t* in,
                                  int bucket_shift, cudaStream_t s) {
  strip_import_T_kernel<<<(d.cols + 255) / 256, 256, 0, s>>>(b, d, row, nb_row, in, bucket_shift);
  return cudaGetLastError();
}
cudaError_t launch_strip_export_lab(const uint32_t* lab, const uint32_t* rim, ImageDims d, int ra, int rb,
                                    uint32_t* out_a, uint32_t* out_b, cudaStream_t s) {
  strip_export_lab_kernel<<<(d.cols + 255) / 256, 256, 0, s>>>(lab, rim, d, ra, rb, out_a, out_b);
  return cudaGetLastError();
}
cudaError_t launch_strip_import_lab(FloodBuffers b, ImageDims d, int row, const uint32_t* in, cudaStream_t s) {
  const size_t slot0 = rim_words(d) - 2 * (size_t)d.cols + (row == 0 ? 0 : (size_t)d.cols);
  strip_import_lab_kernel<<<(d.cols + 255) / 256, 256, 0, s>>>(b, d, row, in, slot0);
  return cudaGetLastError();
}
cudaError_t launch_strip_count_pending(const uint32_t* lab, ImageDims d, int r0, int r1, uint32_t* ctrl,
                                       cudaStream_t s) {
  strip_count_pending_kernel<<<num_sms() * 4, 256, 0, s>>>(lab, d, r0, r1, ctrl);
  return cudaGetLastError();
}

// dense copy of the padded arrival times (diagnostic accessor)
__global__ void __launch_bounds__(256) unpad_T_kernel(const uint32_t* __restrict__ Tp, ImageDims d,
                                                      uint32_t* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  const int r = blockIdx.y, img = blockIdx.z;
  out[(size_t)img * d.px_per_img() + (size_t)r * d.cols + c] = Tp[(size_t)img * d.t_plane() + d.t_index(r, c)];
}

cudaError_t launch_unpad_T(const uint32_t* Tp, ImageDims d, uint32_t* out, cudaStream_t s) {
  dim3 grid((d.cols + 255) / 256, d.rows, d.n_img);
  unpad_T_kernel<<<grid, 256, 0, s>>>(Tp, d, out);
  return cudaGetLastError();
}

}  // namespace ws
