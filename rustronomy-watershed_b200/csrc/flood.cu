// flood.cu -- K2: arrival times of all water levels in one persistent kernel, and the
// parent-pointer kernel that turns them into the reference's colour decision.
//
// Replaces the reference's level loop x 'colouring_loop x find_flooded_px x write-back
// (lib.rs:1379-1438 / 1689-1748 and 196-257).  See common.cuh for the arrival-time
// encoding and DESIGN.md section 2 for why any relaxation order gives the reference's result.
//
// Kernel structure (sm_100a):
//   * persistent cooperative grid, one worklist of active tiles per sweep, grid barrier
//     between sweeps, three rotating lists so pushes never race with the reset;
//   * every CTA = 8 consumer warps + 1 producer warp.  The producer hands out tiles
//     (atomic cursor), stages each tile's 34 x 72-word box of arrival times and its
//     32 x 64 image bytes with the bulk-copy engine (cp.async.bulk -> SASS UBLKCP, completion
//     on an mbarrier) into a two-stage ring, and afterwards publishes the tile's neighbours
//     to the next worklist.  The consumers therefore never wait on a global round trip:
//     they copy the stage into an odd-stride working tile and iterate on it;
//   * the in-tile iteration alternates column and row ownership (8 pixels of Gauss-Seidel
//     along the phase's axis per step) until a phase changes nothing.
#include "kernels.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

#ifndef WS_FLOOD_BULK
#define WS_FLOOD_BULK 3
#endif

namespace ws {

// ---------------------------------------------------------------------------
// small PTX helpers (mbarrier + bulk async copy + named barriers)
// ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy (16-byte aligned, multiple of 16 bytes), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
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

// T = "never", labels = uncoloured, and the image re-encoded for the flood: a pixel that can
// never flood -- not a window centre (lib.rs:220), above the last water level (filter (1),
// lib.rs:224, over levels 0..=max), or padding -- is stored as 255.
__global__ void __launch_bounds__(256) fill_state_kernel(FloodBuffers b, ImageDims d, const uint8_t* __restrict__ img,
                                                         uint32_t lmax) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // arrival times: the padded plane is a multiple of 8 words
  const size_t nt4 = d.t_plane() * d.n_img / 4;
  const uint4 inf4 = make_uint4(T_INF, T_INF, T_INF, T_INF);
  uint4* T4 = reinterpret_cast<uint4*>(b.T);
  for (size_t i = tid; i < nt4; i += stride) __stcg(T4 + i, inf4);
  // labels
  const size_t nl = d.px_total(), nl4 = nl / 4;
  uint4* L4 = reinterpret_cast<uint4*>(b.lab);
  for (size_t i = tid; i < nl4; i += stride) __stcg(L4 + i, make_uint4(0u, 0u, 0u, 0u));
  for (size_t i = nl4 * 4 + tid; i < nl; i += stride) b.lab[i] = 0u;
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

cudaError_t launch_fill_state(FloodBuffers b, ImageDims d, const uint8_t* img, uint32_t lmax, cudaStream_t s) {
  fill_state_kernel<<<148 * 16, 256, 0, s>>>(b, d, img, lmax);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(b.flags, 0, sizeof(uint32_t) * (size_t)d.tiles_total(), s);
  if (e != cudaSuccess) return e;
  return cudaMemsetAsync(b.ctrl, 0, sizeof(uint32_t) * FC_WORDS, s);
}

__device__ __forceinline__ void push_tile(const FloodBuffers& b, uint32_t ntiles, int list, uint32_t tile) {
  const uint32_t bit = 1u << list;
  if (ld_cg(&b.flags[tile]) & bit) return;  // already queued (the common case when seeds are dense)
  const uint32_t old = atomicOr(&b.flags[tile], bit);
  if (!(old & bit)) {
    const uint32_t pos = atomicAdd(&b.ctrl[FC_COUNT0 + list], 1u);
    st_cg(&b.lists[(size_t)list * ntiles + pos], tile);
  }
}

// Colour the starting pixels (lib.rs:1365-1367): T = 0, colour = index + 1, a later
// duplicate overwrites an earlier one (sequential loop) == the largest index wins.
__global__ void __launch_bounds__(256) seed_init_kernel(FloodBuffers b, ImageDims d,
                                                        const uint32_t* __restrict__ seeds_rc,
                                                        const uint32_t* __restrict__ seed_off, uint32_t nseeds,
                                                        uint32_t colour_base) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseeds) return;
  int lo = 0, hi = d.n_img;  // slice of seed i: last b with seed_off[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(seed_off + mid) <= i) lo = mid; else hi = mid;
  }
  const int img = lo;
  const uint32_t r = seeds_rc[2 * (size_t)i], c = seeds_rc[2 * (size_t)i + 1];
  if (r >= (uint32_t)d.rows || c >= (uint32_t)d.cols) {
    atomicOr(&b.ctrl[FC_ERROR], 1u);
    return;
  }
  st_cg(&b.T[(size_t)img * d.t_plane() + d.t_index((int)r, (int)c)], 0u);
  atomicMax(&b.lab[(size_t)img * d.px_per_img() + (size_t)r * d.cols + c],
            LAB_RESOLVED | (colour_base + i - __ldg(seed_off + img) + 1u));
  // Red-black order over the tiles: the first sweep takes the even tiles (tx + ty even), the second the
  // odd ones -- which then already see their neighbours' results, a Gauss-Seidel step at tile level
  // that saves re-activations.  A tile's 4-neighbours have the other parity, so later sweeps alternate
  // by themselves; only the seeding has to split the lists.
  const uint32_t ntiles = (uint32_t)d.tiles_total();
  const int ty = r / TILE_H, tx = c / TILE_W;
  const uint32_t tile = (uint32_t)img * d.tiles_per_img() + ty * d.tiles_x + tx;
  const int par = (tx + ty) & 1;
  push_tile(b, ntiles, par, tile);
  if (r % TILE_H == 0 && ty > 0) push_tile(b, ntiles, par ^ 1, tile - d.tiles_x);
  if (r % TILE_H == TILE_H - 1 && ty + 1 < d.tiles_y) push_tile(b, ntiles, par ^ 1, tile + d.tiles_x);
  if (c % TILE_W == 0 && tx > 0) push_tile(b, ntiles, par ^ 1, tile - 1);
  if (c % TILE_W == TILE_W - 1 && tx + 1 < d.tiles_x) push_tile(b, ntiles, par ^ 1, tile + 1);
}

cudaError_t launch_seed_init(FloodBuffers b, ImageDims d, const uint32_t* seeds_rc, const uint32_t* seed_off,
                             uint32_t nseeds, uint32_t colour_base, cudaStream_t s) {
  if (nseeds == 0) return cudaSuccess;
  seed_init_kernel<<<(nseeds + 255) / 256, 256, 0, s>>>(b, d, seeds_rc, seed_off, nseeds, colour_base);
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
  const unsigned grid = (unsigned)(want < (size_t)148 * 8 ? want : (size_t)148 * 8);
  seeds_convert_kernel<<<grid, 256, 0, s>>>(in, out, n2, rows, cols);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// the flood kernel
// ---------------------------------------------------------------------------

struct FloodArgs {
  FloodBuffers b;
  ImageDims d;
  int check_overflow;
};

enum { EDGE_UP = 1, EDGE_DOWN = 2, EDGE_LEFT = 4, EDGE_RIGHT = 8 };
constexpr uint32_t TILE_NONE = 0xFFFFFFFFu;

struct FloodStage {
  uint32_t T[STG_H * STG_W];      // staged arrival times (also the "before" image of the tile)
  uint8_t pix[TILE_H * TILE_W];   // staged image bytes
};

struct __align__(128) FloodSmem {
  FloodStage st[2];
  uint32_t W[SM_H * SM_W];          // working tile incl. halo
  uint8_t wpix[TILE_H * PIX_W];     // working image tile
  uint64_t full[2], empty[2];       // mbarriers of the ring
  uint32_t tile[2], edge[2];
};

__device__ __forceinline__ uint32_t flood_A(uint32_t pix) { return pix == 255u ? T_INF : ((pix << 24) | 1u); }

// ---- consumer side: one tile to its local fixed point ------------------------------------
__device__ __forceinline__ void flood_consume(const FloodArgs& a, FloodSmem& sm, int s, uint32_t tile) {
  const ImageDims& d = a.d;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const FloodStage& st = sm.st[s];

  // stage -> working tile (odd stride).  Staged row = image columns c0-4 .. c0+67; the working
  // tile keeps c0-1 .. c0+64.
  for (int i = tid; i < SM_H * (TILE_W + 2); i += FLOOD_CONSUMERS) {
    const int lr = i / (TILE_W + 2), lc = i - lr * (TILE_W + 2);
    sm.W[lr * SM_W + lc] = st.T[lr * STG_W + lc + (T_PAD_L - 1)];
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(st.pix);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sm.wpix);
    for (int i = tid; i < TILE_H * TILE_W / 4; i += FLOOD_CONSUMERS) {
      const int r = i / (TILE_W / 4), w = i - r * (TILE_W / 4);
      dst[r * (PIX_W / 4) + w] = src[i];
    }
  }
  consumer_sync();

  // column ownership: column cl, rows cgp*8 .. cgp*8+7;  row ownership: row rl, columns rgp*8 .. +7
  const int cl = tid % TILE_W, cgp = tid / TILE_W;
  uint32_t* colp = sm.W + (cgp * ROWS_PER_THREAD + 1) * SM_W + cl + 1;
  const int rl = lane, rgp = warp;
  uint32_t* rowp = sm.W + (rl + 1) * SM_W + rgp * ROWS_PER_THREAD + 1;
  uint32_t Ac[ROWS_PER_THREAD], Ar[ROWS_PER_THREAD];
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    Ac[i] = flood_A(sm.wpix[(cgp * ROWS_PER_THREAD + i) * PIX_W + cl]);
    Ar[i] = flood_A(sm.wpix[rl * PIX_W + rgp * ROWS_PER_THREAD + i]);
  }

  bool ovf = false;
  uint32_t nphase = 0;
  for (;;) {
    nphase += 2;
    {  // ---- column phase: T(p) = max(A(p), 1 + min over the 4 neighbours) down then up ----
      uint32_t t[ROWS_PER_THREAD], m[ROWS_PER_THREAD];
      const uint32_t up = colp[-SM_W], dn = colp[ROWS_PER_THREAD * SM_W];
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        t[i] = colp[i * SM_W];
        m[i] = min(colp[i * SM_W - 1], colp[i * SM_W + 1]);
      }
      uint32_t it = 0, prev = up;
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        const uint32_t nb = (i < ROWS_PER_THREAD - 1) ? t[i + 1] : dn;
        const uint32_t n = __viaddmax_u32(__vimin3_u32(m[i], prev, nb), 1u, Ac[i]);  // max(A, 1 + min3)
        if (n < t[i]) { t[i] = n; it |= 1u << i; ovf |= ((n & HOP_MASK) == 0u); }
        prev = t[i];
      }
      prev = dn;
#pragma unroll
      for (int i = ROWS_PER_THREAD - 1; i >= 0; --i) {
        const uint32_t nb = (i > 0) ? t[i - 1] : up;
        const uint32_t n = __viaddmax_u32(__vimin3_u32(m[i], prev, nb), 1u, Ac[i]);  // max(A, 1 + min3)
        if (n < t[i]) { t[i] = n; it |= 1u << i; ovf |= ((n & HOP_MASK) == 0u); }
        prev = t[i];
      }
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i)
        if (it & (1u << i)) colp[i * SM_W] = t[i];
      if (!consumer_sync_or(it != 0u)) break;
    }
    {  // ---- row phase: right then left ----------------------------------------------------
      uint32_t t[ROWS_PER_THREAD], m[ROWS_PER_THREAD];
      const uint32_t lf = rowp[-1], rt = rowp[ROWS_PER_THREAD];
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        t[i] = rowp[i];
        m[i] = min(rowp[i - SM_W], rowp[i + SM_W]);
      }
      uint32_t it = 0, prev = lf;
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        const uint32_t nb = (i < ROWS_PER_THREAD - 1) ? t[i + 1] : rt;
        const uint32_t n = __viaddmax_u32(__vimin3_u32(m[i], prev, nb), 1u, Ar[i]);
        if (n < t[i]) { t[i] = n; it |= 1u << i; ovf |= ((n & HOP_MASK) == 0u); }
        prev = t[i];
      }
      prev = rt;
#pragma unroll
      for (int i = ROWS_PER_THREAD - 1; i >= 0; --i) {
        const uint32_t nb = (i > 0) ? t[i - 1] : lf;
        const uint32_t n = __viaddmax_u32(__vimin3_u32(m[i], prev, nb), 1u, Ar[i]);
        if (n < t[i]) { t[i] = n; it |= 1u << i; ovf |= ((n & HOP_MASK) == 0u); }
        prev = t[i];
      }
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i)
        if (it & (1u << i)) rowp[i] = t[i];
      if (!consumer_sync_or(it != 0u)) break;
    }
  }

  // write back what changed against the staged copy (column ownership: coalesced along rows)
  const int tpi = d.tiles_per_img();
  const int img = tile / tpi;
  const int trem = tile - img * tpi;
  const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
  uint32_t* Tg = a.b.T + (size_t)img * d.t_plane() +
                 d.t_index(ty * TILE_H + cgp * ROWS_PER_THREAD, tx * TILE_W + cl);
  const int tp = d.t_pitch();
  // A neighbour tile re-runs only if a changed edge pixel can still lower the pixel facing it:
  // T(edge) + 1 < T(facing pixel), the latter as staged in our halo (never newer than the truth, so the
  // test never drops a needed wake-up).  Without it every tile woke all four neighbours, including the
  // one its values came from, and most activations were such echoes.
  uint32_t e = 0;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const uint32_t v = colp[i * SM_W];
    if (v != st.T[(cgp * ROWS_PER_THREAD + i + 1) * STG_W + cl + T_PAD_L]) {
      st_cg(Tg + (size_t)i * tp, v);
      if (cl == 0 && v + 1u < colp[i * SM_W - 1]) e |= EDGE_LEFT;
      if (cl == TILE_W - 1 && v + 1u < colp[i * SM_W + 1]) e |= EDGE_RIGHT;
      if (cgp == 0 && i == 0 && v + 1u < colp[-SM_W]) e |= EDGE_UP;
      if (cgp == TILE_H / ROWS_PER_THREAD - 1 && i == ROWS_PER_THREAD - 1 && v + 1u < colp[ROWS_PER_THREAD * SM_W])
        e |= EDGE_DOWN;
    }
  }
  if (e) atomicOr(&sm.edge[s], e);
  if (tid == 0) atomicAdd(&a.b.ctrl[FC_PHASES], nphase);
  if (a.check_overflow && ovf) atomicOr(&a.b.ctrl[FC_ERROR], 2u);
}

// Persistent cooperative kernel.  Sweep k drains worklist k%3, fills worklist (k+1)%3 and resets
// worklist (k+2)%3; one grid barrier per sweep; ends when a sweep starts with an empty list.
__global__ void __launch_bounds__(FLOOD_THREADS, 3) flood_kernel(FloodArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ FloodSmem sm;
  const ImageDims& d = a.d;
  const uint32_t ntiles = (uint32_t)d.tiles_total();
  const bool producer = threadIdx.x >= FLOOD_CONSUMERS;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    mbar_init(&sm.empty[0], 1);
    mbar_init(&sm.empty[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t slot = 0;  // ring slots used so far (same sequence on both sides); stage = slot & 1
  int cur = 0;
  for (uint32_t sweep = 0;; ++sweep) {
    const uint32_t n = ld_cg(&a.b.ctrl[FC_COUNT0 + cur]);
    // an empty list ends the flood -- except the very first one (even tiles), after which the odd
    // tiles seeded into list 1 still have to run
    if (n == 0 && (sweep > 0 || ld_cg(&a.b.ctrl[FC_COUNT0 + 1]) == 0)) break;
    const int nxt = (cur + 1) % 3, old = (cur + 2) % 3;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st_cg(&a.b.ctrl[FC_COUNT0 + old], 0u);
      st_cg(&a.b.ctrl[FC_CURSOR0 + old], 0u);
      atomicAdd(&a.b.ctrl[FC_SWEEPS], 1u);
    }

    if (producer) {
      // =========================== producer warp ============================================
      bool pending[2] = {false, false};
      // Publish the neighbours of the tile that last used stage s (after its consumers released it).
      auto retire = [&](int s, uint32_t use_idx) {
        mbar_wait(&sm.empty[s], use_idx & 1u);
        const uint32_t tile = sm.tile[s];
        const uint32_t e = sm.edge[s];
        if (e) {
          const int trem = tile % d.tiles_per_img();
          const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
          __threadfence();  // the consumers' stores (ordered before us by the mbarrier) before the pushes
          if (lane == 0 && (e & EDGE_UP) && ty > 0) push_tile(a.b, ntiles, nxt, tile - d.tiles_x);
          if (lane == 1 && (e & EDGE_DOWN) && ty + 1 < d.tiles_y) push_tile(a.b, ntiles, nxt, tile + d.tiles_x);
          if (lane == 2 && (e & EDGE_LEFT) && tx > 0) push_tile(a.b, ntiles, nxt, tile - 1);
          if (lane == 3 && (e & EDGE_RIGHT) && tx + 1 < d.tiles_x) push_tile(a.b, ntiles, nxt, tile + 1);
        }
        __syncwarp();
      };
      for (;;) {
        const int s = slot & 1;
        if (pending[s]) {
          retire(s, (slot >> 1) - 1u);
          pending[s] = false;
        } else if (slot >= 2) {
          mbar_wait(&sm.empty[s], ((slot >> 1) - 1u) & 1u);  // a "no more tiles" slot: just keep the phases aligned
        }
        uint32_t tile = TILE_NONE;
        if (lane == 0) {
          const uint32_t k = atomicAdd(&a.b.ctrl[FC_CURSOR0 + cur], 1u);
          if (k < n) {
            tile = ld_cg(&a.b.lists[(size_t)cur * ntiles + k]);
            atomicAnd(&a.b.flags[tile], ~(1u << cur));
            atomicAdd(&a.b.ctrl[FC_ACTIVATIONS], 1u);
          }
          sm.tile[s] = tile;
          sm.edge[s] = 0u;
        }
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile == TILE_NONE) {
          if (lane == 0) mbar_arrive(&sm.full[s]);
          ++slot;
          break;
        }
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
          // The tile's pixels were last written through the generic proxy (st.global by consumer
          // threads, possibly of other CTAs, ordered before us by the grid barrier); the bulk copies
          // read them through the async proxy.  Every issuing lane needs the cross-proxy fence.
          asm volatile("fence.proxy.async.global;" ::: "memory");
          // cp.async.bulk is issued from the warp's UNIFORM datapath (SASS: ELECT + R2UR + UBLKCP in a
          // waterfall loop over the lanes).  Letting lanes issue different rows made two such loops run
          // on divergent halves of the warp at once, and they clobbered each other's uniform registers
          // (observed: rows landing late / in the wrong place).  One lane issues every row.
          if (lane == 0) {
            if (BULK_T)
              for (int r = 0; r < STG_H; ++r)
                bulk_g2s(&sm.st[s].T[r * STG_W], tsrc + (size_t)r * d.t_pitch(), STG_W * 4, &sm.full[s]);
            if (BULK_P)
              for (int r = 0; r < TILE_H; ++r)
                bulk_g2s(&sm.st[s].pix[r * TILE_W], psrc + (size_t)r * d.pix_pitch(), TILE_W, &sm.full[s]);
          }
          __syncwarp();
        }
        pending[s] = true;
        ++slot;
      }
      // drain: the slot before the terminating one may still be in flight; then the terminator's ack
      {
        const uint32_t last = slot - 1;        // the "no more tiles" slot
        const int so = (last & 1) ^ 1;
        if (pending[so]) retire(so, (last - 1) >> 1);
        mbar_wait(&sm.empty[last & 1], (last >> 1) & 1u);
      }
    } else {
      // =========================== consumer warps ===========================================
      for (;;) {
        const int s = slot & 1;
        mbar_wait(&sm.full[s], (slot >> 1) & 1u);
        const uint32_t tile = sm.tile[s];
        ++slot;
        if (tile == TILE_NONE) {
          consumer_sync();
          if (threadIdx.x == 0) mbar_arrive(&sm.empty[s]);
          break;
        }
        flood_consume(a, sm, s, tile);
        consumer_sync();  // all stores of the tile issued, all reads of the stage done
        if (threadIdx.x == 0) mbar_arrive(&sm.empty[s]);
      }
    }
    __syncthreads();
    grid.sync();
    cur = nxt;
  }
}

static int coop_max_grid_flood(int device) {
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)flood_kernel, FLOOD_THREADS, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

int flood_max_grid(int device) { return coop_max_grid_flood(device); }

cudaError_t launch_flood(FloodBuffers b, ImageDims d, int check_overflow, int grid, cudaStream_t s) {
  FloodArgs a{b, d, check_overflow};
  void* args[] = {&a};
  const int want = d.tiles_total();
  const int g = want < grid ? (want > 0 ? want : 1) : grid;
  return cudaLaunchCooperativeKernel((const void*)flood_kernel, dim3(g), dim3(FLOOD_THREADS), args, 0, s);
}

// ---------------------------------------------------------------------------
// K3a  parent pointers + level bytes
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256) parent_kernel(FloodBuffers b, ImageDims d) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  const int img = blockIdx.z;
  if (c >= d.cols) return;
  const size_t p = (size_t)img * d.px_per_img() + (size_t)r * d.cols + c;
  const uint32_t* Tq = b.T + (size_t)img * d.t_plane() + d.t_index(r, c);
  const int tp = d.t_pitch();
  const uint32_t t = Tq[0];
  b.lvl[p] = (t >= T_INF) ? (uint8_t)255 : (uint8_t)(t >> 24);
  if (t >= T_INF) {
    b.lab[p] = LAB_RESOLVED;  // UNCOLOURED
    return;
  }
  if (d.is_halo_row(r)) {
    // a neighbouring strip owns this pixel: its label arrives by exchange; until then the word
    // points at itself ("pending"), which pointer jumping leaves alone
    b.lab[p] = (uint32_t)p;
    return;
  }
  if (t == 0u) return;  // seed: coloured by seed_init
  // A coloured non-seed pixel is interior, so all four neighbours exist.  The coloured
  // neighbours the reference sees when it colours p are exactly those with T(q) < T(p);
  // `col0` is the first of them in the order down, right, left, up (lib.rs:190, 245).
  size_t q;
  if (Tq[tp] < t) q = p + d.cols;
  else if (Tq[1] < t) q = p + 1;
  else if (Tq[-1] < t) q = p - 1;
  else if (Tq[-tp] < t) q = p - d.cols;
  else {
    atomicOr(&b.ctrl[FC_ERROR], 4u);  // cannot happen at a fixed point
    b.lab[p] = LAB_RESOLVED;
    return;
  }
  b.lab[p] = (uint32_t)q;
}

cudaError_t launch_parent(FloodBuffers b, ImageDims d, cudaStream_t s) {
  dim3 grid((d.cols + 255) / 256, d.rows, d.n_img);
  parent_kernel<<<grid, 256, 0, s>>>(b, d);
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
                                                             const uint32_t* __restrict__ in) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  const uint32_t v = in[c];
  uint32_t* t = b.T + d.t_index(row, c);
  if (v < ld_cg(t)) {
    st_cg(t, v);
    atomicOr(&b.ctrl[FC_STRIP_CHANGED], 1u);
    push_tile(b, (uint32_t)d.tiles_total(), 0, (uint32_t)((nb_row / TILE_H) * d.tiles_x + c / TILE_W));
  }
}

__global__ void __launch_bounds__(256) strip_export_lab_kernel(const uint32_t* __restrict__ lab, ImageDims d, int ra,
                                                               int rb, uint32_t* __restrict__ out_a,
                                                               uint32_t* __restrict__ out_b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  if (ra >= 0) out_a[c] = ld_cg(lab + (size_t)ra * d.cols + c);
  if (rb >= 0) out_b[c] = ld_cg(lab + (size_t)rb * d.cols + c);
}

// resolved labels of the neighbour's boundary row replace the pending words of halo row `row`
__global__ void __launch_bounds__(256) strip_import_lab_kernel(uint32_t* __restrict__ lab, ImageDims d, int row,
                                                               const uint32_t* __restrict__ in) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.cols) return;
  const uint32_t v = in[c];
  uint32_t* l = lab + (size_t)row * d.cols + c;
  if ((v & LAB_RESOLVED) && !(ld_cg(l) & LAB_RESOLVED)) st_cg(l, v);
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
                                  cudaStream_t s) {
  strip_import_T_kernel<<<(d.cols + 255) / 256, 256, 0, s>>>(b, d, row, nb_row, in);
  return cudaGetLastError();
}
cudaError_t launch_strip_export_lab(const uint32_t* lab, ImageDims d, int ra, int rb, uint32_t* out_a,
                                    uint32_t* out_b, cudaStream_t s) {
  strip_export_lab_kernel<<<(d.cols + 255) / 256, 256, 0, s>>>(lab, d, ra, rb, out_a, out_b);
  return cudaGetLastError();
}
cudaError_t launch_strip_import_lab(uint32_t* lab, ImageDims d, int row, const uint32_t* in, cudaStream_t s) {
  strip_import_lab_kernel<<<(d.cols + 255) / 256, 256, 0, s>>>(lab, d, row, in);
  return cudaGetLastError();
}
cudaError_t launch_strip_count_pending(const uint32_t* lab, ImageDims d, int r0, int r1, uint32_t* ctrl,
                                       cudaStream_t s) {
  strip_count_pending_kernel<<<148 * 4, 256, 0, s>>>(lab, d, r0, r1, ctrl);
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
