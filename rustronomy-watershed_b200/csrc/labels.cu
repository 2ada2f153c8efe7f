// labels.cu -- K3: from arrival times to the reference's colours.
//
// The reference colours a flooded pixel with the colour of the coloured neighbours it sees at that
// moment (lib.rs:235-255): exactly the neighbours q with T(q) < T(p); canonical tie-break = the first of
// them in the order down, right, left, up (lib.rs:190, 245, `col0`).  One parent per pixel gives a forest
// rooted at the seeds, and a pixel's colour is its root's.  Three passes:
//   1. label_tile_kernel: one CTA per 64x32 tile.  Parents from the arrival times (tile + 1-pixel ring in
//      shared memory), then pointer jumping INSIDE the tile in shared memory: every pixel ends up with
//      either its final colour (the chain reached a seed in the tile) or the index of the pixel OUTSIDE
//      the tile where its chain leaves -- a pixel on the rim of a neighbouring tile.  Also writes the
//      level bytes.
//   2. rim_jump_kernel: pointer jumping over the RIM pixels only (9 % of the image; the only possible
//      targets of pass 1), persistent and cooperative, until they all hold colours.  Chains are now
//      counted in tiles, not pixels.
//   3. label_finish_kernel: everything that still points at a rim pixel takes that pixel's word.
// A full-image pointer-jumping loop (the first version) re-read the whole label plane log2(longest
// chain) times: 13-14 passes on smooth fields.
// Row strips: a halo-row pixel belongs to the neighbouring strip; it holds a pointer to itself
// ("pending") until that strip's colour arrives, and whatever points at it keeps doing so.
#include "kernels.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace ws {

constexpr int LT_THREADS = 256;
constexpr int LT_W = TILE_W + 2;  // staged arrival times: tile + ring
constexpr int LT_H = TILE_H + 2;
constexpr uint16_t LT_TERMINAL = 0xFFFFu;

struct LabelSmem {
  uint32_t T[LT_H * LT_W];
  uint32_t term[TILE_H * TILE_W];  // final word of the pixel once nxt == LT_TERMINAL
  uint16_t nxt[TILE_H * TILE_W];   // next pixel of the chain inside the tile
};

__global__ void __launch_bounds__(LT_THREADS) label_tile_kernel(FloodBuffers b, ImageDims d) {
  __shared__ LabelSmem sm;
  const int tid = threadIdx.x;
  const int tpi = d.tiles_per_img();
  const int img = blockIdx.x / tpi;
  const int trem = blockIdx.x - img * tpi;
  const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
  const int r0 = ty * TILE_H, c0 = tx * TILE_W;
  const uint32_t* Tp = b.T + (size_t)img * d.t_plane();
  const int tp = d.t_pitch();
  // the padded layout keeps the ring in bounds (one extra row above / below, 4 words left / right)
  for (int i = tid; i < LT_H * LT_W; i += LT_THREADS) {
    const int r = i / LT_W, c = i - r * LT_W;
    sm.T[i] = __ldcg(Tp + (size_t)(r0 + r) * tp + (c0 + c - 1 + T_PAD_L));  // image (r0 + r - 1, c0 + c - 1)
  }
  __syncthreads();

  const int lc = tid % TILE_W, g = tid / TILE_W;
  const size_t base = (size_t)img * d.px_per_img();
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const int lr = g * ROWS_PER_THREAD + i;
    const int li = lr * TILE_W + lc;
    const int r = r0 + lr, c = c0 + lc;
    uint32_t term = LAB_RESOLVED;  // UNCOLOURED
    uint16_t nx = LT_TERMINAL;
    if (r < d.rows && c < d.cols) {
      const uint32_t* t = sm.T + (lr + 1) * LT_W + lc + 1;
      const uint32_t tv = t[0];
      const size_t p = base + (size_t)r * d.cols + c;
      b.lvl[p] = (tv >= T_INF) ? (uint8_t)255 : (uint8_t)(tv >> 24);
      if (tv >= T_INF) {
        // never coloured
      } else if (d.is_halo_row(r)) {
        term = (uint32_t)p;  // a neighbouring strip owns this pixel: pending
      } else if (tv == 0u) {
        term = __ldcg(b.lab + p);  // seed: coloured by seed_init
      } else {
        // A coloured non-seed pixel is interior, so all four neighbours exist.
        int dr, dc;
        if (t[LT_W] < tv) { dr = 1; dc = 0; }
        else if (t[1] < tv) { dr = 0; dc = 1; }
        else if (t[-1] < tv) { dr = 0; dc = -1; }
        else if (t[-LT_W] < tv) { dr = -1; dc = 0; }
        else { dr = 0; dc = 0; atomicOr(&b.ctrl[FC_ERROR], 4u); }  // cannot happen at a fixed point
        const int pr = lr + dr, pc = lc + dc;
        if (dr == 0 && dc == 0) {
          // leave UNCOLOURED
        } else if (pr >= 0 && pr < TILE_H && pc >= 0 && pc < TILE_W) {
          nx = (uint16_t)(pr * TILE_W + pc);
        } else {
          term = (uint32_t)(p + (ptrdiff_t)dr * d.cols + dc);  // the chain leaves the tile here
        }
      }
    }
    sm.term[li] = term;
    sm.nxt[li] = nx;
  }
  __syncthreads();

  // pointer jumping inside the tile (reads and writes separated by barriers)
  for (;;) {
    uint16_t n2[ROWS_PER_THREAD];
    uint32_t t2[ROWS_PER_THREAD];
    bool moving = false;
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      const int li = (g * ROWS_PER_THREAD + i) * TILE_W + lc;
      const uint16_t n = sm.nxt[li];
      n2[i] = LT_TERMINAL;
      t2[i] = 0;
      if (n != LT_TERMINAL) {
        n2[i] = sm.nxt[n];
        t2[i] = sm.term[n];
        moving = true;
      }
    }
    if (!__syncthreads_or(moving)) break;
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      const int li = (g * ROWS_PER_THREAD + i) * TILE_W + lc;
      if (sm.nxt[li] != LT_TERMINAL) {
        sm.nxt[li] = n2[i];
        if (n2[i] == LT_TERMINAL) sm.term[li] = t2[i];
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const int lr = g * ROWS_PER_THREAD + i;
    const int r = r0 + lr, c = c0 + lc;
    if (r < d.rows && c < d.cols) __stcg(b.lab + base + (size_t)r * d.cols + c, sm.term[lr * TILE_W + lc]);
  }
}

cudaError_t launch_parent(FloodBuffers b, ImageDims d, cudaStream_t s) {
  label_tile_kernel<<<d.tiles_total(), LT_THREADS, 0, s>>>(b, d);
  return cudaGetLastError();
}

// ---- rim pixels --------------------------------------------------------------------------------

constexpr int RIM_PER_TILE = 2 * TILE_W + 2 * (TILE_H - 2);  // 188

// k-th rim pixel of a tile -> tile-local (row, col)
__device__ __forceinline__ void rim_coords(int k, int& lr, int& lc) {
  if (k < TILE_W) { lr = 0; lc = k; }
  else if (k < 2 * TILE_W) { lr = TILE_H - 1; lc = k - TILE_W; }
  else if (k < 2 * TILE_W + TILE_H - 2) { lr = 1 + (k - 2 * TILE_W); lc = 0; }
  else { lr = 1 + (k - 2 * TILE_W - (TILE_H - 2)); lc = TILE_W - 1; }
}

// Pointer jumping over the rim pixels: lab[p] <- lab[lab[p]] until every rim word is a colour (or points
// at a pending halo pixel, which points at itself).  In place and racy on purpose: whatever a thread reads
// is a valid ancestor or the final colour.
__global__ void __launch_bounds__(256) rim_jump_kernel(uint32_t* __restrict__ lab, ImageDims d, uint32_t* ctrl) {
  cg::grid_group grid = cg::this_grid();
  const size_t total = (size_t)d.tiles_total() * RIM_PER_TILE;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int tpi = d.tiles_per_img();
  for (uint32_t round = 0;; ++round) {
    const int cur = round % 3;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st_cg(&ctrl[FC_JUMP_FLAG0 + (round + 1) % 3], 0u);
      atomicAdd(&ctrl[FC_JUMP_ROUNDS], 1u);
    }
    int pending = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const int tile = (int)(i / RIM_PER_TILE), k = (int)(i - (size_t)tile * RIM_PER_TILE);
      int lr, lc;
      rim_coords(k, lr, lc);
      const int img = tile / tpi, trem = tile - img * tpi;
      const int ty = trem / d.tiles_x, tx = trem - ty * d.tiles_x;
      const int r = ty * TILE_H + lr, c = tx * TILE_W + lc;
      if (r >= d.rows || c >= d.cols) continue;
      uint32_t* w = lab + (size_t)img * d.px_per_img() + (size_t)r * d.cols + c;
      const uint32_t v = ld_cg(w);
      if (v & LAB_RESOLVED) continue;
      const uint32_t t = ld_cg(lab + v);
      if (t == v) continue;  // the chain ends at a pending halo pixel of a strip
      st_cg(w, t);
      if (!(t & LAB_RESOLVED)) pending = 1;
    }
    if (__syncthreads_or(pending) && threadIdx.x == 0) st_cg(&ctrl[FC_JUMP_FLAG0 + cur], 1u);
    grid.sync();
    if (ld_cg(&ctrl[FC_JUMP_FLAG0 + cur]) == 0u) break;
  }
}

// Every word that is still a pointer takes the word of the (rim or pending) pixel it points at.
__global__ void __launch_bounds__(256) label_finish_kernel(uint32_t* __restrict__ lab, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n / 4;
  uint4* lab4 = reinterpret_cast<uint4*>(lab);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 v = __ldcg(lab4 + i);
    if ((v.x & v.y & v.z & v.w) & LAB_RESOLVED) continue;
    uint4 o;
    o.x = (v.x & LAB_RESOLVED) ? v.x : ld_cg(lab + v.x);
    o.y = (v.y & LAB_RESOLVED) ? v.y : ld_cg(lab + v.y);
    o.z = (v.z & LAB_RESOLVED) ? v.z : ld_cg(lab + v.z);
    o.w = (v.w & LAB_RESOLVED) ? v.w : ld_cg(lab + v.w);
    if (o.x != v.x || o.y != v.y || o.z != v.z || o.w != v.w) __stcg(lab4 + i, o);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t v = ld_cg(lab + i);
    if (!(v & LAB_RESOLVED)) {
      const uint32_t o = ld_cg(lab + v);
      if (o != v) st_cg(lab + i, o);
    }
  }
}

static int coop_max_grid_rim(int device) {
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)rim_jump_kernel, 256, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}
int jump_max_grid(int device) { return coop_max_grid_rim(device); }

cudaError_t launch_jump(FloodBuffers b, ImageDims d, int grid, cudaStream_t s) {
  uint32_t* lab = b.lab;
  uint32_t* ctrl = b.ctrl;
  const size_t total = (size_t)d.tiles_total() * RIM_PER_TILE;
  const size_t want = (total + 255) / 256;
  const int g = (size_t)grid > want ? (int)(want ? want : 1) : grid;
  void* args[] = {&lab, &d, &ctrl};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)rim_jump_kernel, dim3(g), dim3(256), args, 0, s);
  if (e != cudaSuccess) return e;
  const size_t n = d.px_total();
  const size_t w2 = (n / 4 + 255) / 256;
  const unsigned g2 = (unsigned)(w2 < (size_t)148 * 16 ? (w2 ? w2 : 1) : (size_t)148 * 16);
  label_finish_kernel<<<g2, 256, 0, s>>>(lab, n);
  return cudaGetLastError();
}

}  // namespace ws
