// labels.cu -- K3: from arrival times to the reference's colours.
//
// The reference colours a flooded pixel with the colour of the coloured neighbours it sees at that
// moment (lib.rs:235-255): exactly the neighbours q with T(q) < T(p); canonical tie-break = the first of
// them in the order down, right, left, up (lib.rs:190, 245, `col0`).  One parent per pixel gives a forest
// rooted at the seeds, and a pixel's colour is its root's.  Three passes:
//   1. label_tile_kernel: one CTA per 64x32 tile.  Parents from the arrival times (tile + 1-pixel ring,
//      staged in shared memory by the bulk-copy engine), then pointer jumping INSIDE the tile in shared
//      memory: every pixel ends up with either its final colour (the chain reached a seed in the tile)
//      or a reference to the pixel OUTSIDE the tile where its chain leaves -- always a pixel on the rim
//      of a neighbouring tile.  Rim pixels are also written to a compact array (188 words per tile),
//      and references are indices into that array.  Also writes the level bytes.
//   2. rim_jump_kernel: pointer jumping over the compact rim array only (9 % of the image, L2-resident),
//      persistent and cooperative, until every entry holds a colour.  Chains are now counted in tiles.
//   3. label_finish_kernel: every word that is still a reference takes the rim entry it refers to.
// A full-image pointer-jumping loop (the first version) re-read the whole label plane log2(longest
// chain) times: 13-14 passes on smooth fields.
// Row strips: a halo-row pixel belongs to the neighbouring strip; it has a slot behind the rim entries
// that holds its own index ("pending") until that strip's colour is imported, and whatever refers to it
// keeps doing so.
#include "kernels.cuh"

#include <cooperative_groups.h>
#include <cuda.h>  // CUtensorMap (type only)

#include <cstring>

namespace cg = cooperative_groups;

namespace ws {

constexpr int LT_THREADS = 256;
constexpr int LT_W = STG_W;       // staged arrival times: rows of 72 words (image columns c0-4 .. c0+67)
constexpr int LT_H = TILE_H + 2;
constexpr int LT_C0 = T_PAD_L;    // staged column of the tile's column 0
// working word of a pixel: LT_LOCAL | BYTE offset of the next pixel of its chain inside the tile's word array;
// anything else is final: a colour (bit 31 set: negative as a signed word) or a rim reference (< 2^28).  In-tile
// pointers are the only words in [2^30, 2^31): one signed compare tells them apart.
constexpr uint32_t LT_LOCAL = 0x40000000u;
constexpr uint32_t LT_OFF_MASK = (TILE_H * TILE_W - 1) * 4;
__device__ __forceinline__ bool lt_is_local(uint32_t w) { return (int32_t)w >= (int32_t)LT_LOCAL; }
constexpr int RIM_PER_TILE = 2 * TILE_W + 2 * (TILE_H - 2);  // 188

size_t rim_words(const ImageDims& d) { return (size_t)d.tiles_total() * RIM_PER_TILE + 2 * (size_t)d.cols; }

// k-th rim pixel of a tile <-> tile-local (row, col)
__device__ __forceinline__ int rim_index(int lr, int lc) {
  if (lr == 0) return lc;
  if (lr == TILE_H - 1) return TILE_W + lc;
  if (lc == 0) return 2 * TILE_W + lr - 1;
  return 2 * TILE_W + (TILE_H - 2) + lr - 1;  // lc == TILE_W - 1
}

struct __align__(128) LabelSmem {
  uint32_t T[LT_H * LT_W];
  uint32_t w[TILE_H * TILE_W];     // per pixel: LT_LOCAL | next pixel inside the tile, or its final word
  uint64_t bar;
  uint32_t rb[2 * TILE_H];        // sorted seed lists: rowbase of the tile's rows, one word per 32 columns
  uint32_t nseed_px;
};

// counter-based generator of the random tie-break: splitmix64 of (key, position of the pixel in the field)
__device__ __forceinline__ uint32_t tie_hash(uint64_t key, uint64_t idx) {
  uint64_t z = key + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

// The pixels of one thread (column lc, rows g*8 .. g*8+7): level byte, parent / colour word.  kPlain: the tile lies
// completely inside the image and holds no halo row of a strip -- no per-pixel bounds or ownership tests.
template <bool kTieRandom, bool kPlain>
__device__ __forceinline__ uint32_t lt_parents(LabelSmem& sm, const FloodBuffers& b, const ImageDims& d, int img, int r0,
                                               int c0, int tile, bool sorted, uint32_t col_first, uint64_t tie_seed) {
  const int tid = threadIdx.x;
  const int lc = tid % TILE_W, g = tid / TILE_W;
  const size_t base = (size_t)img * d.px_per_img();
  const uint32_t rim_total = (uint32_t)d.tiles_total() * RIM_PER_TILE;
  const int halo_a = d.halo_top ? 0 : -1, halo_b = d.halo_bottom ? d.rows - 1 : -1;   // this plan's halo rows
  // byte offsets (relative to a pixel's own word) that leave the tile from this thread's column / first / last row
  constexpr int NEVER = 0x7FFFFFFF;
  const int off_r = lc == TILE_W - 1 ? 4 : NEVER, off_l = lc == 0 ? -4 : NEVER;
  const int off_d = g == TILE_H / ROWS_PER_THREAD - 1 ? TILE_W * 4 : NEVER, off_u = g == 0 ? -TILE_W * 4 : NEVER;
  const uint32_t lane_lt = (1u << (tid & 31)) - 1u;
  uint32_t nseed_warp = 0;  // owned pixels of the warp's rows that hold a seed (arrival time 0)
  uint8_t* lvl_p = b.lvl + base + (size_t)(r0 + g * ROWS_PER_THREAD) * d.cols + c0 + lc;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const int lr = g * ROWS_PER_THREAD + i;
    const int r = r0 + lr, c = c0 + lc;
    uint32_t term = LAB_RESOLVED;  // UNCOLOURED
    const uint32_t* t = sm.T + (lr + 1) * LT_W + lc + LT_C0;
    const uint32_t tv = t[0];
    const bool inside = kPlain || (r < d.rows && c < d.cols);
    const bool halo = !kPlain && (r == halo_a || r == halo_b);
    const uint32_t seeds_here = __ballot_sync(0xffffffffu, inside && tv == 0u && !halo);
    nseed_warp += (uint32_t)__popc(seeds_here);
    if (inside) {
      lvl_p[(size_t)i * d.cols] = (uint8_t)(tv >> 24);   // T >= T_INF ("never") reads as level 255
      if (halo) {
        // a neighbouring strip owns this pixel: its slot behind the rim entries holds its own index ("pending")
        // until that strip's colour is imported; a pixel that is never coloured is resolved (UNCOLOURED) at once
        const uint32_t slot = rim_total + (r == 0 ? 0u : (uint32_t)d.cols) + (uint32_t)c;
        if (tv < T_INF) term = slot;
        b.rim[slot] = term;
      } else if (tv >= T_INF) {
        // never coloured
      } else if (tv == 0u) {
        if (sorted) term = LAB_RESOLVED | (col_first + sm.rb[2 * lr + (lc >> 5)] + (uint32_t)__popc(seeds_here & lane_lt));
        else term = __ldcg(b.lab + base + (size_t)r * d.cols + c);  // seed: coloured by seed_init
      } else {
        // A coloured non-seed pixel is interior, so all four neighbours exist.  Byte offset of the parent's word:
        int off = 0;
        const uint32_t dn = t[LT_W], rt = t[1], lf = t[-1], up = t[-LT_W];
        if (kTieRandom) {
          // the reference draws uniformly among the coloured neighbours, in the order down, right, left, up
          const uint32_t em = (dn < tv ? 1u : 0u) | (rt < tv ? 2u : 0u) | (lf < tv ? 4u : 0u) | (up < tv ? 8u : 0u);
          const int k = __popc(em);
          if (k) {
            int pick = 0;
            if (k > 1) {
              const uint64_t pos = ((uint64_t)img * (uint64_t)d.global_rows + (uint64_t)(r + d.row_offset)) * (uint64_t)d.cols + (uint64_t)c;
              pick = (int)(((uint64_t)tie_hash(tie_seed, pos) * (uint64_t)k) >> 32);
            }
            const int bit = __fns(em, 0, pick + 1);  // position of the pick-th set bit
            off = bit == 0 ? TILE_W * 4 : (bit == 1 ? 4 : (bit == 2 ? -4 : -TILE_W * 4));
          }
        } else {
          // canonical: the first coloured neighbour in that order (the last assignment wins)
          off = up < tv ? -TILE_W * 4 : off;
          off = lf < tv ? -4 : off;
          off = rt < tv ? 4 : off;
          off = dn < tv ? TILE_W * 4 : off;
        }
        bool leaves = off == off_r || off == off_l;
        if (i == 0) leaves |= off == off_u;
        if (i == ROWS_PER_THREAD - 1) leaves |= off == off_d;
        if (off == 0) {
          atomicOr(&b.ctrl[FC_ERROR], 4u);  // cannot happen at a fixed point; the pixel stays UNCOLOURED
        } else if (!leaves) {
          term = (uint32_t)((int)(LT_LOCAL | (uint32_t)((lr * TILE_W + lc) * 4)) + off);
        } else {
          // the chain leaves the tile here, onto the rim of the neighbouring tile
          const int dr = off == TILE_W * 4 ? 1 : (off == -TILE_W * 4 ? -1 : 0), dc = off == 4 ? 1 : (off == -4 ? -1 : 0);
          const int ntile = tile + dr * d.tiles_x + dc;
          term = (uint32_t)ntile * RIM_PER_TILE + (uint32_t)rim_index((lr + dr) & (TILE_H - 1), (lc + dc) & (TILE_W - 1));
        }
      }
    }
    sm.w[lr * TILE_W + lc] = term;
  }
  return nseed_warp;
}

// one instruction per tile: 2-D tiled tensor copy global -> shared, completes on `bar` (the flood's map of the
// arrival-time plane: box STG_W x STG_H = LT_W x LT_H)
__device__ __forceinline__ void lt_tensor_g2s(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

template <bool kTieRandom>
__global__ void __launch_bounds__(LT_THREADS, 6) label_tile_kernel(const __grid_constant__ CUtensorMap tmT, FloodBuffers b,
                                                                   ImageDims d, uint32_t* __restrict__ ndistinct,
                                                                   uint64_t tie_seed) {
  __shared__ LabelSmem sm;
  const int tid = threadIdx.x;
  // grid = (tiles_x, tiles_y, slices) whenever that fits a grid (launch_parent): no divisions at the head of the CTA,
  // where they sat in front of the tile's copy
  int img, ty, tx;
  if (gridDim.y * gridDim.z > 1u || d.tiles_y * d.n_img == 1) {
    tx = (int)blockIdx.x;
    ty = (int)blockIdx.y;
    img = (int)blockIdx.z;
  } else {
    const int tpi = d.tiles_per_img();
    img = (int)blockIdx.x / tpi;
    const int trem = (int)blockIdx.x - img * tpi;
    ty = trem / d.tiles_x;
    tx = trem - ty * d.tiles_x;
  }
  const int tile = (img * d.tiles_y + ty) * d.tiles_x + tx;
  const int r0 = ty * TILE_H, c0 = tx * TILE_W;
  if (tid == 0) {
    // box rows r0-1 .. r0+32 (padded row index r0 .. r0+33), columns c0-4 .. c0+67; the flood's results were
    // written by atomics (generic proxy) in an earlier launch, so no cross-proxy fence is needed here
    sm.nseed_px = 0;
    mbar_init(&sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive_expect_tx(&sm.bar, LT_H * LT_W * 4);
    lt_tensor_g2s(sm.T, &tmT, c0, img * d.t_rows() + r0, &sm.bar);
  }
  // Sorted seed list (flood.cu, fill_rows_kernel): a seed's colour is its position in the list -- the index of
  // the row's first seed at or right of this warp's first column (rowbase, one entry per 32 columns), plus the
  // seeds to its left among the warp's 32 columns (a warp holds one row, 32 consecutive columns, per step).
  const bool sorted = __ldcg(&b.ctrl[FC_SEED_UNSORTED]) == 0u;
  if (tid < 2 * TILE_H) {
    const int row = r0 + (tid >> 1);
    sm.rb[tid] = (sorted && row < d.rows)
                     ? __ldg(b.rowbase + ((size_t)img * d.rows + row) * (size_t)(2 * d.tiles_x) + 2 * tx + (tid & 1))
                     : 0u;
  }
  __syncthreads();  // the barrier is initialised before anyone waits on it; rb[] is complete
  const int lc = tid % TILE_W, g = tid / TILE_W;
  const size_t base = (size_t)img * d.px_per_img();
  const bool plain = r0 + TILE_H <= d.rows && c0 + TILE_W <= d.cols && !(d.halo_top && ty == 0) &&
                     !(d.halo_bottom && r0 + TILE_H >= d.rows);
  // (loaded by every thread and used late: behind a barrier the load's latency was on every CTA's critical path)
  const uint32_t col_first = b.colour_base + 1u - __ldg(b.seed_off + img);   // colour of the slice's seed 0, minus its index
  mbar_wait(&sm.bar, 0);

  uint32_t nseed_warp;
  if (plain) nseed_warp = lt_parents<kTieRandom, true>(sm, b, d, img, r0, c0, tile, sorted, col_first, tie_seed);
  else nseed_warp = lt_parents<kTieRandom, false>(sm, b, d, img, r0, c0, tile, sorted, col_first, tie_seed);
  // (a later duplicate seed overwrites an earlier one, lib.rs:1365-1367, so a pixel counts once)
  if ((tid & 31) == 0 && nseed_warp) atomicAdd(&sm.nseed_px, nseed_warp);
  __syncthreads();
  if (tid == 0 && sm.nseed_px) atomicAdd(&ndistinct[img], sm.nseed_px);  // one global atomic per tile

  // Pointer jumping inside the tile, without barriers: every thread keeps replacing the words of its pixels by
  // their successors' CURRENT words until none of them is an in-tile pointer any more.  A word only ever changes
  // from a pointer to a pointer further down its own chain or to the chain's final word, so whatever a racing
  // reader sees is a valid ancestor; chains shorten for everybody as results land (rounds of doubling separated
  // by barriers cost two barriers and eight bit tests per round and thread, whether or not it had work left).
  const char* wbytes = reinterpret_cast<const char*>(sm.w);
  uint32_t* mine = sm.w + g * ROWS_PER_THREAD * TILE_W + lc;
  uint32_t cur[ROWS_PER_THREAD];
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) cur[i] = mine[i * TILE_W];
  for (;;) {
    int top = (int)0x80000000;  // in-tile pointers are the largest words as signed numbers
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      if (lt_is_local(cur[i])) {
        cur[i] = *reinterpret_cast<const volatile uint32_t*>(wbytes + (cur[i] & LT_OFF_MASK));
        *reinterpret_cast<volatile uint32_t*>(mine + i * TILE_W) = cur[i];
      }
      top = max(top, (int)cur[i]);
    }
    if (top < (int)LT_LOCAL) break;
  }
  __syncthreads();  // the rim entries below are other threads' words

  if (plain) {
    uint32_t* out = b.lab + base + (size_t)(r0 + g * ROWS_PER_THREAD) * d.cols + c0 + lc;
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) __stcg(out + (size_t)i * d.cols, mine[i * TILE_W]);
  } else {
#pragma unroll
    for (int i = 0; i < ROWS_PER_THREAD; ++i) {
      const int lr = g * ROWS_PER_THREAD + i;
      const int r = r0 + lr, c = c0 + lc;
      if (r < d.rows && c < d.cols) __stcg(b.lab + base + (size_t)r * d.cols + c, sm.w[lr * TILE_W + lc]);
    }
  }
  // the tile's rim, compact (pixels outside the image: UNCOLOURED, so that every entry is defined)
  if (tid < RIM_PER_TILE) {
    int lr, lcc;
    if (tid < TILE_W) { lr = 0; lcc = tid; }
    else if (tid < 2 * TILE_W) { lr = TILE_H - 1; lcc = tid - TILE_W; }
    else if (tid < 2 * TILE_W + TILE_H - 2) { lr = 1 + (tid - 2 * TILE_W); lcc = 0; }
    else { lr = 1 + (tid - 2 * TILE_W - (TILE_H - 2)); lcc = TILE_W - 1; }
    const bool in = (r0 + lr < d.rows) && (c0 + lcc < d.cols);
    __stcg(b.rim + (size_t)tile * RIM_PER_TILE + tid, in ? sm.w[lr * TILE_W + lcc] : LAB_RESOLVED);
  }
}

cudaError_t launch_parent(FloodBuffers b, ImageDims d, uint32_t* ndistinct, bool tie_random, uint64_t tie_seed,
                          const void* tensor_maps, cudaStream_t s) {
  CUtensorMap tmT;
  memcpy(&tmT, tensor_maps, sizeof(CUtensorMap));  // the arrival-time plane's map (flood_make_tensor_maps)
  cudaError_t e = cudaMemsetAsync(ndistinct, 0, sizeof(uint32_t) * (size_t)d.n_img, s);
  if (e != cudaSuccess) return e;
  // the pending slots of halo rows this plan does not have read as resolved (bit 31 set) to whoever counts them
  e = cudaMemsetAsync(b.rim + (size_t)d.tiles_total() * RIM_PER_TILE, 0x80, 2 * (size_t)d.cols * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  const dim3 grid = (d.tiles_y <= 65535 && d.n_img <= 65535) ? dim3(d.tiles_x, d.tiles_y, d.n_img) : dim3(d.tiles_total());
  if (tie_random) label_tile_kernel<true><<<grid, LT_THREADS, 0, s>>>(tmT, b, d, ndistinct, tie_seed);
  else label_tile_kernel<false><<<grid, LT_THREADS, 0, s>>>(tmT, b, d, ndistinct, 0ull);
  return cudaGetLastError();
}

// Pointer jumping over the compact rim array: rim[i] <- rim[rim[i]] until every entry is a colour (or refers
// to a pending halo slot, which refers to itself).  In place and racy on purpose: whatever a thread reads
// is a valid ancestor or the final colour.
__global__ void __launch_bounds__(256) rim_jump_kernel(uint32_t* __restrict__ rim, size_t n, uint32_t* ctrl) {
  cg::grid_group grid = cg::this_grid();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (uint32_t round = 0;; ++round) {
    const int cur = round % 3;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      st_cg(&ctrl[FC_JUMP_FLAG0 + (round + 1) % 3], 0u);
      atomicAdd(&ctrl[FC_JUMP_ROUNDS], 1u);
    }
    int pending = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const uint32_t v = ld_cg(rim + i);
      if (v & LAB_RESOLVED) continue;
      const uint32_t t = ld_cg(rim + v);
      if (t == v) continue;  // the chain ends at a pending halo pixel of a strip
      st_cg(rim + i, t);
      if (!(t & LAB_RESOLVED)) pending = 1;
    }
    if (__syncthreads_or(pending) && threadIdx.x == 0) st_cg(&ctrl[FC_JUMP_FLAG0 + cur], 1u);
    grid.sync();
    if (ld_cg(&ctrl[FC_JUMP_FLAG0 + cur]) == 0u) break;
  }
}

// Every label word that is still a reference takes the rim entry (or pending slot) it refers to.
__global__ void __launch_bounds__(256) label_finish_kernel(uint32_t* __restrict__ lab, size_t n,
                                                           const uint32_t* __restrict__ rim) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n4 = n / 4;
  uint4* lab4 = reinterpret_cast<uint4*>(lab);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint4 v = __ldcg(lab4 + i);
    if ((v.x & v.y & v.z & v.w) & LAB_RESOLVED) continue;
    uint4 o;
    o.x = (v.x & LAB_RESOLVED) ? v.x : ld_cg(rim + v.x);
    o.y = (v.y & LAB_RESOLVED) ? v.y : ld_cg(rim + v.y);
    o.z = (v.z & LAB_RESOLVED) ? v.z : ld_cg(rim + v.z);
    o.w = (v.w & LAB_RESOLVED) ? v.w : ld_cg(rim + v.w);
    if (o.x != v.x || o.y != v.y || o.z != v.z || o.w != v.w) __stcg(lab4 + i, o);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t v = ld_cg(lab + i);
    if (!(v & LAB_RESOLVED)) {
      const uint32_t o = ld_cg(rim + v);
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

cudaError_t launch_label_finish(FloodBuffers b, ImageDims d, int sms, cudaStream_t s) {
  const size_t n = d.px_total();
  const size_t w2 = (n / 4 + 255) / 256;
  const size_t cap = (size_t)sms * 16;
  const unsigned g2 = (unsigned)(w2 < cap ? (w2 ? w2 : 1) : cap);
  label_finish_kernel<<<g2, 256, 0, s>>>(b.lab, n, b.rim);
  return cudaGetLastError();
}

// finish = 0: only the rim array is resolved (strips between two exchange rounds: the label plane is finished
// once, after the last round -- it is 10x the rim array)
cudaError_t launch_jump(FloodBuffers b, ImageDims d, int grid, int finish, cudaStream_t s) {
  uint32_t* rim = b.rim;
  uint32_t* ctrl = b.ctrl;
  size_t total = (size_t)d.tiles_total() * RIM_PER_TILE;  // the pending slots never move by themselves
  const size_t want = (total + 255) / 256;
  const int g = (size_t)grid > want ? (int)(want ? want : 1) : grid;
  void* args[] = {&rim, &total, &ctrl};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)rim_jump_kernel, dim3(g), dim3(256), args, 0, s);
  if (e != cudaSuccess || !finish) return e;
  const size_t n = d.px_total();
  const size_t w2 = (n / 4 + 255) / 256;
  const size_t g2cap = (size_t)num_sms() * 16;
  const unsigned g2 = (unsigned)(w2 < g2cap ? (w2 ? w2 : 1) : g2cap);
  label_finish_kernel<<<g2, 256, 0, s>>>(b.lab, n, rim);
  return cudaGetLastError();
}

}  // namespace ws
