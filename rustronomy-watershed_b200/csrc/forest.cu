// forest.cu -- K4b: the minimum spanning forest of the basin graph that the tiles left over, by Boruvka rounds
// in global memory (make_colour_map's closure over all water levels at once, lib.rs:467-542).
//
// merge_reduce (merge.cu) emits, per tile, FINAL forest edges (only counted) and DEFERRED edges between the
// identities of basins that reach the tile's rim.  The lakes per level need, of the DEFERRED graph, nothing but
// the number of forest edges at every level:
//     lakes(L) = colours on the canvas - FINAL edges with level <= L - forest edges of the DEFERRED graph <= L.
// The first version ran Kruskal: a lock-free union-find over the edges bucketed by level, one grid barrier per
// level -- 255 barriers with ~3e4 edges between two of them, latency bound (2.2 ms at 16384^2).  Boruvka needs
// no order: every component offers its lightest edge with a 64-bit atomicMin, picked edges ARE forest edges
// with their true level, and ~log2(largest component) rounds of two grid barriers replace the 255.
//
// The rounds follow merge.cu's rules exactly (tests/merge_model.py restates them):
//   * a pick made by a component of CLOSED nodes only is FINAL, any other pick is DEFERRED;
//   * of a mutual pick the closed side moves if only one is closed, else the larger root;
//   * a closed mover takes the identity of the node at the far end of its edge (link[]).
// On one GPU every node is closed -- the edge list is the whole graph -- so every pick is FINAL and the result
// is the histogram of forest-edge levels.  In a row strip (engine.cu, ws_plan_strip_forest) the basins that
// touch the strip's boundary rows are open: what comes out is the strip's FINAL histogram and a short list of
// DEFERRED edges between boundary basins, the only thing the strips have to exchange (SURVEY.md 8(e):
// "boundary union-merge").
#include "kernels.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace ws {

constexpr unsigned long long FB_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t FOREST_MAX_ROUNDS = 250;   // (the round index lives in 8 bits of the offer key)

// Everything a phase reads was written by other CTAs in the phase before (same kernel, across a grid barrier):
// all of it goes through L2 (ld.cg / st.cg), never through the read-only path.
__device__ __forceinline__ uint32_t ld_u32(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ uint2 ld_u2(const uint2* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ld_u8(const uint8_t* p) { return (uint32_t)__ldcg(p); }
__device__ __forceinline__ unsigned long long ld_u64(const unsigned long long* p) { return __ldcg(p); }

// root of x: parent[] does not change while this runs (hooks are written in the other phase)
__device__ __forceinline__ uint32_t forest_root(const uint32_t* parent, uint32_t x) {
  for (uint32_t p = ld_u32(parent + x); p != x; p = ld_u32(parent + x)) x = p;
  return x;
}

// Offers of round r always beat what older rounds left in best[] (no reset, one array); inside a round the
// order is (level, position in the live list).
__device__ __forceinline__ unsigned long long forest_key(uint32_t round, uint32_t w, uint32_t pos) {
  return ((unsigned long long)(255u - round) << 56) | ((unsigned long long)w << 32) | pos;
}

__device__ __forceinline__ int forest_slice_of(const uint32_t* __restrict__ seed_off, int n_img, uint32_t colour) {
  int s0 = 0, s1 = n_img;
  while (s1 - s0 > 1) {
    const int mid = (s0 + s1) >> 1;
    if (__ldg(seed_off + mid) <= colour) s0 = mid; else s1 = mid;
  }
  return s0;
}

__global__ void __launch_bounds__(256) forest_init_kernel(ForestBuffers f, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    f.parent[i] = i;
    f.link[i] = i;
    f.best[i] = FB_NONE;
  }
}

// open_[colour id] = 1 for the colours of `count` label words (a boundary row of a strip)
__global__ void __launch_bounds__(256) forest_mark_open_kernel(const uint32_t* __restrict__ lab, size_t count,
                                                               uint32_t ncolours, uint8_t* __restrict__ open_) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint32_t v = lab[i] & LAB_MASK;
  if (v == 0u) return;
  if (v - 1u < ncolours) open_[v - 1u] = 1;   // labels are colours (>= 1); ids are 0-based
}

struct ForestArgs {
  ForestBuffers f;
  const uint2* in_ab;          // edges to start from (merge_reduce's list, or the strips' gathered picks)
  const uint8_t* in_w;
  const uint32_t* in_count;
  int skip_final;              // 1: entries with bit 31 of .y set are FINAL tile edges -- counted into tile_hist
  const uint32_t* seed_off;    // batches: slice of a colour (histograms are per slice)
  int n_img;
};

// A CTA works on chunks of FK * 256 consecutive entries, FK per thread with all their loads in flight (the
// phases are chains of dependent L2 / DRAM round trips: latency is all there is to hide), and appends what stays
// alive with ONE global atomic per chunk (a warp-level atomicAdd per 32 entries on a single counter serialised
// the whole pass: 1.1 M same-address atomics, ~3 ms).
constexpr int FK = 4;
constexpr uint32_t FCHUNK = FK * 256;

struct ForestCta {
  uint32_t cnt, base;
};

// positions for the calling thread's live items inside the output list; all threads of the CTA call this
__device__ __forceinline__ void cta_append(ForestCta& sc, uint32_t* counter, const bool (&live)[FK], uint32_t (&pos)[FK]) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) sc.cnt = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < FK; ++k) {
    const uint32_t m = __ballot_sync(0xffffffffu, live[k]);
    uint32_t wb = 0;
    if (m) {
      const int leader = __ffs((int)m) - 1;
      if (lane == leader) wb = atomicAdd(&sc.cnt, (uint32_t)__popc(m));
      wb = __shfl_sync(0xffffffffu, wb, leader);
    }
    pos[k] = wb + (uint32_t)__popc(m & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (threadIdx.x == 0) sc.base = sc.cnt ? atomicAdd(counter, sc.cnt) : 0u;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < FK; ++k) pos[k] += sc.base;
}

// kOpen = false: every node is closed (one GPU: the list is the whole graph) -- every pick is FINAL, no
// identities, the live list only carries the current roots of an edge's ends.
// kOpen = true (row strips): open_[] marks the boundary basins, the live list also carries the edge's own ends
// (a closed mover takes the identity of the BASIN at the far end, not of that basin's component).
template <bool kOpen>
__global__ void __launch_bounds__(256) forest_boruvka_kernel(const ForestArgs a) {
  cg::grid_group grid = cg::this_grid();
  const ForestBuffers& f = a.f;
  __shared__ uint32_t s_tile[256], s_forest[256];
  __shared__ ForestCta sc;
  s_tile[threadIdx.x] = 0;
  s_forest[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const bool one = a.n_img == 1;

  // ---- pass 0: every DEFERRED tile edge is alive (its ends are different identities) and offers itself
  {
    const uint32_t n = *a.in_count;
    for (uint32_t c0 = blockIdx.x * FCHUNK; c0 < n; c0 += gridDim.x * FCHUNK) {
      uint2 e[FK];
      uint32_t w[FK], pos[FK];
      bool live[FK];
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const uint32_t i = c0 + k * 256 + threadIdx.x;
        e[k] = make_uint2(0u, 0u);
        w[k] = 0;
        if (i < n) {
          e[k] = a.in_ab[i];
          w[k] = a.in_w[i];
        }
      }
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const uint32_t i = c0 + k * 256 + threadIdx.x;
        live[k] = false;
        if (i < n) {
          if (a.skip_final && (e[k].y >> 31)) {
            if (one) atomicAdd(&s_tile[w[k]], 1u);
            else atomicAdd(&f.tile_hist[(size_t)forest_slice_of(a.seed_off, a.n_img, e[k].x) * 256 + w[k]], 1u);
          } else {
            e[k].y &= LAB_MASK;
            live[k] = e[k].x != e[k].y;
          }
        }
      }
      cta_append(sc, &f.count[1], live, pos);
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        if (!live[k]) continue;
        __stcg(&f.ab[1][pos[k]], e[k]);
        __stcg(&f.w[1][pos[k]], (uint8_t)w[k]);
        if (kOpen) __stcg(&f.orig[1][pos[k]], e[k]);
        const unsigned long long key = forest_key(0u, w[k], pos[k]);
        atomicMin(&f.best[e[k].x], key);
        atomicMin(&f.best[e[k].y], key);
      }
    }
  }
  if (gtid == 0) f.count[0] = 0;
  grid.sync();

  int cur = 1;
  for (uint32_t round = 0;; ++round) {
    const uint32_t n = ld_u32(&f.count[cur]);
    if (n == 0) break;
    if (round >= FOREST_MAX_ROUNDS) {   // cannot happen (components at least halve every round)
      if (gtid == 0) atomicOr(f.error, 1u);
      break;
    }
    const uint2* ab = f.ab[cur];
    const uint8_t* wv = f.w[cur];
    // ---- hooks: the thread that holds a root's picked edge moves the root
    for (uint32_t c0 = blockIdx.x * FCHUNK; c0 < n; c0 += gridDim.x * FCHUNK) {
      uint2 e[FK];
      uint32_t w[FK];
      unsigned long long ba[FK], bb[FK];
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const uint32_t i = c0 + k * 256 + threadIdx.x;
        e[k] = make_uint2(0u, 0u);
        w[k] = 0;
        if (i < n) {
          e[k] = ld_u2(ab + i);          // the roots of the two ends at the start of this round
          w[k] = ld_u8(wv + i);
        }
      }
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const uint32_t i = c0 + k * 256 + threadIdx.x;
        ba[k] = bb[k] = FB_NONE;
        if (i < n) {
          ba[k] = ld_u64(f.best + e[k].x);
          bb[k] = ld_u64(f.best + e[k].y);
        }
      }
#pragma unroll
      for (int k = 0; k < FK; ++k) {
        const uint32_t i = c0 + k * 256 + threadIdx.x;
        if (i >= n) continue;
        const unsigned long long key = forest_key(round, w[k], i);
        const uint32_t ca = e[k].x, cb = e[k].y;
        const bool pa = ba[k] == key, pb = bb[k] == key;
        if (!pa && !pb) continue;
        bool cla = true, clb = true;            // open_ of a root is current
        if (kOpen) {
          cla = !ld_u8(f.open_ + ca);
          clb = !ld_u8(f.open_ + cb);
        }
        bool a_moves;
        if (pa && pb) a_moves = cla == clb ? ca > cb : cla;   // mutual: the closed side moves if only one is closed,
        else a_moves = pa;                                     // else the larger root
        const uint32_t mover = a_moves ? ca : cb, other = a_moves ? cb : ca;
        const bool mover_closed = a_moves ? cla : clb;
        const bool fin = mover_closed || (pa && pb && (cla || clb));
        __stcg(f.parent + mover, other);
        if (fin) {
          if (kOpen && mover_closed) {
            const uint2 o = ld_u2(f.orig[cur] + i);
            __stcg(f.link + mover, a_moves ? o.y : o.x);   // the basin at the far end
          }
          if (one) atomicAdd(&s_forest[w[k]], 1u);
          else atomicAdd(&f.forest_hist[(size_t)forest_slice_of(a.seed_off, a.n_img, ca) * 256 + w[k]], 1u);
        } else if (kOpen) {
          const uint32_t d = atomicAdd(f.n_deferred, 1u);
          if (d < f.def_cap) {
            __stcg(&f.def_ab[d], ld_u2(f.orig[cur] + i));
            __stcg(&f.def_w[d], (uint8_t)w[k]);
          } else {
            atomicOr(f.error, 2u);
          }
        }
      }
    }
    if (gtid == 0) {
      __stcg(&f.count[cur ^ 1], 0u);
      atomicAdd(f.rounds, 1u);
    }
    grid.sync();
    // ---- new roots, liveness, offers of the next round (parent[] is fixed during this pass)
    {
      uint2* oab = f.ab[cur ^ 1];
      uint8_t* ow = f.w[cur ^ 1];
      for (uint32_t c0 = blockIdx.x * FCHUNK; c0 < n; c0 += gridDim.x * FCHUNK) {
        uint2 e[FK], o[FK];
        uint32_t w[FK], ra[FK], rb[FK], pa[FK], pb[FK], pos[FK];
        bool live[FK];
#pragma unroll
        for (int k = 0; k < FK; ++k) {
          const uint32_t i = c0 + k * 256 + threadIdx.x;
          e[k] = o[k] = make_uint2(0u, 0u);
          w[k] = 0;
          if (i < n) {
            e[k] = ld_u2(ab + i);
            w[k] = ld_u8(wv + i);
            if (kOpen) o[k] = ld_u2(f.orig[cur] + i);
          }
        }
#pragma unroll
        for (int k = 0; k < FK; ++k) {   // the first hop of all walks together
          const uint32_t i = c0 + k * 256 + threadIdx.x;
          ra[k] = e[k].x;
          rb[k] = e[k].y;
          pa[k] = ra[k];
          pb[k] = rb[k];
          if (i < n) {
            pa[k] = ld_u32(f.parent + ra[k]);
            pb[k] = ld_u32(f.parent + rb[k]);
          }
        }
        for (bool more = true; more;) {  // then in lockstep until every walk has reached its root
          more = false;
#pragma unroll
          for (int k = 0; k < FK; ++k) {
            if (pa[k] != ra[k]) { ra[k] = pa[k]; pa[k] = ld_u32(f.parent + ra[k]); more = true; }
            if (pb[k] != rb[k]) { rb[k] = pb[k]; pb[k] = ld_u32(f.parent + rb[k]); more = true; }
          }
        }
#pragma unroll
        for (int k = 0; k < FK; ++k) {
          const uint32_t i = c0 + k * 256 + threadIdx.x;
          live[k] = false;
          if (i < n) {
            if (kOpen) {   // (only ever set: racing writers store the same value)
              if (ra[k] != e[k].x && ld_u8(f.open_ + e[k].x)) __stcg(f.open_ + ra[k], (uint8_t)1);
              if (rb[k] != e[k].y && ld_u8(f.open_ + e[k].y)) __stcg(f.open_ + rb[k], (uint8_t)1);
            }
            live[k] = ra[k] != rb[k];
          }
        }
        cta_append(sc, &f.count[cur ^ 1], live, pos);
#pragma unroll
        for (int k = 0; k < FK; ++k) {
          if (!live[k]) continue;
          __stcg(&oab[pos[k]], make_uint2(ra[k], rb[k]));
          __stcg(&ow[pos[k]], (uint8_t)w[k]);
          if (kOpen) __stcg(&f.orig[cur ^ 1][pos[k]], o[k]);
          const unsigned long long key = forest_key(round + 1u, w[k], pos[k]);
          atomicMin(&f.best[ra[k]], key);
          atomicMin(&f.best[rb[k]], key);
        }
      }
    }
    grid.sync();
    cur ^= 1;
  }

  // ---- histograms of this CTA
  __syncthreads();
  if (one) {
    if (s_tile[threadIdx.x]) atomicAdd(&f.tile_hist[threadIdx.x], s_tile[threadIdx.x]);
    if (s_forest[threadIdx.x]) atomicAdd(&f.forest_hist[threadIdx.x], s_forest[threadIdx.x]);
  }
}

// DEFERRED picks between identities: every end follows its chain of FINAL moves (link[] is final now)
__global__ void __launch_bounds__(256) forest_ident_kernel(ForestBuffers f) {
  const uint32_t n = min(*f.n_deferred, f.def_cap);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint2 e = f.def_ab[i];
    int guard = 1 << 22;   // chains of FINAL moves are acyclic; the bound only keeps a broken invariant from hanging
    for (uint32_t l = f.link[e.x]; l != e.x && guard > 0; l = f.link[e.x], --guard) e.x = l;
    for (uint32_t l = f.link[e.y]; l != e.y && guard > 0; l = f.link[e.y], --guard) e.y = l;
    if (guard <= 0) atomicOr(f.error, 4u);
    f.def_ab[i] = e;
  }
}

// counts[img][l] = colours present - FINAL tile edges - forest picks at levels <= l
__global__ void forest_lake_counts_kernel(const uint32_t* __restrict__ ndistinct, const uint32_t* __restrict__ tile_hist,
                                          const uint32_t* __restrict__ forest_hist, int n_img, uint32_t lmax,
                                          uint32_t* __restrict__ counts) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n_img) return;
  uint32_t n = ndistinct[img];
  for (uint32_t l = 0; l < 256; ++l) {
    if (l <= lmax) n -= tile_hist[(size_t)img * 256 + l] + forest_hist[(size_t)img * 256 + l];
    counts[(size_t)img * 256 + l] = (l <= lmax) ? n : 0u;
  }
}

__global__ void ctrl_accumulate_kernel(const uint32_t* src, uint32_t* dst, int as_flag) {
  const uint32_t v = *src;
  if (as_flag) { if (v) atomicOr(dst, 1u); }
  else if (v) atomicAdd(dst, v);
}
cudaError_t launch_ctrl_accumulate(const uint32_t* src, uint32_t* dst, int as_flag, cudaStream_t s) {
  ctrl_accumulate_kernel<<<1, 1, 0, s>>>(src, dst, as_flag);
  return cudaGetLastError();
}

constexpr uint32_t PACKET_HEADER_WORDS = 260;

// packet = header { edges, colours present, error bits, 0, FINAL per level [256] } + cap pairs + cap level bytes
__global__ void __launch_bounds__(256) strip_packet_kernel(ForestBuffers f, const uint32_t* __restrict__ ndistinct,
                                                           uint32_t cap, uint32_t* __restrict__ packet) {
  const uint32_t n_all = *f.n_deferred;
  const uint32_t n = n_all < cap ? n_all : cap;
  uint2* pab = reinterpret_cast<uint2*>(packet + PACKET_HEADER_WORDS);
  uint8_t* pw = reinterpret_cast<uint8_t*>(pab + cap);
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      packet[0] = n;
      packet[1] = *ndistinct;
      packet[2] = *f.error | (n_all > cap ? 2u : 0u);
      packet[3] = *f.rounds;
    }
    packet[4 + threadIdx.x] = f.tile_hist[threadIdx.x] + f.forest_hist[threadIdx.x];
  }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    pab[i] = f.def_ab[i];
    pw[i] = f.def_w[i];
  }
}
cudaError_t launch_strip_packet(ForestBuffers f, const uint32_t* ndistinct, uint32_t cap, void* packet, cudaStream_t s) {
  strip_packet_kernel<<<32, 256, 0, s>>>(f, ndistinct, cap, reinterpret_cast<uint32_t*>(packet));
  return cudaGetLastError();
}

// one CTA per packet: edges appended to f.ab[0] / f.w[0], FINAL counts summed into f.tile_hist, colours into *ndistinct
__global__ void __launch_bounds__(256) strip_unpack_kernel(const uint8_t* __restrict__ packets, uint32_t cap,
                                                           ForestBuffers f, uint32_t* out_count, uint32_t* ndistinct) {
  const size_t bytes = (size_t)PACKET_HEADER_WORDS * 4 + (size_t)cap * 9;
  const uint32_t* pk = reinterpret_cast<const uint32_t*>(packets + (size_t)blockIdx.x * bytes);
  const uint2* pab = reinterpret_cast<const uint2*>(pk + PACKET_HEADER_WORDS);
  const uint8_t* pw = reinterpret_cast<const uint8_t*>(pab + cap);
  __shared__ uint32_t s_base;
  const uint32_t n = pk[0] < cap ? pk[0] : cap;
  if (threadIdx.x == 0) {
    s_base = atomicAdd(out_count, n);
    atomicAdd(ndistinct, pk[1]);
    if (pk[2]) atomicOr(f.error, pk[2] | 8u);
  }
  const uint32_t h = pk[4 + threadIdx.x];
  if (h) atomicAdd(&f.tile_hist[threadIdx.x], h);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    f.ab[0][s_base + i] = pab[i];
    f.w[0][s_base + i] = pw[i];
  }
}
cudaError_t launch_strip_unpack(const void* packets, uint32_t n_packets, uint32_t cap, ForestBuffers f,
                                uint32_t* out_count, uint32_t* ndistinct, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out_count, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ndistinct, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  strip_unpack_kernel<<<n_packets, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(packets), cap, f, out_count, ndistinct);
  return cudaGetLastError();
}

int forest_max_grid(int device) {
  int per_sm = 0, per_sm_open = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)forest_boruvka_kernel<false>, 256, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_open, (const void*)forest_boruvka_kernel<true>, 256, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  per_sm = per_sm < per_sm_open ? per_sm : per_sm_open;
  const int cap = 4 * sms;   // rounds are latency bound: more CTAs only make the barriers dearer
  return per_sm * sms < cap ? per_sm * sms : cap;
}

cudaError_t launch_forest_init(ForestBuffers f, uint32_t ncolours, int n_img, int with_open, int sms, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(f.count, 0, 16 * sizeof(uint32_t), s);   // count[2], n_deferred, rounds, error
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(f.forest_hist, 0, sizeof(uint32_t) * 256 * (size_t)n_img, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(f.tile_hist, 0, sizeof(uint32_t) * 256 * (size_t)n_img, s);
  if (e != cudaSuccess) return e;
  if (with_open) {
    e = cudaMemsetAsync(f.open_, 0, ncolours ? ncolours : 1, s);
    if (e != cudaSuccess) return e;
  }
  if (ncolours == 0) return cudaSuccess;
  const uint32_t want = (ncolours + 255) / 256;
  const uint32_t cap = (uint32_t)sms * 16u;
  forest_init_kernel<<<want < cap ? want : cap, 256, 0, s>>>(f, ncolours);
  return cudaGetLastError();
}

cudaError_t launch_forest_mark_open(const uint32_t* lab, size_t count, uint32_t ncolours, uint8_t* open_,
                                    cudaStream_t s) {
  if (count == 0) return cudaSuccess;
  forest_mark_open_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(lab, count, ncolours, open_);
  return cudaGetLastError();
}

cudaError_t launch_forest(ForestBuffers f, const uint2* in_ab, const uint8_t* in_w, const uint32_t* in_count,
                          int skip_final, int with_open, const uint32_t* seed_off, int n_img, int grid, int sms,
                          cudaStream_t s) {
  ForestArgs a;
  a.f = f;
  a.in_ab = in_ab;
  a.in_w = in_w;
  a.in_count = in_count;
  a.skip_final = skip_final;
  a.seed_off = seed_off;
  a.n_img = n_img;
  void* args[] = {&a};
  const void* fn = with_open ? (const void*)forest_boruvka_kernel<true> : (const void*)forest_boruvka_kernel<false>;
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, s);
  if (e != cudaSuccess || !with_open) return e;
  forest_ident_kernel<<<sms * 2, 256, 0, s>>>(f);
  return cudaGetLastError();
}

cudaError_t launch_forest_lake_counts(const uint32_t* ndistinct, ForestBuffers f, int n_img, uint32_t lmax,
                                      uint32_t* counts, cudaStream_t s) {
  forest_lake_counts_kernel<<<(n_img + 63) / 64, 64, 0, s>>>(ndistinct, f.tile_hist, f.forest_hist, n_img, lmax, counts);
  return cudaGetLastError();
}

}  // namespace ws
