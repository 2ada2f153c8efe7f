// merge.cu -- K4a: basin-adjacency edges of the merging transform, contracted and reduced per tile.
//
// find_merge (lib.rs:393-445) looks from every coloured window centre at its coloured
// 4-neighbours of a different colour; make_colour_map (lib.rs:467-542) closes those pairs
// transitively at every water level.  In arrival-time terms: two adjacent coloured pixels p, q
// (at least one of them a window centre, lib.rs:411-414) with different segmenting labels a != b
// put an edge (a, b) of weight w = max(level(p), level(q)) into the basin graph G, and the lakes at
// level L are the components of the edges with w <= L, so
//     lakes(L) = colours on the canvas - (edges of a minimum spanning forest of G with w <= L).
// Every CTA takes ONE 64x32 tile (plus the pixels right of / below it) and runs Boruvka on the tile's basin
// graph in shared memory, under the total edge order (w, edge id in the tile); every component picks its
// lightest edge in every round, and a picked edge is classified by who picked it:
//   FINAL.  A basin none of whose pixels touches the tile's rim is *closed*: all its edges are in this
//      tile.  The lightest edge leaving a component made of closed basins only is the lightest edge
//      leaving it in the WHOLE graph, hence a forest edge for good (cut property).  It is only counted.
//      The set of basins joined by FINAL edges holds at most one open basin (the edge that would join two
//      sets with an open basin each is picked by an open component); its *identity* -- the open basin, or
//      the closed basin where the chain of FINAL moves ends -- stands for all its basins in what follows.
//   DEFERRED.  Every other pick.  Together with the FINAL ones they are a minimum spanning forest of the
//      tile's graph, and only those can be in the global forest (cycle property); they are emitted between
//      the identities of their basins and go through the global level-ordered union-find.
// (Round 1 used to let only closed components pick, then a second Boruvka reduced the rest: 8.9 rounds and
// 7.4 k edge looks per tile on a noise field; one Boruvka with classified picks needs 5.3 rounds and 4.8 k
// looks -- 3.2 k after the first round, which needs no liveness test -- for 15 % more DEFERRED edges.)
// On a noise field ~3/4 of all basins are closed in their tile, so the global pass -- random accesses
// over tens of millions of colours -- sees a quarter of the graph.  Final edges are kept (flagged) in the
// same list: the merge tree that per-level representatives need is built from all of them on demand.
#include "kernels.cuh"

namespace ws {

constexpr int MR_NW = TILE_W + 1;              // node grid = tile pixels + right / bottom neighbours
constexpr int MR_NH = TILE_H + 1;
constexpr int MR_NODES = MR_NW * MR_NH;        // 2145
constexpr int MR_THREADS = 256;
constexpr int MR_WARPS = MR_THREADS / 32;
constexpr int MR_PER_THREAD = (MR_NODES + MR_THREADS - 1) / MR_THREADS;  // 9
constexpr uint32_t MR_NONE = 0xFFFFFFFFu;
constexpr uint16_t MR_NOLAB = 0xFFFFu;

// Two sizes of the same kernel.  A tile can hold up to 2145 basins and 4096 edges, a real one holds a few hundred
// (noise field: 260 basins, 1600 edges): the SMALL size -- up to 1024 basins, 384 edges per warp -- needs 28 KB
// of shared memory and 40 registers, so six CTAs share an SM instead of four (the kernel is bound by the latency
// of dependent shared-memory accesses: more warps is what it wants).  A tile that does not fit is put on a list
// and redone by the FULL size, which fits everything.
struct MrSmall { static constexpr int MAXN = 1024, HASH = 2048, HSHIFT = 21, SEG = 384; };
struct MrFull  { static constexpr int MAXN = MR_NODES, HASH = 4096, HSHIFT = 20, SEG = 2 * 32 * ROWS_PER_THREAD; };

// Shared memory of one tile.  An edge is one word: id(a) | id(b) << 12 | level << 24.
template <typename Z>
struct MergeSmem {
  uint32_t label_of[Z::MAXN];    // dense id -> colour
  uint16_t parent[Z::MAXN];      // union-find over dense ids; only a root's own thread re-parents it
  uint8_t open_[(Z::MAXN + 3) & ~3];  // dense id: the basin has pixels on the tile's rim; propagated to the roots
                                 // (only ever set, so plain byte stores of 1 are race-free)
  union {
    uint32_t table[Z::HASH];     // while dense ids are handed out: 0 = free, else the label / the id
    uint32_t edge[MR_WARPS * Z::SEG];  // afterwards: 8 warp-private segments of live edges, compacted every round
  } a;
  union {
    struct { uint32_t node[MR_NODES]; } n;  // until the edge list is built: dense id | level << 16 of every node
    struct {
      uint32_t best[Z::MAXN];    // root: smallest (level << 16 | slot) offered this round; after the id went
                                 // under another component: the edge it went along (an edge word)
      uint16_t comp[Z::MAXN];    // root of every id at the start of the round
      uint16_t link[Z::MAXN];    // FINAL moves only: the basin at the far end of the edge (else the id itself);
                                 // following the links gives the identity of a contracted basin; bit 15: FINAL
    } r;
  } b;
  uint32_t nlab, nout, gpos, first, overflow;
};

// The right / down edges of a pixel of a tile on the image's border or of a row strip (one copy of the code for
// the eight pixels of a thread: these tiles are few, the kernel is short of instruction cache).
__device__ __noinline__ uint2 mr_border_edges(int gr, int gc, int rows, int cols, int row_offset, int global_rows,
                                              int halos, uint32_t na, uint32_t nr, uint32_t nd) {
  const uint32_t a = na & 0xFFFFu, br = nr & 0xFFFFu, bd = nd & 0xFFFFu;
  uint2 e = make_uint2(MR_NONE, MR_NONE);  // (right, down)
  // window centres of the WHOLE field (lib.rs:220, 411-414), in local coordinates: ImageDims::is_centre
  auto centre = [&](int r, int c) {
    const int g = r + row_offset;
    return g >= 1 && g <= global_rows - 2 && c >= 1 && c <= cols - 2;
  };
  // An edge belongs to the strip that owns its upper / left pixel; a halo row's own edges are the
  // neighbouring strip's.  Plain plans own every row.  (halos: bit 0 = halo_top, bit 1 = halo_bottom)
  if (a != MR_NOLAB && gr < rows && !((halos & 1) && gr == 0) && !((halos & 2) && gr == rows - 1)) {
    const bool pin = centre(gr, gc);
    if (br != MR_NOLAB && br != a && (pin || centre(gr, gc + 1)))
      e.x = a | (br << 12) | (max(na >> 16, nr >> 16) << 24);
    if (bd != MR_NOLAB && bd != a && (pin || centre(gr + 1, gc)))
      e.y = a | (bd << 12) | (max(na >> 16, nd >> 16) << 24);
  }
  return e;
}

// One Boruvka over the tile's basin graph; every component picks its lightest edge in every round.  A pick
// made by a component that holds closed basins only is FINAL (cut property: all edges of a closed basin lie
// in this tile), every other pick is DEFERRED.  `contract` = 0 treats every basin as open (no FINAL edges).
// Returns false when the tile does not fit this size (nothing has been emitted then).
template <typename Z>
__device__ __forceinline__ bool merge_tile(MergeSmem<Z>& sm, const int img, const int ty, const int tx,
                                           const uint32_t* __restrict__ lab,
                                           const uint8_t* __restrict__ lvl, const ImageDims& d,
                                           const uint32_t* __restrict__ seed_off, const int contract,
                                           uint2* __restrict__ red_ab, uint8_t* __restrict__ red_w,
                                           uint32_t* __restrict__ red_count) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = ty * TILE_H, c0 = tx * TILE_W;
  const size_t base = (size_t)img * d.px_per_img();

  // (a) labels and levels of the tile and of its right / bottom neighbours: all loads first, so that they are in
  // flight together (every load used at once cost a round trip each: 11 % of the kernel's stall samples)
  uint32_t L[MR_PER_THREAD], V[MR_PER_THREAD];
  if (tid == 0) { sm.nlab = 0; sm.nout = 0; sm.first = 0; sm.overflow = 0; }
  // node i = tid + 256 k sits at (row, column) = (i / 65, i % 65): 256 = 3 * 65 + 61, so no division per node
  static_assert(MR_THREADS == 3 * MR_NW + 61, "incremental node coordinates");
  const int nr0 = tid / MR_NW, nc0 = tid - nr0 * MR_NW;
  // the node grid lies inside the image (all tiles but those of the last tile row / column): no bounds tests,
  // offsets by addition
  const bool whole = r0 + MR_NH <= d.rows && c0 + MR_NW <= d.cols;
  if (whole) {
    const size_t p0 = base + (size_t)(r0 + nr0) * d.cols + c0 + nc0;
    const uint32_t* lp = lab + p0;
    const uint8_t* vp = lvl + p0;
    // node + 256 = 3 rows down and 61 right, or 4 down and 4 left (unsigned: up to 36 rows of a very wide image)
    const uint32_t step3 = 3u * (uint32_t)d.cols + 61u, back = (uint32_t)d.cols - (uint32_t)MR_NW;
    int c = nc0;
    uint32_t off = 0;
#pragma unroll
    for (int k = 0; k < MR_PER_THREAD; ++k) {
      const int i = tid + k * MR_THREADS;
      L[k] = 0;
      V[k] = 255;
      if (i < MR_NODES) {
        L[k] = __ldg(lp + off);
        V[k] = __ldg(vp + off);
      }
      off += step3;
      c += 61;
      if (c >= MR_NW) { c -= MR_NW; off += back; }
    }
  } else {
    int r = nr0, c = nc0;
#pragma unroll
    for (int k = 0; k < MR_PER_THREAD; ++k) {
      const int i = tid + k * MR_THREADS;
      L[k] = 0;
      V[k] = 255;
      if (i < MR_NODES) {
        const int gr = r0 + r, gc = c0 + c;
        if (gr < d.rows && gc < d.cols) {
          const size_t p = base + (size_t)gr * d.cols + gc;
          L[k] = __ldg(lab + p);
          V[k] = __ldg(lvl + p);
        }
      }
      r += 3;
      c += 61;
      if (c >= MR_NW) { c -= MR_NW; ++r; }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < MR_PER_THREAD; ++k) {
    const int i = tid + k * MR_THREADS;
    L[k] &= LAB_MASK;
    if (i < MR_NODES) {
      sm.b.n.node[i] = V[k] << 16;
      if (L[k] != 0u && sm.first == 0u) sm.first = L[k];  // any coloured label (benign race)
    }
  }
  for (int i = tid; i < Z::MAXN; i += MR_THREADS) {
    sm.open_[i] = 0;
    sm.parent[i] = (uint16_t)i;
  }
  for (int i = tid; i < Z::HASH; i += MR_THREADS) sm.a.table[i] = 0u;
  __syncthreads();
  {  // a tile inside one basin (most tiles of a smooth field) has no edge at all
    const uint32_t f = sm.first;
    bool differs = false;
#pragma unroll
    for (int k = 0; k < MR_PER_THREAD; ++k) differs |= (L[k] != 0u && L[k] != f);
    if (!__syncthreads_or(differs)) return true;
  }

  // (b) dense ids, one per distinct label: insert the labels into an open-addressing table, number the
  // occupied slots, then replace every slot's label by its id.  Consecutive lanes hold consecutive nodes of
  // a row, and a basin is a few pixels wide: only the first lane of a run of equal labels probes the table,
  // the others take its slot by shuffle.
  uint16_t slot[MR_PER_THREAD];
#pragma unroll
  for (int k = 0; k < MR_PER_THREAD; ++k) {
    const uint32_t l = L[k];
    const uint32_t prev = __shfl_up_sync(0xffffffffu, l, 1);
    const bool lead = lane == 0 || prev != l;
    const uint32_t leaders = __ballot_sync(0xffffffffu, lead);
    uint32_t h = (l * 2654435761u) >> Z::HSHIFT;
    if (lead && l != 0u) {
      int probes = 0;
#pragma unroll 1   // (unrolled four times in each of the nine copies it was a quarter of the kernel's code: the
                   //  kernel's top stall reason was instruction fetch)
      for (;; ++probes) {
        if (probes == Z::HASH) {   // more distinct labels than slots (small size only)
          sm.overflow = 1;
          break;
        }
        uint32_t cur = ((volatile uint32_t*)sm.a.table)[h];
        if (cur == 0u) cur = atomicCAS(&sm.a.table[h], 0u, l);
        if (cur == 0u || cur == l) break;
        h = (h + 1u) & (Z::HASH - 1);
      }
    }
    const int src = 31 - __clz((int)(leaders & (0xFFFFFFFFu >> (31 - lane))));  // my run's first lane
    slot[k] = (uint16_t)__shfl_sync(0xffffffffu, h, src);
  }
  __syncthreads();
  {
    uint32_t key[Z::HASH / MR_THREADS];
#pragma unroll
    for (int k = 0; k < Z::HASH / MR_THREADS; ++k) key[k] = sm.a.table[tid + k * MR_THREADS];
    // ids in blocks: one shared atomic per warp
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < Z::HASH / MR_THREADS; ++k) mine += (key[k] != 0u);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    uint32_t wbase = 0;
    if (lane == 31 && incl) wbase = atomicAdd(&sm.nlab, incl);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    uint32_t id = wbase + incl - mine;
#pragma unroll
    for (int k = 0; k < Z::HASH / MR_THREADS; ++k) {
      if (key[k] == 0u) continue;
      if (id < (uint32_t)Z::MAXN) sm.label_of[id] = key[k];
      sm.a.table[tid + k * MR_THREADS] = id;  // (only this thread touches the slot in this phase)
      ++id;
    }
  }
  __syncthreads();
  const int nlab = (int)sm.nlab;
  if (nlab > Z::MAXN || sm.overflow) return false;   // (uniform: both were written before the barrier)
  int nr = nr0, nc = nc0;
  const bool all_open = !contract, strip = d.halo_top || d.halo_bottom;
  const int rim_r = r0 > 0 ? 0 : -1, rim_c = c0 > 0 ? 0 : -1;   // first row / column, when a tile lies beyond it
#pragma unroll
  for (int k = 0; k < MR_PER_THREAD; ++k, nr += 3, nc += 61) {
    if (nc >= MR_NW) { nc -= MR_NW; ++nr; }   // (row, column) of node tid + 256 k
    const int i = tid + k * MR_THREADS;
    if (i >= MR_NODES) continue;
    uint32_t id = MR_NOLAB;
    if (L[k] != 0u) {
      id = sm.a.table[slot[k]];
      const int r = nr, c = nc;
      // rim: pixels whose up / left neighbour lies outside the tile, and the neighbour row / column itself;
      // in a row strip also the halo rows (the neighbouring strip's pixels) and the first owned row below a
      // halo (its upward edges belong to the neighbouring strip)
      bool rim = all_open || r == TILE_H || c == TILE_W || r == rim_r || c == rim_c;
      if (strip) {
        const int gr = r0 + r;
        rim = rim || (d.halo_top && gr <= 1) || (d.halo_bottom && gr == d.rows - 1);
      }
      if (rim) sm.open_[id] = 1;
    }
    sm.b.n.node[i] |= id;  // (MR_NOLAB = 0xFFFF for uncoloured nodes)
  }
  __syncthreads();  // last use of the table: its memory becomes the edge list

  // (c) the edges of my 8 pixels, compacted into my warp's segment of the edge list.  A tile whose nodes are all
  // window centres of an unstriped image (every tile but those on the image's border) needs none of the
  // per-pixel geometry tests.
  const int lc = tid % TILE_W, g = tid / TILE_W;
  uint32_t* seg = sm.a.edge + warp * Z::SEG;
  uint32_t cnt = 0;  // live edges in the segment (uniform across the warp)
  const bool interior = !d.halo_top && !d.halo_bottom && d.row_offset == 0 && r0 >= 1 && c0 >= 1 &&
                        r0 + TILE_H <= d.rows - 2 && c0 + TILE_W <= d.cols - 2;
#pragma unroll
  for (int i = 0; i < ROWS_PER_THREAD; ++i) {
    const int r = g * ROWS_PER_THREAD + i;
    const int n = r * MR_NW + lc;
    const uint32_t na = sm.b.n.node[n], nr = sm.b.n.node[n + 1], nd = sm.b.n.node[n + MR_NW];
    const uint32_t a = na & 0xFFFFu, br = nr & 0xFFFFu, bd = nd & 0xFFFFu;
    uint32_t er = MR_NONE, ed = MR_NONE;
    if (interior) {
      if (a != MR_NOLAB) {
        if (br != MR_NOLAB && br != a) er = a | (br << 12) | (max(na >> 16, nr >> 16) << 24);
        if (bd != MR_NOLAB && bd != a) ed = a | (bd << 12) | (max(na >> 16, nd >> 16) << 24);
      }
    } else {
      const uint2 e = mr_border_edges(r0 + r, c0 + lc, d.rows, d.cols, d.row_offset, d.global_rows,
                                      (d.halo_top ? 1 : 0) | (d.halo_bottom ? 2 : 0), na, nr, nd);
      er = e.x;
      ed = e.y;
    }
    // (no write of this loop can hit the table's last readers: they are behind the barrier above)
    uint32_t m = __ballot_sync(0xffffffffu, er != MR_NONE);
    uint32_t pos = cnt + __popc(m & ((1u << lane) - 1u));
    if (er != MR_NONE && pos < (uint32_t)Z::SEG) seg[pos] = er;
    cnt += __popc(m);
    m = __ballot_sync(0xffffffffu, ed != MR_NONE);
    pos = cnt + __popc(m & ((1u << lane) - 1u));
    if (ed != MR_NONE && pos < (uint32_t)Z::SEG) seg[pos] = ed;
    cnt += __popc(m);
  }
  if (cnt > (uint32_t)Z::SEG) sm.overflow = 1;   // more edges than this size holds (small size only)
  if (!__syncthreads_or(cnt != 0u)) return true;  // no edge between different basins in this tile
  if (sm.overflow) return false;                  // (written before the barrier: uniform)
  // (node[] is dead from here on: its memory becomes best / comp / link)
#pragma unroll 1
  for (int i = tid; i < nlab; i += MR_THREADS) {
    sm.b.r.best[i] = MR_NONE;
    sm.b.r.comp[i] = (uint16_t)i;
    sm.b.r.link[i] = (uint16_t)i;
  }
  __syncthreads();

  // (d) Boruvka.  Keys (level << 16 | slot in the edge list) are distinct within a round, so the picks of a
  // round form a forest apart from mutual picks of one edge; of such a pair exactly one side moves: the closed
  // one if only one is closed (so that an open basin never goes along a FINAL edge), else the larger root.
  // (The order among equal levels may differ from round to round: every round is a valid Boruvka step on the
  // graph contracted so far.)
  // Round 1: every basin is its own component and every edge is alive -- offers only.
  for (uint32_t j = lane; j < cnt; j += 32) {
    const uint32_t e = seg[j];
    const uint32_t key = ((e >> 24) << 16) | (uint32_t)(warp * Z::SEG + j);
    atomicMin(&sm.b.r.best[e & 0xFFFu], key);
    atomicMin(&sm.b.r.best[(e >> 12) & 0xFFFu], key);
  }
  constexpr uint32_t FIN = 0x8000u, IDM = 0x7FFFu;
  for (;;) {
    __syncthreads();
#ifdef WS_MERGE_STATS  // rounds and edge looks (scripts/merge_stats.py)
    if (tid == 0) atomicAdd(&red_count[4], 1u);
    if (lane == 0) atomicAdd(&red_count[6], cnt);
#endif
    // every picking root goes under the component at the other end of its edge
#pragma unroll 1
    for (int i = tid; i < nlab; i += MR_THREADS) {
      if (sm.b.r.comp[i] != (uint16_t)i) continue;
      const uint32_t key = sm.b.r.best[i];
      if (key == MR_NONE) continue;
      const uint32_t e = sm.a.edge[key & 0xFFFFu];
      const uint32_t ia = e & 0xFFFu, ib = (e >> 12) & 0xFFFu;
      const uint32_t cu = sm.b.r.comp[ia], cv = sm.b.r.comp[ib];
      const uint32_t other = cu == (uint32_t)i ? cv : cu;
      const bool mutual = sm.b.r.best[other] == key;
      const bool ci = !sm.open_[i], co = !sm.open_[other];  // closed components (open_ of a root is current)
      if (mutual && (ci == co ? (uint32_t)i < other : !ci)) continue;  // the other one moves
      sm.parent[i] = (uint16_t)other;
      if (ci || (mutual && co))  // FINAL; a closed mover takes the identity of the basin at the far end
        sm.b.r.link[i] = (uint16_t)(FIN | (ci ? (cu == (uint32_t)i ? ib : ia) : (uint32_t)i));
    }
    __syncthreads();
    // roots (reads racing with the flattening stores still see an ancestor).  An id that was a root and has
    // just gone under another component keeps the edge it went along (the list is not compacted in between).
#pragma unroll 1
    for (int i = tid; i < nlab; i += MR_THREADS) {
      uint32_t x = (uint32_t)i;
      for (uint32_t p = sm.parent[x]; p != x; p = sm.parent[x]) x = p;
      if (sm.b.r.comp[i] == (uint16_t)i) {
        if (x != (uint32_t)i) sm.b.r.best[i] = sm.a.edge[sm.b.r.best[i] & 0xFFFFu];
        else sm.b.r.best[i] = MR_NONE;
      }
      sm.parent[i] = (uint16_t)x;
      sm.b.r.comp[i] = (uint16_t)x;
      if (x != (uint32_t)i && sm.open_[i]) sm.open_[x] = 1;
    }
    __syncthreads();
    // one pass over my warp's live edges: drop those inside one component, keep the rest compacted,
    // offer each to the components at its ends
    bool any = false;
    uint32_t kept = 0;
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
      const uint32_t j = j0 + lane;
      uint32_t e = 0, cu = 0, cv = 0;
      if (j < cnt) {
        e = seg[j];
        cu = sm.b.r.comp[e & 0xFFFu];
        cv = sm.b.r.comp[(e >> 12) & 0xFFFu];
      }
      const bool alive = cu != cv;
      const uint32_t m = __ballot_sync(0xffffffffu, alive);  // (also orders this step's reads before its writes)
      if (alive) {
        const uint32_t pos = kept + __popc(m & ((1u << lane) - 1u));
        seg[pos] = e;
        const uint32_t key = ((e >> 24) << 16) | (uint32_t)(warp * Z::SEG + pos);
        atomicMin(&sm.b.r.best[cu], key);
        atomicMin(&sm.b.r.best[cv], key);
        any = true;
      }
      kept += __popc(m);
    }
    cnt = kept;
    if (!__syncthreads_or(any)) break;
  }

  // (e) every id that is not a root went under another component along exactly one edge: append those
  // edges to the global list as global colour ids -- FINAL edges (bit 31 of .y) between the two basins
  // themselves, DEFERRED ones between the identities of the two basins (the end of their chains of FINAL moves)
  uint32_t mine = 0;
  if (tid == 0) sm.first = 0;  // from here on: cursor into this tile's part of the global list
#pragma unroll 1
  for (int i = tid; i < nlab; i += MR_THREADS) mine += (sm.parent[i] != (uint16_t)i);
  if (mine) atomicAdd(&sm.nout, mine);
  __syncthreads();
  const uint32_t nout = sm.nout;
#ifdef WS_MERGE_STATS
  if (tid == 0) { atomicAdd(&red_count[8], 1u); atomicAdd(&red_count[9], (uint32_t)nlab); atomicAdd(&red_count[10], nout); }
#endif
  if (nout == 0u) return true;
  if (tid == 0) sm.gpos = atomicAdd(red_count, nout);
  __syncthreads();
  const uint32_t gbase = __ldg(seed_off + img) - 1u;  // global colour id = seed_off[img] + colour - 1
  uint32_t at = mine ? sm.gpos + atomicAdd(&sm.first, mine) : 0u;
#pragma unroll 1
  for (int i = tid; i < nlab; i += MR_THREADS) {
    if (sm.parent[i] == (uint16_t)i) continue;
    const uint32_t e = sm.b.r.best[i];
    const uint32_t fin = (sm.b.r.link[i] & FIN) ? 1u : 0u;
    uint32_t ia = e & 0xFFFu, ib = (e >> 12) & 0xFFFu;
    if (!fin) {  // (read-only walks: no link changes any more; chains of FINAL moves are acyclic -- the bound
                 // only keeps a broken invariant from hanging the device)
      int guard = nlab;
      for (uint32_t l = sm.b.r.link[ia] & IDM; l != ia && guard > 0; l = sm.b.r.link[ia] & IDM, --guard) ia = l;
      for (uint32_t l = sm.b.r.link[ib] & IDM; l != ib && guard > 0; l = sm.b.r.link[ib] & IDM, --guard) ib = l;
    }
    red_ab[at] = make_uint2(gbase + sm.label_of[ia], (gbase + sm.label_of[ib]) | (fin << 31));
    red_w[at] = (uint8_t)(e >> 24);
    ++at;
  }
  return true;
}

// every tile with the small size; the ones that do not fit go on ovf_list (ovf_count = red_count[12])
__global__ void __launch_bounds__(MR_THREADS, 6) merge_reduce_kernel(const uint32_t* __restrict__ lab,
                                                                     const uint8_t* __restrict__ lvl, ImageDims d,
                                                                     const uint32_t* __restrict__ seed_off, int contract,
                                                                     uint2* __restrict__ red_ab, uint8_t* __restrict__ red_w,
                                                                     uint32_t* __restrict__ red_count,
                                                                     uint32_t* __restrict__ ovf_list) {
  __shared__ MergeSmem<MrSmall> sm;
  // grid = (tiles_x, tiles_y, slices) whenever that fits a grid: no divisions per thread
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
  if (!merge_tile<MrSmall>(sm, img, ty, tx, lab, lvl, d, seed_off, contract, red_ab, red_w, red_count) &&
      threadIdx.x == 0)
    ovf_list[atomicAdd(&red_count[12], 1u)] = (uint32_t)((img * d.tiles_y + ty) * d.tiles_x + tx);
}

// the listed tiles with the full size (a persistent grid: the list is usually empty)
__global__ void __launch_bounds__(MR_THREADS) merge_reduce_full_kernel(const uint32_t* __restrict__ lab,
                                                                       const uint8_t* __restrict__ lvl, ImageDims d,
                                                                       const uint32_t* __restrict__ seed_off, int contract,
                                                                       uint2* __restrict__ red_ab, uint8_t* __restrict__ red_w,
                                                                       uint32_t* __restrict__ red_count,
                                                                       const uint32_t* __restrict__ ovf_list) {
  __shared__ MergeSmem<MrFull> sm;
  const uint32_t n = red_count[12];
  for (uint32_t k = blockIdx.x; k < n; k += gridDim.x) {
    const int tile = (int)ovf_list[k], tpi = d.tiles_per_img();
    const int img = tile / tpi, trem = tile - img * tpi;
    merge_tile<MrFull>(sm, img, trem / d.tiles_x, trem % d.tiles_x, lab, lvl, d, seed_off, contract, red_ab, red_w, red_count);
    __syncthreads();  // the next tile reuses the shared memory
  }
}

size_t merge_reduce_capacity(const ImageDims& d) { return (size_t)d.tiles_total() * (MR_NODES - 1); }

cudaError_t launch_merge_reduce(const uint32_t* lab, const uint8_t* lvl, ImageDims d, const uint32_t* seed_off,
                                int contract, uint2* red_ab, uint8_t* red_w, uint32_t* red_count, uint32_t* ovf_list,
                                cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(red_count, 0, 16 * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  const dim3 grid = (d.tiles_y <= 65535 && d.n_img <= 65535) ? dim3(d.tiles_x, d.tiles_y, d.n_img) : dim3(d.tiles_total());
  merge_reduce_kernel<<<grid, MR_THREADS, 0, s>>>(lab, lvl, d, seed_off, contract, red_ab, red_w, red_count, ovf_list);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int want = d.tiles_total(), cap = num_sms() * 4;
  merge_reduce_full_kernel<<<want < cap ? want : cap, MR_THREADS, 0, s>>>(lab, lvl, d, seed_off, contract, red_ab, red_w,
                                                                             red_count, ovf_list);
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
  red_hist_kernel<<<num_sms() * 8, 256, 0, s>>>(red_ab, red_w, red_count, all, seed_off, n_img, level_hist, fin_hist);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  edge_scan_kernel<<<1, 256, 0, s>>>(level_hist, level_cursor);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  red_scatter_kernel<<<num_sms() * 8, 256, 0, s>>>(red_ab, red_w, red_count, all, level_cursor, edges);
  return cudaGetLastError();
}

}  // namespace ws
