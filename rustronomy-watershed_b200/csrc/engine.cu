// engine.cu -- host side of the engine: context, workspaces, the kernel
// pipeline, and the extern "C" boundary declared in include/ws_b200.h.
//
// Pipeline of one run (all on the context's stream, no host round trip except
// where noted):
//   fill_state -> seed_init -> flood (persistent, asynchronous worklist)
//   -> label_tile -> rim_jump (persistent, cooperative) -> label_finish      [labels + levels]
//   merging only: merge_reduce (per-tile contraction + spanning forest in shared memory) -> red_hist ->
//   edge_scan -> red_scatter -> uf_init -> union_levels (persistent, cooperative) -> lake_counts;
//   the merge tree for per-level representatives is built on demand (plan_build_tree)
#include "../../include/ws_b200.h"
#include "hostpipe.h"
#include "kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <random>
#include <string>
#include <thread>
#include <vector>

using namespace ws;

// ---------------------------------------------------------------------------
// context / plan objects
// ---------------------------------------------------------------------------

struct ws_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  int flood_grid = 0, jump_grid = 0, union_grid = 0, forest_grid = 0, sms = 0;
  std::string err;
  uint64_t tie_seed = 0;      // key of the WS_TIE_RANDOM generator
  ws_plan* cached = nullptr;  // workspace reused by the host-level entry points
  // host-level scratch (device)
  uint8_t* d_img = nullptr;   size_t d_img_cap = 0;
  uint8_t* d_raw = nullptr;   size_t d_raw_cap = 0;
  uint32_t* d_seeds = nullptr; size_t d_seeds_cap = 0;  // [cap][2]
  uint64_t* d_seeds64 = nullptr; size_t d_seeds64_cap = 0;  // staging of the caller's (usize, usize) pairs
  void* d_pp_scratch = nullptr;  // pre_processor partial minima / maxima
  uint32_t* d_seed_off = nullptr; size_t d_seed_off_cap = 0;
  uint64_t* d_out[2] = {nullptr, nullptr}; size_t d_out_cap[2] = {0, 0};
  uint64_t* h_pin[2] = {nullptr, nullptr}; size_t h_pin_cap[2] = {0, 0};
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  // pageable caller memory: worker threads + a ring of page-locked slots (hostpipe.h), made on first use
  HostPool pool;
  int host_threads = 0;            // 0: not started yet
  int host_threads_wanted = 0;     // 0: default (min(16, CPUs this process may use), WS_HOST_THREADS)
  bool pinned_host_widen = false;  // page-locked label outputs: u32 over the link + widening by the workers
  uint8_t* ring = nullptr;
  bool ring_used[8] = {false, false, false, false, false, false, false, false};  // ring_ev[i] has been recorded
  size_t ring_pos = 0;             // slots are handed out round robin ACROSS calls
  cudaEvent_t ring_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct ws_plan {
  ws_ctx* ctx = nullptr;
  ImageDims d{};
  FloodBuffers fb{};
  MergeBuffers mb{};
  ForestBuffers fo{};
  size_t fo_col_cap = 0, fo_edge_cap = 0, fo_def_cap = 0;
  size_t uf_cap = 0, edges_cap = 0, rep_cap = 0;
  uint32_t* rep = nullptr;
  int rep_level = -1;
  uint32_t* T_dense = nullptr;       // lazily made dense copy of the arrival times (diagnostic accessor)
  uint32_t* lvl_hist = nullptr;      // [n_img][256] pixels coloured at each level
  uint32_t* d_strip_off = nullptr;   // [2] seed offsets {0, nseeds} of a strip run
  uint32_t colour_base = 0;          // strips: colour of local seed i = colour_base + i + 1
  int bucket_shift = 2;              // priority granularity of the flood's worklist in the last run
  int check_ovf = 0;                 // hop counters can overflow: the (whole) field has more than 2^24 pixels
  alignas(64) unsigned char tmaps[FLOOD_TENSOR_MAP_BYTES];  // tensor maps of fb.T / fb.pix for the flood's tile copies
  uint2* union_edges = nullptr;      // ws_plan_union_edges: the gathered edges bucketed by level
  size_t union_edges_cap = 0;
  uint32_t* chunk_counts = nullptr;  // minima scratch
  uint32_t* d_total = nullptr;
  uint32_t* h_ctrl = nullptr;        // pinned mirror of fb.ctrl + scalars
  // last run
  const uint32_t* seeds = nullptr;
  const uint32_t* seed_off = nullptr;
  std::vector<uint32_t> h_seed_off;
  size_t nseeds = 0;
  ws_config cfg{};
  bool ran = false, merged = false;
  bool tree_built = false;           // hook_to / hook_lvl hold the full merge tree of the last merging run
  uint64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // phase boundaries of the last run
  // kernel boundaries of the last run: fill_state | seed_init | flood | label_tile | rim_jump | label_finish |
  // merge_reduce | forest_init | forest rounds (or the level-ordered union) | lake counts |
  cudaEvent_t kev[WS_KERNEL_SLOTS + 1] = {};
  bool kev_valid[WS_KERNEL_SLOTS] = {};
};

namespace {

ws_status fail(ws_ctx* ctx, ws_status s, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return s;
}

ws_status cuda_fail(ws_ctx* ctx, cudaError_t e, const char* what) {
  cudaGetLastError();  // clear the sticky-free error state
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  return fail(ctx, e == cudaErrorMemoryAllocation ? WS_ERR_OOM : WS_ERR_CUDA, m);
}

#define WS_CUDA(ctx, expr)                                        \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return cuda_fail((ctx), _e, #expr);    \
  } while (0)

#define WS_TRY(expr)                      \
  do {                                    \
    ws_status _s = (expr);                \
    if (_s != WS_OK) return _s;           \
  } while (0)

template <typename T>
ws_status grow(ws_ctx* ctx, T*& p, size_t& cap, size_t need) {
  if (need <= cap && p) return WS_OK;
  if (p) WS_CUDA(ctx, cudaFree(p));
  p = nullptr;
  cap = 0;
  const size_t n = std::max<size_t>(need, 1);
  WS_CUDA(ctx, cudaMalloc((void**)&p, n * sizeof(T)));
  cap = n;
  return WS_OK;
}

ImageDims make_dims(size_t n_img, size_t rows, size_t cols) {
  ImageDims d;
  d.n_img = (int)n_img;
  d.rows = (int)rows;
  d.cols = (int)cols;
  d.tiles_x = (int)((cols + TILE_W - 1) / TILE_W);
  d.tiles_y = (int)((rows + TILE_H - 1) / TILE_H);
  d.row_offset = 0;
  d.global_rows = (int)rows;
  d.halo_top = d.halo_bottom = 0;
  return d;
}

ws_status check_cfg(ws_ctx* ctx, const ws_config* cfg) {
  if (!cfg) return fail(ctx, WS_ERR_INVALID_ARG, "cfg is NULL");
  ws_status s = ws_config_validate(cfg);
  if (s != WS_OK) return fail(ctx, s, ws_status_str(s));
  return WS_OK;
}

}  // namespace

// ---------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------

extern "C" const char* ws_status_str(ws_status s) {
  switch (s) {
    case WS_OK: return "ok";
    case WS_ERR_INVALID_ARG: return "invalid argument";
    case WS_ERR_MAX_TOO_HIGH: return "maximum water level higher than the maximum allowed value 254";
    case WS_ERR_MAX_TOO_LOW: return "maximum water level lower than the minimum allowed value 1";
    case WS_ERR_SEED_OOB: return "seed outside the (padded) image";
    case WS_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case WS_ERR_CUDA: return "CUDA error";
    case WS_ERR_OOM: return "out of device memory";
    case WS_ERR_TOO_LARGE: return "image or seed list too large";
    case WS_ERR_HOP_OVERFLOW: return "flood path longer than 2^24-2 steps inside one level";
    case WS_ERR_INTERNAL: return "internal error";
  }
  return "unknown status";
}

extern "C" int ws_abi_version(void) { return WS_ABI_VERSION; }
extern "C" void ws_free(void* p) { free(p); }
extern "C" const char* ws_last_error(const ws_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }

extern "C" ws_status ws_config_validate(const ws_config* cfg) {
  if (!cfg) return WS_ERR_INVALID_ARG;
  if (cfg->kind != WS_SEGMENTING && cfg->kind != WS_MERGING) return WS_ERR_INVALID_ARG;
  if (cfg->max_water_level > WS_NORMAL_MAX) return WS_ERR_MAX_TOO_HIGH;   // lib.rs:1000 / 1026
  if (cfg->max_water_level <= WS_ALWAYS_FILL) return WS_ERR_MAX_TOO_LOW;  // lib.rs:1002 / 1028
  if (cfg->tie_break != WS_TIE_FIRST && cfg->tie_break != WS_TIE_RANDOM) return WS_ERR_INVALID_ARG;
  return WS_OK;
}

extern "C" ws_status ws_output_shape(const ws_config* cfg, size_t rows, size_t cols, size_t* out_rows,
                                     size_t* out_cols) {
  if (!cfg || !out_rows || !out_cols) return WS_ERR_INVALID_ARG;
  const size_t pad = cfg->edge_correction ? 2 : 0;  // lib.rs:1330-1336
  *out_rows = rows + pad;
  *out_cols = cols + pad;
  return WS_OK;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------

extern "C" ws_status ws_ctx_create(int device, ws_ctx** out) {
  if (!out) return WS_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return WS_ERR_NO_DEVICE;
  }
  if (cudaSetDevice(device) != cudaSuccess) return WS_ERR_NO_DEVICE;
  int coop = 0, major = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (!coop || major < 10) return WS_ERR_NO_DEVICE;  // sm_100a code only
  ws_ctx* c = new (std::nothrow) ws_ctx();
  if (!c) return WS_ERR_INTERNAL;
  c->device = device;
  {  // like thread_rng(): seeded from the OS, different for every context unless the caller sets it
    std::random_device rd;
    c->tie_seed = ((uint64_t)rd() << 32) ^ (uint64_t)rd() ^ (uint64_t)(uintptr_t)c;
  }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return WS_ERR_CUDA;
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&c->ev_ready[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming);
  }
  if (const char* e = getenv("WS_PINNED_HOST_WIDEN")) c->pinned_host_widen = atoi(e) != 0;
  c->flood_grid = flood_max_grid(device);
  c->jump_grid = jump_max_grid(device);
  c->union_grid = union_max_grid(device);
  c->forest_grid = forest_max_grid(device);
  cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device);
  set_num_sms(c->sms);
  if (c->flood_grid <= 0 || c->jump_grid <= 0 || c->union_grid <= 0 || c->forest_grid <= 0) {
    cudaGetLastError();
    ws_ctx_destroy(c);
    return WS_ERR_CUDA;  // e.g. the fatbin holds no code for this GPU
  }
  *out = c;
  return WS_OK;
}

extern "C" void ws_ctx_destroy(ws_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->cached) ws_plan_destroy(c->cached);
  cudaFree(c->d_img);
  cudaFree(c->d_raw);
  cudaFree(c->d_seeds);
  cudaFree(c->d_seeds64);
  cudaFree(c->d_pp_scratch);
  cudaFree(c->d_seed_off);
  c->pool.stop();
  if (c->ring) cudaFreeHost(c->ring);
  for (auto& ev : c->ring_ev)
    if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) {
    cudaFree(c->d_out[i]);
    if (c->h_pin[i]) cudaFreeHost(c->h_pin[i]);
    if (c->ev_ready[i]) cudaEventDestroy(c->ev_ready[i]);
    if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
}

extern "C" ws_status ws_ctx_set_tie_seed(ws_ctx* ctx, uint64_t seed) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  ctx->tie_seed = seed;
  return WS_OK;
}

extern "C" ws_status ws_ctx_set_host_threads(ws_ctx* ctx, int nthreads) {
  if (!ctx || nthreads < 0 || nthreads > 64) return WS_ERR_INVALID_ARG;
  ctx->host_threads_wanted = nthreads;
  if (ctx->host_threads) {  // restart with the new size on the next staged copy
    ctx->pool.stop();
    ctx->host_threads = 0;
  }
  return WS_OK;
}

extern "C" ws_status ws_ctx_set_option(ws_ctx* ctx, ws_option opt, int value) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  switch (opt) {
    case WS_OPT_PINNED_HOST_WIDEN: ctx->pinned_host_widen = value != 0; return WS_OK;
  }
  return fail(ctx, WS_ERR_INVALID_ARG, "unknown option");
}

extern "C" void* ws_ctx_stream(ws_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" ws_status ws_ctx_synchronize(ws_ctx* ctx) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return WS_OK;
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------

extern "C" ws_status ws_plan_create(ws_ctx* ctx, size_t n_img, size_t rows, size_t cols, ws_plan** out) {
  if (!ctx || !out) return WS_ERR_INVALID_ARG;
  *out = nullptr;
  if (n_img == 0 || rows == 0 || cols == 0) return fail(ctx, WS_ERR_INVALID_ARG, "empty image");
  // pixel indices live in 31 bits of a label word
  if (rows > 0x7fffffffull || cols > 0x7fffffffull || n_img > 0x7fffffffull ||
      (double)n_img * (double)rows * (double)cols >= 2147483648.0)
    return fail(ctx, WS_ERR_TOO_LARGE, "more than 2^31 - 1 pixels in one plan");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  ws_plan* p = new (std::nothrow) ws_plan();
  if (!p) return WS_ERR_INTERNAL;
  p->ctx = ctx;
  p->d = make_dims(n_img, rows, cols);
  const size_t npx = p->d.px_total();
  const size_t ntiles = (size_t)p->d.tiles_total();
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** ptr, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(ptr, std::max<size_t>(bytes, 16));
  };
  alloc((void**)&p->fb.T, p->d.t_plane() * n_img * 4);
  alloc((void**)&p->fb.pix, p->d.pix_plane() * n_img);
  alloc((void**)&p->fb.lab, npx * 4);
  alloc((void**)&p->fb.lvl, npx);
  alloc((void**)&p->fb.rim, rim_words(p->d) * 4);
  p->fb.qcap = (uint32_t)ntiles + FLOOD_QSLACK;
  alloc((void**)&p->fb.qslots, (size_t)FLOOD_BUCKETS * p->fb.qcap * 4);
  alloc((void**)&p->fb.qmask, ntiles * 8);
  alloc((void**)&p->fb.ctrl, FC_WORDS * 4);
  alloc((void**)&p->fb.row_start, (n_img * rows + 1) * 4);
  alloc((void**)&p->fb.rowbase, n_img * rows * 2 * (size_t)p->d.tiles_x * 4);
  alloc((void**)&p->mb.level_hist, 257 * 4);
  alloc((void**)&p->mb.level_cursor, 256 * 4);
  alloc((void**)&p->mb.red_count, 64);
  alloc((void**)&p->mb.ovf_list, ntiles * 4);
  alloc((void**)&p->lvl_hist, n_img * 256 * 4);
  alloc((void**)&p->d_strip_off, 16);
  alloc((void**)&p->mb.unions, n_img * 256 * 4);
  alloc((void**)&p->mb.fin_hist, n_img * 256 * 4);
  alloc((void**)&p->mb.ndistinct, n_img * 4);
  alloc((void**)&p->mb.counts, n_img * 256 * 4);
  alloc((void**)&p->fo.count, 64);
  alloc((void**)&p->fo.tile_hist, n_img * 256 * 4);
  alloc((void**)&p->fo.forest_hist, n_img * 256 * 4);
  alloc((void**)&p->chunk_counts, minima_num_chunks(p->d) * 4);
  alloc((void**)&p->d_total, 16);
  if (e == cudaSuccess) e = cudaMemset(p->fo.count, 0, 64);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_ctrl, (FC_WORDS + 16) * 4);
  if (e == cudaSuccess) e = flood_make_tensor_maps(p->fb, p->d, p->tmaps);
  for (auto& ev : p->ev)
    if (e == cudaSuccess) e = cudaEventCreate(&ev);
  for (auto& ev : p->kev)
    if (e == cudaSuccess) e = cudaEventCreate(&ev);
  if (e != cudaSuccess) {
    ws_plan_destroy(p);
    return cuda_fail(ctx, e, "ws_plan_create");
  }
  *out = p;
  return WS_OK;
}

extern "C" void ws_plan_destroy(ws_plan* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  if (p->ctx->cached == p) p->ctx->cached = nullptr;
  cudaFree(p->fb.T);
  cudaFree(p->fb.pix);
  cudaFree(p->T_dense);
  cudaFree(p->lvl_hist);
  cudaFree(p->d_strip_off);
  cudaFree(p->union_edges);
  cudaFree(p->fb.lab);
  cudaFree(p->fb.lvl);
  cudaFree(p->fb.rim);
  cudaFree(p->fb.qslots);
  cudaFree(p->fb.qmask);
  cudaFree(p->fb.ctrl);
  cudaFree(p->fb.row_start);
  cudaFree(p->fb.rowbase);
  cudaFree(p->mb.level_hist);
  cudaFree(p->mb.level_cursor);
  cudaFree(p->mb.edges);
  cudaFree(p->mb.red_ab);
  cudaFree(p->mb.red_w);
  cudaFree(p->mb.red_count);
  cudaFree(p->mb.ovf_list);
  cudaFree(p->mb.parent);
  cudaFree(p->mb.hook_to);
  cudaFree(p->mb.hook_lvl);
  cudaFree(p->mb.unions);
  cudaFree(p->mb.fin_hist);
  cudaFree(p->mb.ndistinct);
  cudaFree(p->mb.counts);
  cudaFree(p->fo.count);
  cudaFree(p->fo.tile_hist);
  cudaFree(p->fo.forest_hist);
  cudaFree(p->fo.best);
  cudaFree(p->fo.orig[0]);
  cudaFree(p->fo.orig[1]);
  cudaFree(p->fo.ab[1]);
  cudaFree(p->fo.w[0]);
  cudaFree(p->fo.w[1]);
  cudaFree(p->fo.def_ab);
  cudaFree(p->fo.def_w);
  cudaFree(p->rep);
  cudaFree(p->chunk_counts);
  cudaFree(p->d_total);
  if (p->h_ctrl) cudaFreeHost(p->h_ctrl);
  for (auto& ev : p->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : p->kev)
    if (ev) cudaEventDestroy(ev);
  delete p;
}

extern "C" ws_status ws_plan_find_local_minima(ws_plan* p, const uint8_t* d_imgs, uint32_t* d_seeds_rc, size_t cap,
                                               uint32_t* d_seed_off, size_t* out_total) {
  if (!p || !d_imgs || !d_seed_off || !out_total) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t nchunks = minima_num_chunks(p->d);
  WS_CUDA(ctx, launch_minima_count(d_imgs, p->d, p->chunk_counts, s));
  WS_CUDA(ctx, launch_minima_scan(p->chunk_counts, nchunks, p->d, d_seed_off, p->d_total, s));
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl + FC_WORDS, p->d_total, 4, cudaMemcpyDeviceToHost, s));
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  const size_t total = p->h_ctrl[FC_WORDS];
  *out_total = total;
  p->stats[4] += 2;
  if (!d_seeds_rc) return WS_OK;  // count + offsets only
  if (total > cap) return fail(ctx, WS_ERR_TOO_LARGE, "seed buffer too small for the minima found");
  if (total) {
    WS_CUDA(ctx, launch_minima_write(d_imgs, p->d, p->chunk_counts, d_seeds_rc, (uint32_t)cap, s));
    p->stats[4] += 1;
  }
  return WS_OK;
}

// union-find arrays of the level-ordered pass (merge tree) -- also lent to the forest rounds
static ws_status plan_uf_buffers(ws_plan* p, size_t ncolours) {
  ws_ctx* ctx = p->ctx;
  if (ncolours > p->uf_cap || !p->mb.parent) {
    cudaFree(p->mb.parent); cudaFree(p->mb.hook_to); cudaFree(p->mb.hook_lvl);
    p->mb.parent = p->mb.hook_to = nullptr; p->mb.hook_lvl = nullptr; p->uf_cap = 0;
    const size_t n = std::max<size_t>(ncolours, 1);
    WS_CUDA(ctx, cudaMalloc((void**)&p->mb.parent, n * 4));
    WS_CUDA(ctx, cudaMalloc((void**)&p->mb.hook_to, n * 4));
    WS_CUDA(ctx, cudaMalloc((void**)&p->mb.hook_lvl, n));
    p->uf_cap = n;
  }
  return WS_OK;
}

// edge buffers: bounded by (nodes per tile - 1) forest edges per tile, so no host round trip is needed
static ws_status plan_edge_buffers(ws_plan* p, size_t min_cap = 0) {
  ws_ctx* ctx = p->ctx;
  const size_t cap = std::max(merge_reduce_capacity(p->d), min_cap);
  if (cap > p->edges_cap || !p->mb.edges) {
    cudaFree(p->mb.edges); cudaFree(p->mb.red_ab); cudaFree(p->mb.red_w);
    p->mb.edges = p->mb.red_ab = nullptr; p->mb.red_w = nullptr; p->edges_cap = 0;
    WS_CUDA(ctx, cudaMalloc((void**)&p->mb.red_ab, cap * sizeof(uint2)));
    WS_CUDA(ctx, cudaMalloc((void**)&p->mb.red_w, cap));
    WS_CUDA(ctx, cudaMalloc((void**)&p->mb.edges, cap * sizeof(uint2)));
    p->edges_cap = cap;
  }
  return WS_OK;
}

// Buffers of the Boruvka rounds (forest.cu) for `ncolours` nodes and `def_cap` DEFERRED picks (0: one GPU, no
// open nodes).  parent / link / open_ are the union-find arrays of the level-ordered pass (never in use at the
// same time: that pass re-initialises them), ab[0] is its bucket array.
static ws_status plan_forest_buffers(ws_plan* p, size_t ncolours, size_t def_cap, size_t min_edge_cap = 0) {
  ws_ctx* ctx = p->ctx;
  WS_TRY(plan_uf_buffers(p, ncolours));
  WS_TRY(plan_edge_buffers(p, min_edge_cap));
  if (ncolours > p->fo_col_cap || !p->fo.best) {
    cudaFree(p->fo.best);
    p->fo.best = nullptr; p->fo_col_cap = 0;
    const size_t n = std::max<size_t>(ncolours, 1);
    WS_CUDA(ctx, cudaMalloc((void**)&p->fo.best, n * 8));
    p->fo_col_cap = n;
  }
  const size_t edge_cap = p->edges_cap;
  if (edge_cap > p->fo_edge_cap || !p->fo.ab[1]) {
    cudaFree(p->fo.ab[1]); cudaFree(p->fo.w[0]); cudaFree(p->fo.w[1]); cudaFree(p->fo.orig[0]); cudaFree(p->fo.orig[1]);
    p->fo.ab[1] = p->fo.orig[0] = p->fo.orig[1] = nullptr; p->fo.w[0] = p->fo.w[1] = nullptr; p->fo_edge_cap = 0;
    WS_CUDA(ctx, cudaMalloc((void**)&p->fo.ab[1], edge_cap * sizeof(uint2)));
    WS_CUDA(ctx, cudaMalloc((void**)&p->fo.w[0], edge_cap));
    WS_CUDA(ctx, cudaMalloc((void**)&p->fo.w[1], edge_cap));
    p->fo_edge_cap = edge_cap;
  }
  if (def_cap && (!p->fo.orig[0] || def_cap > p->fo_def_cap)) {   // strips only
    cudaFree(p->fo.def_ab); cudaFree(p->fo.def_w);
    p->fo.def_ab = nullptr; p->fo.def_w = nullptr; p->fo_def_cap = 0;
    if (!p->fo.orig[0]) {
      WS_CUDA(ctx, cudaMalloc((void**)&p->fo.orig[0], edge_cap * sizeof(uint2)));
      WS_CUDA(ctx, cudaMalloc((void**)&p->fo.orig[1], edge_cap * sizeof(uint2)));
    }
    WS_CUDA(ctx, cudaMalloc((void**)&p->fo.def_ab, def_cap * sizeof(uint2)));
    WS_CUDA(ctx, cudaMalloc((void**)&p->fo.def_w, def_cap));
    p->fo_def_cap = def_cap;
  }
  p->fo.def_cap = (uint32_t)p->fo_def_cap;
  p->fo.parent = p->mb.parent;
  p->fo.link = p->mb.hook_to;
  p->fo.open_ = p->mb.hook_lvl;
  p->fo.ab[0] = p->mb.edges;
  p->fo.n_deferred = p->fo.count + 2;
  p->fo.rounds = p->fo.count + 3;
  p->fo.error = p->fo.count + 4;
  return WS_OK;
}

static bool merge_by_kruskal() {  // A/B switch of the first version (level-ordered union-find, 255 grid barriers)
  static const bool v = [] { const char* e = getenv("WS_MERGE_KRUSKAL"); return e && atoi(e) != 0; }();
  return v;
}

static ws_status plan_merge(ws_plan* p) {
  ws_ctx* ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  const uint32_t lmax = p->cfg.max_water_level;
  WS_TRY(plan_uf_buffers(p, p->nseeds));
  WS_TRY(plan_edge_buffers(p));
  // Per tile: FINAL forest edges (only counted) and DEFERRED edges between basins that reach the tile's rim.
  WS_CUDA(ctx, launch_merge_reduce(p->fb.lab, p->fb.lvl, p->d, p->seed_off, 1, p->mb.red_ab, p->mb.red_w,
                                   p->mb.red_count, p->mb.ovf_list, s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[7], s));
  p->kev_valid[6] = true;
#ifdef WS_MERGE_STATS
  {
    uint32_t st[16];
    cudaStreamSynchronize(s);
    cudaMemcpy(st, p->mb.red_count, 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "merge stats: edges %u | tiles with edges %u, rounds %.2f per tile, live-edge looks %.1f per tile, "
            "ids %.1f, out %.1f per tile\n", st[0], st[8], st[4] / (double)st[8], st[6] / (double)st[8],
            st[9] / (double)st[8], st[10] / (double)st[8]);
  }
#endif
  if (merge_by_kruskal()) {
    WS_CUDA(ctx, launch_red_sort(p->mb.red_ab, p->mb.red_w, p->mb.red_count, 0, p->seed_off, p->d.n_img,
                                 p->mb.level_hist, p->mb.level_cursor, p->mb.fin_hist, p->mb.edges, s));
    WS_CUDA(ctx, launch_uf_init(p->mb, p->d, (uint32_t)p->nseeds, s));
    WS_CUDA(ctx, cudaEventRecord(p->kev[8], s));
    WS_CUDA(ctx, launch_union_levels(p->mb, p->seed_off, p->d.n_img, lmax, ctx->union_grid, s));
    WS_CUDA(ctx, cudaEventRecord(p->kev[9], s));
    WS_CUDA(ctx, launch_lake_counts(p->mb, p->d.n_img, lmax, s));
    WS_CUDA(ctx, cudaEventRecord(p->kev[10], s));
    for (int i = 7; i < 10; ++i) p->kev_valid[i] = true;
    p->stats[4] += 8;
  } else {
    // The forest of the DEFERRED graph by Boruvka rounds: every node is closed here (the list is the whole
    // graph), so every pick is a forest edge with its true level -- no order, no barrier per level.
    WS_TRY(plan_forest_buffers(p, p->nseeds, 0));
    WS_CUDA(ctx, launch_forest_init(p->fo, (uint32_t)p->nseeds, p->d.n_img, 0, ctx->sms, s));
    WS_CUDA(ctx, cudaEventRecord(p->kev[8], s));
    WS_CUDA(ctx, launch_forest(p->fo, p->mb.red_ab, p->mb.red_w, p->mb.red_count, 1, 0, p->seed_off, p->d.n_img,
                               ctx->forest_grid, ctx->sms, s));
    WS_CUDA(ctx, cudaEventRecord(p->kev[9], s));
    WS_CUDA(ctx, launch_forest_lake_counts(p->mb.ndistinct, p->fo, p->d.n_img, lmax, p->mb.counts, s));
    WS_CUDA(ctx, cudaEventRecord(p->kev[10], s));
    for (int i = 7; i < 10; ++i) p->kev_valid[i] = true;
    p->stats[4] += 5;
  }
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl + FC_WORDS + 1, p->mb.red_count, 4, cudaMemcpyDeviceToHost, s));
  p->merged = true;
  p->tree_built = false;
  p->rep_level = -1;
  return WS_OK;
}

// The merge tree (hook_to / hook_lvl: representative at level L = follow the links of level <= L) needs
// the unions in level order over ALL forest edges, the tile-contracted ones included: built on the first
// request for per-level representatives (merging snapshots, lake sizes, hooks), not for lake counts.
static ws_status plan_build_tree(ws_plan* p) {
  ws_ctx* ctx = p->ctx;
  cudaStream_t s = ctx->stream;
  const uint32_t lmax = p->cfg.max_water_level;
  WS_CUDA(ctx, launch_red_sort(p->mb.red_ab, p->mb.red_w, p->mb.red_count, 1, p->seed_off, p->d.n_img,
                               p->mb.level_hist, p->mb.level_cursor, p->mb.fin_hist, p->mb.edges, s));
  WS_CUDA(ctx, launch_uf_init(p->mb, p->d, (uint32_t)p->nseeds, s));
  WS_CUDA(ctx, launch_union_levels(p->mb, p->seed_off, p->d.n_img, lmax, ctx->union_grid, s));
  p->stats[4] += 6;
  p->tree_built = true;
  p->rep_level = -1;
  return WS_OK;
}

extern "C" ws_status ws_plan_run(ws_plan* p, const ws_config* cfg, const uint8_t* d_imgs, const uint32_t* d_seeds_rc,
                                 const uint32_t* d_seed_off, size_t nseeds_total) {
  if (!p || !d_imgs || !d_seed_off || (nseeds_total && !d_seeds_rc)) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_TRY(check_cfg(ctx, cfg));
  if (nseeds_total >= 0x7fffffffull) return fail(ctx, WS_ERR_TOO_LARGE, "more than 2^31 - 2 seeds");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  p->d.row_offset = 0;  // a plain run: not a strip of a larger field
  p->d.global_rows = p->d.rows;
  p->d.halo_top = p->d.halo_bottom = 0;
  p->cfg = *cfg;
  p->seeds = d_seeds_rc;
  p->seed_off = d_seed_off;
  p->nseeds = nseeds_total;
  p->ran = false;
  p->merged = false;
  for (auto& v : p->stats) v = 0;

  // hop counters can only overflow when a single slice has more than 2^24 pixels
  p->check_ovf = p->d.px_per_img() > (size_t)HOP_MASK ? 1 : 0;
  const int check_ovf = p->check_ovf;
  for (auto& v : p->kev_valid) v = false;
  WS_CUDA(ctx, cudaEventRecord(p->ev[0], s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[0], s));
  p->fb.seed_off = d_seed_off;
  p->fb.colour_base = 0u;
  WS_CUDA(ctx, launch_fill_state(p->fb, p->d, d_imgs, cfg->max_water_level, d_seeds_rc, d_seed_off,
                                 (uint32_t)nseeds_total, ctx->sms, s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[1], s));
  WS_CUDA(ctx, launch_seed_init(p->fb, p->d, d_seeds_rc, d_seed_off, (uint32_t)nseeds_total, 0u, s));
  WS_CUDA(ctx, cudaEventRecord(p->ev[1], s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[2], s));
  p->bucket_shift = flood_bucket_shift(nseeds_total, p->d);
  WS_CUDA(ctx, launch_flood(p->fb, p->d, check_ovf, p->bucket_shift, ctx->flood_grid, p->tmaps, s));
  WS_CUDA(ctx, cudaEventRecord(p->ev[2], s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[3], s));
  WS_CUDA(ctx, launch_parent(p->fb, p->d, p->mb.ndistinct, cfg->tie_break == WS_TIE_RANDOM, ctx->tie_seed, p->tmaps, s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[4], s));
  WS_CUDA(ctx, launch_jump(p->fb, p->d, ctx->jump_grid, 0, s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[5], s));
  WS_CUDA(ctx, launch_label_finish(p->fb, p->d, ctx->sms, s));
  WS_CUDA(ctx, cudaEventRecord(p->kev[6], s));
  for (int i = 0; i < 6; ++i) p->kev_valid[i] = true;
  WS_CUDA(ctx, cudaEventRecord(p->ev[3], s));
  // seeds_scan_empty, fill_rows, fill_state, flood, label_tile, rim_jump, label_finish (+ seeds_scan, seed_init, seed_dup)
  p->stats[4] += 7 + (nseeds_total ? 3 : 0);
  if (cfg->kind == WS_MERGING) WS_TRY(plan_merge(p));
  WS_CUDA(ctx, cudaEventRecord(p->ev[4], s));
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl, p->fb.ctrl, FC_WORDS * 4, cudaMemcpyDeviceToHost, s));
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl + FC_WORDS + 10, p->fo.count + 3, 8, cudaMemcpyDeviceToHost, s));  // rounds, error
  p->h_seed_off.resize((size_t)p->d.n_img + 1);  // colour base of every slice, for merging snapshots
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_seed_off.data(), d_seed_off, ((size_t)p->d.n_img + 1) * 4,
                               cudaMemcpyDeviceToHost, s));
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  p->stats[0] = p->h_ctrl[FC_STALE];
  if (p->merged) p->stats[3] = p->h_ctrl[FC_WORDS + 1];
  p->stats[1] = p->h_ctrl[FC_ACTIVATIONS];
  p->stats[2] = p->h_ctrl[FC_JUMP_ROUNDS];
  p->stats[5] = p->h_ctrl[FC_PHASES];
  p->stats[6] = p->h_ctrl[FC_WAIT_KCYC];
  p->stats[7] = p->h_ctrl[FC_BUSY_KCYC];
  {
    static const bool env = [] { const char* e = getenv("WS_FLOOD_STATS"); return e && atoi(e) != 0; }();
    if (env)
      fprintf(stderr, "[ws] flood: %u activations, %u phases, %u stale; consumers busy %u waiting %u, producers idle %u (kilo-cycles)\n",
              p->h_ctrl[FC_ACTIVATIONS], p->h_ctrl[FC_PHASES], p->h_ctrl[FC_STALE], p->h_ctrl[FC_BUSY_KCYC],
              p->h_ctrl[FC_WAIT_KCYC], p->h_ctrl[FC_IDLE]);
  }
  const uint32_t err = p->h_ctrl[FC_ERROR];
  if (err & 1u) return fail(ctx, WS_ERR_SEED_OOB, "a seed lies outside the image");
  if (err & 2u) return fail(ctx, WS_ERR_HOP_OVERFLOW, ws_status_str(WS_ERR_HOP_OVERFLOW));
  if (err & 4u) return fail(ctx, WS_ERR_INTERNAL, "flood did not reach a fixed point");
  if (err & 8u) {
    char msg[256];
    snprintf(msg, sizeof msg, "flood worklist: a claimed slot was never written (bucket %u slot %u tail %u head %u avail %d cap %u outstanding %u now 0x%x)",
             p->h_ctrl[16], p->h_ctrl[17], p->h_ctrl[18], p->h_ctrl[19], (int)p->h_ctrl[20], p->h_ctrl[21], p->h_ctrl[22], p->h_ctrl[23]);
    return fail(ctx, WS_ERR_INTERNAL, msg);
  }
  if (err & 16u) return fail(ctx, WS_ERR_INTERNAL, "flood worklist watchdog fired");
  if (err & 32u) return fail(ctx, WS_ERR_INTERNAL, "flood worklist: a ring slot was overwritten while in use");
  if (p->merged && !merge_by_kruskal() && p->h_ctrl[FC_WORDS + 11])
    return fail(ctx, WS_ERR_INTERNAL, "forest rounds failed");
  p->ran = true;
  return WS_OK;
}

namespace {  // (defined with the staged copies further down)
ws_status copy_h2d(ws_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
ws_status copy_d2h(ws_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
}  // namespace

extern "C" ws_status ws_dev_malloc(ws_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return WS_ERR_INVALID_ARG;
  *out = nullptr;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_CUDA(ctx, cudaMalloc(out, std::max<size_t>(bytes, 16)));
  return WS_OK;
}
extern "C" ws_status ws_dev_free(ws_ctx* ctx, void* d_ptr) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  WS_CUDA(ctx, cudaFree(d_ptr));
  return WS_OK;
}
extern "C" ws_status ws_memcpy_h2d(ws_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (!ctx || (bytes && (!d_dst || !h_src))) return WS_ERR_INVALID_ARG;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_TRY(copy_h2d(ctx, d_dst, h_src, bytes));   // pageable memory goes through the page-locked ring
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return WS_OK;
}
extern "C" ws_status ws_memcpy_d2h(ws_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (!ctx || (bytes && (!h_dst || !d_src))) return WS_ERR_INVALID_ARG;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_TRY(copy_d2h(ctx, h_dst, d_src, bytes));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return WS_OK;
}

extern "C" ws_status ws_memcpy_d2d(ws_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
  if (!ctx || (bytes && (!d_dst || !d_src))) return WS_ERR_INVALID_ARG;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return WS_OK;
}

extern "C" const uint32_t* ws_plan_arrival_times(const ws_plan* cp) {
  ws_plan* p = const_cast<ws_plan*>(cp);
  if (!p) return nullptr;  // also after a failed run: the diagnostic is most useful then
  if (cudaSetDevice(p->ctx->device) != cudaSuccess) return nullptr;
  if (!p->T_dense && cudaMalloc((void**)&p->T_dense, p->d.px_total() * 4) != cudaSuccess) return nullptr;
  if (launch_unpad_T(p->fb.T, p->d, p->T_dense, p->ctx->stream) != cudaSuccess) return nullptr;
  if (cudaStreamSynchronize(p->ctx->stream) != cudaSuccess) return nullptr;
  return p->T_dense;
}
extern "C" const uint32_t* ws_plan_labels(const ws_plan* p) { return p ? p->fb.lab : nullptr; }
extern "C" const uint8_t* ws_plan_levels(const ws_plan* p) { return p ? p->fb.lvl : nullptr; }
extern "C" const uint32_t* ws_plan_lake_counts(const ws_plan* p) { return p ? p->mb.counts : nullptr; }

static ws_status plan_rep_table(ws_plan* p, uint32_t level) {
  ws_ctx* ctx = p->ctx;
  if (!p->merged) return fail(ctx, WS_ERR_INVALID_ARG, "merging snapshot requested but the last run was not WS_MERGING");
  if (!p->tree_built) WS_TRY(plan_build_tree(p));
  if (p->nseeds > p->rep_cap || !p->rep) {
    cudaFree(p->rep);
    p->rep = nullptr; p->rep_cap = 0; p->rep_level = -1;
    const size_t n = std::max<size_t>(p->nseeds, 1);
    WS_CUDA(ctx, cudaMalloc((void**)&p->rep, n * 4));
    p->rep_cap = n;
  }
  if (p->rep_level == (int)level) return WS_OK;
  const int incremental = (p->rep_level >= 0 && p->rep_level < (int)level) ? 1 : 0;
  WS_CUDA(ctx, launch_rep_table(p->mb.hook_to, p->mb.hook_lvl, (uint32_t)p->nseeds, level, incremental, p->rep,
                                ctx->stream));
  p->rep_level = (int)level;
  p->stats[4] += 1;
  return WS_OK;
}

extern "C" ws_status ws_plan_snapshot(ws_plan* p, ws_kind kind, size_t i, uint8_t level, uint64_t* d_out) {
  if (!p || !d_out) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  if (!p->ran) return fail(ctx, WS_ERR_INVALID_ARG, "no completed run to snapshot");
  if (i >= (size_t)p->d.n_img) return fail(ctx, WS_ERR_INVALID_ARG, "slice index out of range");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = p->d.px_per_img();
  const uint32_t* rep = nullptr;
  uint32_t base = 0;
  if (kind == WS_MERGING) {
    WS_TRY(plan_rep_table(p, level));
    rep = p->rep;
    base = p->h_seed_off.size() > i ? p->h_seed_off[i] : 0;
  }
  WS_CUDA(ctx, launch_snapshot(p->fb.lab + i * n, p->fb.lvl + i * n, n, level, rep, base, d_out, ctx->stream));
  p->stats[4] += 1;
  return WS_OK;
}

// the snapshot of slice 0 as 32-bit words (internal: the pageable history / hook paths)
static ws_status plan_snapshot32(ws_plan* p, ws_kind kind, uint8_t level, uint32_t* d_out) {
  ws_ctx* ctx = p->ctx;
  const size_t n = p->d.px_per_img();
  const uint32_t* rep = nullptr;
  uint32_t base = 0;
  if (kind == WS_MERGING) {
    WS_TRY(plan_rep_table(p, level));
    rep = p->rep;
    base = !p->h_seed_off.empty() ? p->h_seed_off[0] : 0;
  }
  WS_CUDA(ctx, launch_snapshot32(p->fb.lab, p->fb.lvl, n, level, rep, base, d_out, ctx->stream));
  p->stats[4] += 1;
  return WS_OK;
}

// ---------------------------------------------------------------------------
// row-strip decomposition of one field over several plans / GPUs
// ---------------------------------------------------------------------------

namespace {
ws_status check_flood_errors(ws_plan* p) {
  ws_ctx* ctx = p->ctx;
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl, p->fb.ctrl, FC_WORDS * 4, cudaMemcpyDeviceToHost, ctx->stream));
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl + FC_WORDS + 10, p->fo.count + 4, 4, cudaMemcpyDeviceToHost, ctx->stream));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (p->h_ctrl[FC_WORDS + 10]) {
    char msg[128];
    snprintf(msg, sizeof msg, "forest rounds failed (error bits 0x%x)", p->h_ctrl[FC_WORDS + 10]);
    return fail(ctx, WS_ERR_INTERNAL, msg);
  }
  p->stats[0] += p->h_ctrl[FC_STALE];
  p->stats[1] += p->h_ctrl[FC_ACTIVATIONS];
  const uint32_t err = p->h_ctrl[FC_ERROR];
  if (err & 1u) return fail(ctx, WS_ERR_SEED_OOB, "a seed lies outside the strip");
  if (err & 2u) return fail(ctx, WS_ERR_HOP_OVERFLOW, ws_status_str(WS_ERR_HOP_OVERFLOW));
  if (err & 4u) return fail(ctx, WS_ERR_INTERNAL, "flood did not reach a fixed point");
  if (err & 8u) {
    char msg[256];
    snprintf(msg, sizeof msg, "flood worklist: a claimed slot was never written (bucket %u slot %u tail %u head %u avail %d cap %u outstanding %u now 0x%x)",
             p->h_ctrl[16], p->h_ctrl[17], p->h_ctrl[18], p->h_ctrl[19], (int)p->h_ctrl[20], p->h_ctrl[21], p->h_ctrl[22], p->h_ctrl[23]);
    return fail(ctx, WS_ERR_INTERNAL, msg);
  }
  if (err & 16u) return fail(ctx, WS_ERR_INTERNAL, "flood worklist watchdog fired");
  if (err & 32u) return fail(ctx, WS_ERR_INTERNAL, "flood worklist: a ring slot was overwritten while in use");
  return WS_OK;
}
int first_owned(const ws_plan* p) { return p->d.halo_top ? 1 : 0; }
int last_owned(const ws_plan* p) { return p->d.rows - 1 - (p->d.halo_bottom ? 1 : 0); }
}  // namespace

static ws_status strip_begin_impl(ws_plan* p, const ws_config* cfg, const ws_strip* st, const uint8_t* d_img,
                                  const uint32_t* d_seeds_rc, size_t nseeds, bool sync) {
  if (!p || !st || !d_img || (nseeds && !d_seeds_rc)) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_TRY(check_cfg(ctx, cfg));
  if (p->d.n_img != 1) return fail(ctx, WS_ERR_INVALID_ARG, "a strip plan holds one image");
  if (st->row_offset + (size_t)p->d.rows > st->global_rows || (st->halo_top && st->row_offset == 0) ||
      (size_t)p->d.rows < (size_t)(st->halo_top ? 1 : 0) + (st->halo_bottom ? 1 : 0) + 1)
    return fail(ctx, WS_ERR_INVALID_ARG, "inconsistent strip geometry");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  p->d.row_offset = (int)st->row_offset;
  p->d.global_rows = (int)st->global_rows;
  p->d.halo_top = st->halo_top ? 1 : 0;
  p->d.halo_bottom = st->halo_bottom ? 1 : 0;
  p->colour_base = st->colour_base;
  p->cfg = *cfg;
  p->seeds = d_seeds_rc;
  p->seed_off = p->d_strip_off;
  p->nseeds = nseeds;
  p->ran = false;
  p->merged = false;
  for (auto& v : p->stats) v = 0;
  p->h_seed_off.assign(2, 0);
  p->h_seed_off[1] = (uint32_t)nseeds;
  // (a slot of its own: the copy reads it when it executes, which may be after this call has returned)
  p->h_ctrl[FC_WORDS + 8] = 0;
  p->h_ctrl[FC_WORDS + 9] = (uint32_t)nseeds;
  WS_CUDA(ctx, cudaMemcpyAsync(p->d_strip_off, p->h_ctrl + FC_WORDS + 8, 8, cudaMemcpyHostToDevice, s));
  // hop counters cross strip boundaries through the halo exchange: the WHOLE field decides
  p->check_ovf = (size_t)st->global_rows * (size_t)p->d.cols > (size_t)HOP_MASK ? 1 : 0;
  p->fb.seed_off = p->d_strip_off;
  p->fb.colour_base = st->colour_base;
  WS_CUDA(ctx, launch_fill_state(p->fb, p->d, d_img, cfg->max_water_level, d_seeds_rc, p->d_strip_off, (uint32_t)nseeds,
                                 ctx->sms, s));
  WS_CUDA(ctx, launch_seed_init(p->fb, p->d, d_seeds_rc, p->d_strip_off, (uint32_t)nseeds, st->colour_base, s));
  p->bucket_shift = flood_bucket_shift(nseeds, p->d);
  WS_CUDA(ctx, launch_flood(p->fb, p->d, p->check_ovf, p->bucket_shift, ctx->flood_grid, p->tmaps, s));
  p->stats[4] += 3;
  return sync ? check_flood_errors(p) : WS_OK;
}

extern "C" ws_status ws_plan_strip_begin(ws_plan* p, const ws_config* cfg, const ws_strip* st, const uint8_t* d_img,
                                         const uint32_t* d_seeds_rc, size_t nseeds) {
  return strip_begin_impl(p, cfg, st, d_img, d_seeds_rc, nseeds, true);
}
extern "C" ws_status ws_plan_strip_begin_async(ws_plan* p, const ws_config* cfg, const ws_strip* st,
                                               const uint8_t* d_img, const uint32_t* d_seeds_rc, size_t nseeds) {
  return strip_begin_impl(p, cfg, st, d_img, d_seeds_rc, nseeds, false);
}

extern "C" ws_status ws_plan_strip_check(ws_plan* p) {
  if (!p) return WS_ERR_INVALID_ARG;
  WS_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
  return check_flood_errors(p);
}

static ws_status strip_export_times_impl(ws_plan* p, uint32_t* d_top, uint32_t* d_bottom, bool sync) {
  if (!p) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_CUDA(ctx, launch_strip_export_T(p->fb.T, p->d, d_top ? first_owned(p) : -1, d_bottom ? last_owned(p) : -1, d_top,
                                     d_bottom, ctx->stream));
  if (sync) WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller hands the rows to another stream / NCCL
  return WS_OK;
}
extern "C" ws_status ws_plan_strip_export_times(ws_plan* p, uint32_t* d_top, uint32_t* d_bottom) {
  return strip_export_times_impl(p, d_top, d_bottom, true);
}
extern "C" ws_status ws_plan_strip_export_times_async(ws_plan* p, uint32_t* d_top, uint32_t* d_bottom) {
  return strip_export_times_impl(p, d_top, d_bottom, false);
}

static ws_status strip_import_times_impl(ws_plan* p, const uint32_t* d_top, const uint32_t* d_bottom, int* changed,
                                         uint32_t* d_changed) {
  if (!p || (!changed && !d_changed)) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  if (changed) *changed = 0;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // the worklist is empty after a completed flood; reset the statistics, keep the error word
  WS_CUDA(ctx, cudaMemsetAsync(p->fb.ctrl, 0, sizeof(uint32_t) * FC_ERROR, s));
  WS_CUDA(ctx, cudaMemsetAsync(p->fb.ctrl + FC_STRIP_CHANGED, 0, sizeof(uint32_t), s));
  if (d_top && p->d.halo_top) WS_CUDA(ctx, launch_strip_import_T(p->fb, p->d, 0, 1, d_top, p->bucket_shift, s));
  if (d_bottom && p->d.halo_bottom)
    WS_CUDA(ctx, launch_strip_import_T(p->fb, p->d, p->d.rows - 1, p->d.rows - 2, d_bottom, p->bucket_shift, s));
  WS_CUDA(ctx, launch_flood(p->fb, p->d, p->check_ovf, p->bucket_shift, ctx->flood_grid, p->tmaps, s));  // returns at once if nothing woke up
  p->stats[4] += 3;
  if (d_changed) {  // asynchronous: the flag goes to the caller's device word, nothing waits
    WS_CUDA(ctx, launch_ctrl_accumulate(p->fb.ctrl + FC_STRIP_CHANGED, d_changed, 1, s));
    return WS_OK;
  }
  WS_TRY(check_flood_errors(p));
  *changed = p->h_ctrl[FC_STRIP_CHANGED] ? 1 : 0;
  return WS_OK;
}
extern "C" ws_status ws_plan_strip_import_times(ws_plan* p, const uint32_t* d_top, const uint32_t* d_bottom,
                                                int* changed) {
  if (!changed) return WS_ERR_INVALID_ARG;
  return strip_import_times_impl(p, d_top, d_bottom, changed, nullptr);
}
extern "C" ws_status ws_plan_strip_import_times_async(ws_plan* p, const uint32_t* d_top, const uint32_t* d_bottom,
                                                      uint32_t* d_changed) {
  if (!d_changed) return WS_ERR_INVALID_ARG;
  return strip_import_times_impl(p, d_top, d_bottom, nullptr, d_changed);
}

static ws_status strip_labels_impl(ws_plan* p, bool sync) {
  if (!p) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_CUDA(ctx, launch_parent(p->fb, p->d, p->mb.ndistinct, p->cfg.tie_break == WS_TIE_RANDOM, ctx->tie_seed, p->tmaps, ctx->stream));
  // asynchronous protocol: the label plane is finished once, after the last exchange round
  WS_CUDA(ctx, launch_jump(p->fb, p->d, ctx->jump_grid, sync ? 1 : 0, ctx->stream));
  p->stats[4] += 2;
  if (sync) WS_TRY(check_flood_errors(p));
  p->ran = true;
  return WS_OK;
}
extern "C" ws_status ws_plan_strip_labels(ws_plan* p) { return strip_labels_impl(p, true); }
extern "C" ws_status ws_plan_strip_labels_async(ws_plan* p) { return strip_labels_impl(p, false); }

static ws_status strip_export_labels_impl(ws_plan* p, uint32_t* d_top, uint32_t* d_bottom, bool sync) {
  if (!p) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_CUDA(ctx, launch_strip_export_lab(p->fb.lab, p->fb.rim, p->d, d_top ? first_owned(p) : -1,
                                       d_bottom ? last_owned(p) : -1, d_top, d_bottom, ctx->stream));
  if (sync) WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return WS_OK;
}
extern "C" ws_status ws_plan_strip_export_labels(ws_plan* p, uint32_t* d_top, uint32_t* d_bottom) {
  return strip_export_labels_impl(p, d_top, d_bottom, true);
}
extern "C" ws_status ws_plan_strip_export_labels_async(ws_plan* p, uint32_t* d_top, uint32_t* d_bottom) {
  return strip_export_labels_impl(p, d_top, d_bottom, false);
}

static ws_status strip_import_labels_impl(ws_plan* p, const uint32_t* d_top, const uint32_t* d_bottom, size_t* pending,
                                          uint32_t* d_pending) {
  if (!p || (!pending && !d_pending)) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (d_top && p->d.halo_top) WS_CUDA(ctx, launch_strip_import_lab(p->fb, p->d, 0, d_top, s));
  if (d_bottom && p->d.halo_bottom) WS_CUDA(ctx, launch_strip_import_lab(p->fb, p->d, p->d.rows - 1, d_bottom, s));
  WS_CUDA(ctx, cudaMemsetAsync(p->fb.ctrl + FC_JUMP_FLAG0, 0, sizeof(uint32_t) * 3, s));
  WS_CUDA(ctx, cudaMemsetAsync(p->fb.ctrl + FC_STRIP_PENDING, 0, sizeof(uint32_t), s));
  if (d_pending) {  // asynchronous: only the rim array (a tenth of the label plane) is touched per round
    WS_CUDA(ctx, launch_jump(p->fb, p->d, ctx->jump_grid, 0, s));
    WS_CUDA(ctx, launch_strip_count_pending_rim(p->fb.rim, rim_words(p->d), p->fb.ctrl, ctx->sms, s));
  } else {
    WS_CUDA(ctx, launch_jump(p->fb, p->d, ctx->jump_grid, 1, s));
    WS_CUDA(ctx, launch_strip_count_pending(p->fb.lab, p->d, first_owned(p), last_owned(p), p->fb.ctrl, s));
  }
  p->stats[4] += 4;
  if (d_pending) {
    WS_CUDA(ctx, launch_ctrl_accumulate(p->fb.ctrl + FC_STRIP_PENDING, d_pending, 0, s));
    return WS_OK;
  }
  WS_TRY(check_flood_errors(p));
  *pending = p->h_ctrl[FC_STRIP_PENDING];
  return WS_OK;
}
extern "C" ws_status ws_plan_strip_import_labels(ws_plan* p, const uint32_t* d_top, const uint32_t* d_bottom,
                                                 size_t* pending) {
  if (!pending) return WS_ERR_INVALID_ARG;
  return strip_import_labels_impl(p, d_top, d_bottom, pending, nullptr);
}
extern "C" ws_status ws_plan_strip_import_labels_async(ws_plan* p, const uint32_t* d_top, const uint32_t* d_bottom,
                                                       uint32_t* d_pending) {
  if (!d_pending) return WS_ERR_INVALID_ARG;
  return strip_import_labels_impl(p, d_top, d_bottom, nullptr, d_pending);
}

// after the last exchange round of the asynchronous protocol: every label word takes its rim entry
extern "C" ws_status ws_plan_strip_labels_finish_async(ws_plan* p) {
  if (!p) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  WS_CUDA(ctx, launch_label_finish(p->fb, p->d, ctx->sms, ctx->stream));
  p->stats[4] += 1;
  return WS_OK;
}

// ---- strips, merging: the strip's FINAL counts and its boundary picks as one device-resident packet ----------
// packet = uint32 header[260] { edges, colours present, error bits, 0, FINAL edges per level [256] }
//          + cap x (colour a - 1, colour b - 1) uint32 pairs + cap level bytes,   cap = 3 * cols + 64
// (at most one DEFERRED pick per boundary basin, and those live in at most three rows)
static size_t strip_packet_cap(size_t cols) { return (3 * cols + 64 + 15) & ~(size_t)15; }
extern "C" size_t ws_strip_packet_bytes(size_t cols) { return 260 * 4 + strip_packet_cap(cols) * 9; }

extern "C" ws_status ws_plan_strip_forest(ws_plan* p, size_t ncolours_total, void* d_packet) {
  if (!p || !d_packet) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  if (!p->ran) return fail(ctx, WS_ERR_INVALID_ARG, "labels are not resolved yet");
  if (ncolours_total >= 0x7fffffffull) return fail(ctx, WS_ERR_TOO_LARGE, "colour count too large");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t cap = strip_packet_cap((size_t)p->d.cols);
  WS_TRY(plan_forest_buffers(p, ncolours_total, cap));
  // labels are GLOBAL colours here (colour_base + i + 1), so the colour id is label - 1: offsets {0, ...}
  WS_CUDA(ctx, launch_merge_reduce(p->fb.lab, p->fb.lvl, p->d, p->d_strip_off, 1, p->mb.red_ab, p->mb.red_w,
                                   p->mb.red_count, p->mb.ovf_list, s));
  WS_CUDA(ctx, launch_count_present(p->fb.lab, p->d, p->seeds, (uint32_t)p->nseeds, p->colour_base, p->mb.ndistinct, s));
  WS_CUDA(ctx, launch_forest_init(p->fo, (uint32_t)ncolours_total, 1, 1, ctx->sms, s));
  // open = the basins other strips may hold edges of: those on the halo rows, and on the first owned row below a
  // halo (its upward edges belong to the strip above) -- the rows merge_reduce treats as rim
  const size_t cols = (size_t)p->d.cols;
  if (p->d.halo_top)
    WS_CUDA(ctx, launch_forest_mark_open(p->fb.lab, 2 * cols, (uint32_t)ncolours_total, p->fo.open_, s));
  if (p->d.halo_bottom)
    WS_CUDA(ctx, launch_forest_mark_open(p->fb.lab + (size_t)(p->d.rows - 1) * cols, cols, (uint32_t)ncolours_total,
                                         p->fo.open_, s));
  WS_CUDA(ctx, launch_forest(p->fo, p->mb.red_ab, p->mb.red_w, p->mb.red_count, 1, 1, p->d_strip_off, 1,
                             ctx->forest_grid, ctx->sms, s));
  WS_CUDA(ctx, launch_strip_packet(p->fo, p->mb.ndistinct, (uint32_t)cap, d_packet, s));
  p->stats[4] += 8;
  p->merged = false;
  return WS_OK;
}

// All strips' packets (gathered, back to back) -> lakes per level of the whole field in ws_plan_lake_counts().
extern "C" ws_status ws_plan_forest_packets(ws_plan* p, const void* d_packets, size_t n_packets, size_t ncolours_total,
                                            uint8_t max_water_level) {
  if (!p || !d_packets || n_packets == 0) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  if (ncolours_total >= 0x7fffffffull) return fail(ctx, WS_ERR_TOO_LARGE, "colour count too large");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t cap = strip_packet_cap((size_t)p->d.cols);
  WS_TRY(plan_forest_buffers(p, ncolours_total, cap, n_packets * cap));
  WS_CUDA(ctx, launch_forest_init(p->fo, (uint32_t)ncolours_total, 1, 0, ctx->sms, s));
  // edges -> ab[0] / w[0] (consumed by the first pass before the rounds reuse them), FINAL counts and colours summed
  WS_CUDA(ctx, launch_strip_unpack(d_packets, (uint32_t)n_packets, (uint32_t)cap, p->fo, p->fo.count + 5,
                                   p->mb.ndistinct, s));
  WS_CUDA(ctx, launch_forest(p->fo, p->fo.ab[0], p->fo.w[0], p->fo.count + 5, 0, 0, p->d_strip_off, 1,
                             ctx->forest_grid, ctx->sms, s));
  WS_CUDA(ctx, launch_forest_lake_counts(p->mb.ndistinct, p->fo, 1, max_water_level, p->mb.counts, s));
  p->stats[4] += 5;
  return WS_OK;
}

extern "C" ws_status ws_plan_strip_edges(ws_plan* p, const void** d_ab, const void** d_w, size_t* n,
                                         uint32_t* ndistinct) {
  if (!p || !d_ab || !d_w || !n || !ndistinct) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  if (!p->ran) return fail(ctx, WS_ERR_INVALID_ARG, "labels are not resolved yet");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  WS_TRY(plan_edge_buffers(p));
  // labels are GLOBAL colours here (colour_base + i + 1), so the colour id is label - 1: offsets {0, ...}
  WS_CUDA(ctx, launch_merge_reduce(p->fb.lab, p->fb.lvl, p->d, p->d_strip_off, 1, p->mb.red_ab, p->mb.red_w,
                                   p->mb.red_count, p->mb.ovf_list, s));
  WS_CUDA(ctx, launch_count_present(p->fb.lab, p->d, p->seeds, (uint32_t)p->nseeds, p->colour_base, p->mb.ndistinct, s));
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl + FC_WORDS, p->mb.red_count, 4, cudaMemcpyDeviceToHost, s));
  WS_CUDA(ctx, cudaMemcpyAsync(p->h_ctrl + FC_WORDS + 1, p->mb.ndistinct, 4, cudaMemcpyDeviceToHost, s));
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  p->stats[4] += 2;
  *d_ab = p->mb.red_ab;
  *d_w = p->mb.red_w;
  *n = p->h_ctrl[FC_WORDS];
  *ndistinct = p->h_ctrl[FC_WORDS + 1];
  p->stats[3] = *n;
  return WS_OK;
}

extern "C" ws_status ws_plan_union_edges(ws_plan* p, const void* d_ab, const void* d_w, size_t n, size_t ncolours,
                                         uint32_t ndistinct, uint8_t max_water_level) {
  if (!p || (n && (!d_ab || !d_w))) return WS_ERR_INVALID_ARG;
  ws_ctx* ctx = p->ctx;
  if (ncolours >= 0x7fffffffull || n >= 0xffffffffull) return fail(ctx, WS_ERR_TOO_LARGE, "edge or colour count too large");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  WS_TRY(plan_uf_buffers(p, ncolours));
  if (n > p->union_edges_cap || !p->union_edges) {
    cudaFree(p->union_edges);
    p->union_edges = nullptr; p->union_edges_cap = 0;
    WS_CUDA(ctx, cudaMalloc((void**)&p->union_edges, std::max<size_t>(n, 16) * sizeof(uint2)));
    p->union_edges_cap = std::max<size_t>(n, 16);
  }
  p->h_ctrl[FC_WORDS + 2] = (uint32_t)n;
  p->h_ctrl[FC_WORDS + 3] = ndistinct;
  WS_CUDA(ctx, cudaMemcpyAsync(p->mb.red_count, p->h_ctrl + FC_WORDS + 2, 4, cudaMemcpyHostToDevice, s));
  MergeBuffers m = p->mb;
  m.edges = p->union_edges;
  WS_CUDA(ctx, launch_red_sort((const uint2*)d_ab, (const uint8_t*)d_w, m.red_count, 1, p->d_strip_off, 1, m.level_hist,
                               m.level_cursor, m.fin_hist, m.edges, s));
  WS_CUDA(ctx, cudaMemsetAsync(m.fin_hist, 0, 256 * sizeof(uint32_t), s));  // every edge handed in is unioned
  WS_CUDA(ctx, launch_uf_reset(m, (uint32_t)ncolours, s));
  WS_CUDA(ctx, cudaMemcpyAsync(m.ndistinct, p->h_ctrl + FC_WORDS + 3, 4, cudaMemcpyHostToDevice, s));
  WS_CUDA(ctx, launch_union_levels(m, p->d_strip_off, 1, max_water_level, ctx->union_grid, s));
  WS_CUDA(ctx, launch_lake_counts(m, 1, max_water_level, s));
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  p->stats[4] += 7;
  return WS_OK;
}

extern "C" ws_status ws_plan_phase_ms(ws_plan* p, float out[4]) {
  if (!p || !out) return WS_ERR_INVALID_ARG;
  if (!p->ran) return fail(p->ctx, WS_ERR_INVALID_ARG, "no completed run");
  for (int i = 0; i < 4; ++i) WS_CUDA(p->ctx, cudaEventElapsedTime(&out[i], p->ev[i], p->ev[i + 1]));
  return WS_OK;
}

extern "C" ws_status ws_plan_kernel_ms(ws_plan* p, float out[WS_KERNEL_SLOTS]) {
  if (!p || !out) return WS_ERR_INVALID_ARG;
  if (!p->ran) return fail(p->ctx, WS_ERR_INVALID_ARG, "no completed run");
  for (int i = 0; i < WS_KERNEL_SLOTS; ++i) {
    out[i] = 0.0f;
    if (p->kev_valid[i]) WS_CUDA(p->ctx, cudaEventElapsedTime(&out[i], p->kev[i], p->kev[i + 1]));
  }
  return WS_OK;
}

extern "C" ws_status ws_plan_stats(ws_plan* p, uint64_t out[8]) {
  if (!p || !out) return WS_ERR_INVALID_ARG;
  for (int i = 0; i < 8; ++i) out[i] = p->stats[i];
  return WS_OK;
}

// ---------------------------------------------------------------------------
// host-level entry points (the reference-facing ones)
// ---------------------------------------------------------------------------

namespace {

ws_status check_image(ws_ctx* ctx, const ws_image* img) {
  if (!img || !img->data) return fail(ctx, WS_ERR_INVALID_ARG, "image is NULL");
  if (img->rows == 0 || img->cols == 0) return fail(ctx, WS_ERR_INVALID_ARG, "empty image");
  return WS_OK;
}

// WS_TRACE=1: wall time of the phases of a host-level call on stderr (synchronises between the phases)
struct Trace {
  bool on;
  ws_ctx* ctx;
  std::chrono::steady_clock::time_point t;
  explicit Trace(ws_ctx* c) : ctx(c) {
    static const bool env = [] { const char* e = getenv("WS_TRACE"); return e && atoi(e) != 0; }();
    on = env;
    if (on) t = std::chrono::steady_clock::now();
  }
  void lap(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[ws trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

// ---- staged copies between pageable caller memory and the device (hostpipe.h) --------------------------
// Ring geometry, measured on the pool's hosts (one 16384^2 segmenting transform into pageable usize labels, medians;
// scripts/hostbench/e2e_ab.py): 4 x 32 MB 50 ms, 4 x 16 MB 46 ms, 4 x 8 MB 44 ms, 6 x 8 MB 35.8 ms, 8 x 8 MB 35.5 ms,
// 8 x 4 MB 42 ms, 8 x 16 MB 43 ms.  The device -> host direction keeps RING_SLOTS - 1 copies queued ahead of the
// workers: with three of them the link idled between a slot's arrival and the next copy's start; slots below
// 8 MB pay for the hand-over to the worker threads (one wake-up and join per slot).
#ifndef WS_RING_SLOT_MB
#define WS_RING_SLOT_MB 8
#endif
constexpr size_t RING_SLOT_BYTES = (size_t)WS_RING_SLOT_MB << 20;
#ifndef WS_RING_SLOTS
#define WS_RING_SLOTS 8
#endif
constexpr int RING_SLOTS = WS_RING_SLOTS;   // (at most 8: ws_ctx::ring_ev)

ws_status ensure_ring(ws_ctx* ctx) {
  if (!ctx->host_threads) {
    const int n = ctx->host_threads_wanted > 0 ? ctx->host_threads_wanted : host_threads_default();
    ctx->pool.start(n);
    ctx->host_threads = n;
  }
  if (!ctx->ring) {
    WS_CUDA(ctx, cudaHostAlloc((void**)&ctx->ring, RING_SLOT_BYTES * RING_SLOTS, cudaHostAllocDefault));
    for (int i = 0; i < RING_SLOTS; ++i) WS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
  }
  return WS_OK;
}

// a chunk of `n` bytes split over the pool in pieces that are multiples of 64 bytes
template <typename F>
void pool_split(ws_ctx* ctx, size_t n, F&& piece) {
  const size_t want = std::max<size_t>(1, std::min<size_t>((size_t)ctx->pool.size(), n >> 18));
  const size_t per = ((n + want - 1) / want + 63) & ~(size_t)63;
  const size_t parts = (n + per - 1) / per;
  ctx->pool.run(parts, [&](size_t t) {
    const size_t lo = t * per, hi = std::min(n, lo + per);
    if (lo < hi) piece(lo, hi - lo);
  });
}

// Host -> device through the ring: fill(slot pointer, byte offset in the destination, bytes) is run by the
// workers for every piece while the copy engine moves the previous slot.
template <typename Fill>
ws_status staged_h2d(ws_ctx* ctx, void* d_dst, size_t total, Fill fill) {
  if (total == 0) return WS_OK;
  WS_TRY(ensure_ring(ctx));
  cudaStream_t s = ctx->stream;
  const size_t nchunks = (total + RING_SLOT_BYTES - 1) / RING_SLOT_BYTES;
  for (size_t k = 0; k < nchunks; ++k) {
    // the slot's last copy -- of this call or of an earlier one (the image, when the seeds follow) -- must
    // have left it before the workers write into it again
    const int slot = (int)(ctx->ring_pos++ % RING_SLOTS);
    if (ctx->ring_used[slot]) WS_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));
    const size_t off = k * RING_SLOT_BYTES, n = std::min(RING_SLOT_BYTES, total - off);
    uint8_t* sp = ctx->ring + (size_t)slot * RING_SLOT_BYTES;
    pool_split(ctx, n, [&](size_t lo, size_t len) { fill(sp + lo, off + lo, len); });
    WS_CUDA(ctx, cudaMemcpyAsync((char*)d_dst + off, sp, n, cudaMemcpyHostToDevice, s));
    WS_CUDA(ctx, cudaEventRecord(ctx->ring_ev[slot], s));
    ctx->ring_used[slot] = true;
  }
  return WS_OK;
}

// Device -> host through the ring, RING_SLOTS - 1 copies in flight ahead of the workers:
// drain(slot pointer, byte offset in the source, bytes).
template <typename Drain>
ws_status staged_d2h(ws_ctx* ctx, const void* d_src, size_t total, Drain drain) {
  if (total == 0) return WS_OK;
  WS_TRY(ensure_ring(ctx));
  cudaStream_t s = ctx->stream;
  // (a slot still being read by an earlier staged_h2d is safe: these copies are behind it on the same stream)
  const size_t nchunks = (total + RING_SLOT_BYTES - 1) / RING_SLOT_BYTES;
  for (size_t k = 0; k < nchunks + RING_SLOTS - 1; ++k) {
    if (k < nchunks) {
      const int slot = (int)(k % RING_SLOTS);
      ctx->ring_used[slot] = true;
      const size_t off = k * RING_SLOT_BYTES, n = std::min(RING_SLOT_BYTES, total - off);
      WS_CUDA(ctx, cudaMemcpyAsync(ctx->ring + (size_t)slot * RING_SLOT_BYTES, (const char*)d_src + off, n,
                                   cudaMemcpyDeviceToHost, s));
      WS_CUDA(ctx, cudaEventRecord(ctx->ring_ev[slot], s));
    }
    if (k + 1 >= (size_t)RING_SLOTS) {
      const size_t j = k + 1 - RING_SLOTS;
      if (j < nchunks) {
        const int slot = (int)(j % RING_SLOTS);
        const size_t off = j * RING_SLOT_BYTES, n = std::min(RING_SLOT_BYTES, total - off);
        WS_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));
        const uint8_t* sp = ctx->ring + (size_t)slot * RING_SLOT_BYTES;
        pool_split(ctx, n, [&](size_t lo, size_t len) { drain(sp + lo, off + lo, len); });
      }
    }
  }
  return WS_OK;
}

// plain bytes: page-locked memory goes straight over the link, pageable memory through the ring
ws_status copy_h2d(ws_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (bytes == 0) return WS_OK;
  if (bytes < ((size_t)1 << 20) || host_ptr_is_pinned(h_src)) {
    WS_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return WS_OK;
  }
  return staged_h2d(ctx, d_dst, bytes, [&](uint8_t* sp, size_t off, size_t n) { memcpy(sp, (const char*)h_src + off, n); });
}
ws_status copy_d2h(ws_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return WS_OK;
  if (bytes < ((size_t)1 << 20) || host_ptr_is_pinned(h_dst)) {
    WS_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return WS_OK;
  }
  return staged_d2h(ctx, d_src, bytes, [&](const uint8_t* sp, size_t off, size_t n) { memcpy((char*)h_dst + off, sp, n); });
}

// Upload an ArrayView2<u8> (arbitrary strides) as a dense C-order device image.
ws_status upload_image(ws_ctx* ctx, const ws_image* img, uint8_t* dst) {
  cudaStream_t s = ctx->stream;
  const size_t rows = img->rows, cols = img->cols;
  if (img->col_stride == 1 && img->row_stride == (ptrdiff_t)cols) return copy_h2d(ctx, dst, img->data, rows * cols);
  if (img->col_stride == 1 && img->row_stride > (ptrdiff_t)cols && host_ptr_is_pinned(img->data)) {
    WS_CUDA(ctx, cudaMemcpy2DAsync(dst, cols, img->data, (size_t)img->row_stride, cols, rows,
                                   cudaMemcpyHostToDevice, s));
    return WS_OK;
  }
  // general strides (sub-views, transposed / reversed views): the workers gather into the ring
  const uint8_t* base = img->data;
  const ptrdiff_t rs = img->row_stride, cs = img->col_stride;
  return staged_h2d(ctx, dst, rows * cols, [=](uint8_t* sp, size_t off, size_t n) {
    size_t r = off / cols, c = off - r * cols;
    for (size_t i = 0; i < n; ++i) {
      sp[i] = base[(ptrdiff_t)r * rs + (ptrdiff_t)c * cs];
      if (++c == cols) { c = 0; ++r; }
    }
  });
}

// Final segmenting labels of `n` pixels as usize into caller memory.  Pageable destination (an ordinary
// Array2<usize>): the u32 label words cross the link as they are and the workers widen them out of the ring
// -- half the bytes on the link and no extra pass over the destination.  Page-locked destination: widened
// on the device chunk by chunk, copied by the second copy engine while the next chunk is widened (or the
// pageable way if the caller asked for it: WS_OPT_PINNED_HOST_WIDEN).
ws_status download_labels_u64(ws_ctx* ctx, ws_plan* p, const uint32_t* d_lab, size_t n, uint64_t* out) {
  cudaStream_t s = ctx->stream;
  Trace tr(ctx);
  if (!host_ptr_is_pinned(out) || ctx->pinned_host_widen) {
    WS_TRY(staged_d2h(ctx, d_lab, n * 4, [=](const uint8_t* sp, size_t off, size_t len) {
      widen_labels_host(out + off / 4, reinterpret_cast<const uint32_t*>(sp), len / 4);
    }));
    WS_CUDA(ctx, cudaStreamSynchronize(s));
    tr.lap("labels: u32 down + host widen");
    return WS_OK;
  }
  const size_t chunk = (size_t)8 << 20;  // pixels per chunk: 64 MB of usize
  for (int i = 0; i < 2; ++i) WS_TRY(grow(ctx, ctx->d_out[i], ctx->d_out_cap[i], std::min(chunk, n)));
  size_t k = 0;
  for (size_t off = 0; off < n; off += chunk, ++k) {
    const int buf = (int)(k & 1);
    const size_t m = std::min(chunk, n - off);
    if (k >= 2) WS_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_copied[buf], 0));
    WS_CUDA(ctx, launch_widen_labels(d_lab + off, m, ctx->d_out[buf], s));
    WS_CUDA(ctx, cudaEventRecord(ctx->ev_ready[buf], s));
    WS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready[buf], 0));
    WS_CUDA(ctx, cudaMemcpyAsync(out + off, ctx->d_out[buf], m * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
    WS_CUDA(ctx, cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream));
    p->stats[4] += 1;
  }
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  tr.lap("labels: device widen + u64 down");
  return WS_OK;
}

ws_status get_plan(ws_ctx* ctx, size_t n_img, size_t rows, size_t cols, ws_plan** out) {
  ws_plan* p = ctx->cached;
  if (p && p->d.n_img == (int)n_img && p->d.rows == (int)rows && p->d.cols == (int)cols) {
    *out = p;
    return WS_OK;
  }
  if (p) ws_plan_destroy(p);
  ctx->cached = nullptr;
  WS_TRY(ws_plan_create(ctx, n_img, rows, cols, &p));
  ctx->cached = p;
  *out = p;
  return WS_OK;
}

// Everything the host-level calls share: validate, upload image (+pad) and seeds, run.
struct HostRun {
  ws_plan* plan = nullptr;
  size_t orows = 0, ocols = 0, npx = 0;
  size_t nseeds = 0;  // as given, or as found on the device (WS_SEEDS_AUTO)
};

ws_status host_run_batch(ws_ctx* ctx, const ws_config* cfg, const uint8_t* imgs, const ws_image* view, size_t n_img,
                         size_t rows, size_t cols, const uint64_t* seeds_rc, const uint64_t* seed_offsets,
                         size_t nseeds, HostRun* hr) {
  WS_TRY(check_cfg(ctx, cfg));
  const bool auto_seeds = nseeds == WS_SEEDS_AUTO;
  if (!auto_seeds && nseeds && !seeds_rc) return fail(ctx, WS_ERR_INVALID_ARG, "seeds is NULL");
  if (!auto_seeds && nseeds >= 0x7fffffffull) return fail(ctx, WS_ERR_TOO_LARGE, "more than 2^31 - 2 seeds");
  // the reference's caller finds the seeds on the UNPADDED image and hands them over unshifted
  // (lib.rs:1365-1367): that needs the caller's own list
  if (auto_seeds && cfg->edge_correction)
    return fail(ctx, WS_ERR_INVALID_ARG, "WS_SEEDS_AUTO is not available with edge correction");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t orows, ocols;
  ws_output_shape(cfg, rows, cols, &orows, &ocols);
  ws_plan* p = nullptr;
  WS_TRY(get_plan(ctx, n_img, orows, ocols, &p));
  cudaStream_t s = ctx->stream;
  const size_t npx_in = rows * cols, npx_out = orows * ocols;
  WS_TRY(grow(ctx, ctx->d_img, ctx->d_img_cap, n_img * npx_out));
  if (cfg->edge_correction) {
    WS_TRY(grow(ctx, ctx->d_raw, ctx->d_raw_cap, n_img * npx_in));
    if (view) WS_TRY(upload_image(ctx, view, ctx->d_raw));
    else WS_TRY(copy_h2d(ctx, ctx->d_raw, imgs, n_img * npx_in));
    for (size_t b = 0; b < n_img; ++b)
      WS_CUDA(ctx, launch_pad_image(ctx->d_raw + b * npx_in, (int)rows, (int)cols, ctx->d_img + b * npx_out, s));
  } else {
    if (view) WS_TRY(upload_image(ctx, view, ctx->d_img));
    else WS_TRY(copy_h2d(ctx, ctx->d_img, imgs, n_img * npx_in));
  }
  Trace tr(ctx);
  tr.lap("image upload (+ earlier work)");
  WS_TRY(grow(ctx, ctx->d_seed_off, ctx->d_seed_off_cap, n_img + 1));
  if (auto_seeds) {
    // WatershedUtils::find_local_minima on the device: nothing but the image crosses the link
    size_t total = 0;
    WS_TRY(ws_plan_find_local_minima(p, ctx->d_img, nullptr, 0, ctx->d_seed_off, &total));
    if (total >= 0x7fffffffull) return fail(ctx, WS_ERR_TOO_LARGE, "more than 2^31 - 2 seeds");
    WS_TRY(grow(ctx, ctx->d_seeds, ctx->d_seeds_cap, 2 * total));
    if (total) WS_CUDA(ctx, launch_minima_write(ctx->d_img, p->d, p->chunk_counts, ctx->d_seeds, (uint32_t)total, s));
    nseeds = total;
  } else {
    // seeds: the (usize, usize) pairs become u32 pairs, those outside the (padded) output shape are marked
    // -- the reference indexes output[seed] and panics there.  Page-locked lists go up as they are and a
    // kernel narrows them; pageable lists are narrowed by the workers on their way into the ring.
    p->h_seed_off.assign(n_img + 1, 0);
    if (seed_offsets) {
      for (size_t b = 0; b <= n_img; ++b) p->h_seed_off[b] = (uint32_t)seed_offsets[b];
    } else {
      p->h_seed_off[n_img] = (uint32_t)nseeds;
    }
    WS_TRY(grow(ctx, ctx->d_seeds, ctx->d_seeds_cap, 2 * nseeds));
    if (nseeds) {
      if (nseeds * 16 < ((size_t)1 << 20) || host_ptr_is_pinned(seeds_rc)) {
        WS_TRY(grow(ctx, ctx->d_seeds64, ctx->d_seeds64_cap, 2 * nseeds));
        WS_CUDA(ctx, cudaMemcpyAsync(ctx->d_seeds64, seeds_rc, 2 * nseeds * 8, cudaMemcpyHostToDevice, s));
        WS_CUDA(ctx, launch_seeds_convert(ctx->d_seeds64, ctx->d_seeds, nseeds, orows, ocols, s));
      } else {
        const uint64_t lr = orows, lc = ocols;
        WS_TRY(staged_h2d(ctx, ctx->d_seeds, 2 * nseeds * 4, [=](uint8_t* sp, size_t off, size_t n) {
          narrow_seeds_host(reinterpret_cast<uint32_t*>(sp), seeds_rc + off / 4, n / 4, off / 4, lr, lc);
        }));
      }
    }
    WS_CUDA(ctx, cudaMemcpyAsync(ctx->d_seed_off, p->h_seed_off.data(), (n_img + 1) * 4, cudaMemcpyHostToDevice, s));
  }
  tr.lap("seeds upload / search");
  WS_TRY(ws_plan_run(p, cfg, ctx->d_img, ctx->d_seeds, ctx->d_seed_off, nseeds));
  tr.lap("device run");
  hr->plan = p;
  hr->orows = orows;
  hr->ocols = ocols;
  hr->npx = npx_out;
  hr->nseeds = nseeds;
  return WS_OK;
}

ws_status host_run(ws_ctx* ctx, const ws_config* cfg, const ws_image* img, const uint64_t* seeds_rc, size_t nseeds,
                   HostRun* hr) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  WS_TRY(check_image(ctx, img));
  return host_run_batch(ctx, cfg, nullptr, img, 1, img->rows, img->cols, seeds_rc, nullptr, nseeds, hr);
}

// Per-level snapshots streamed to the host, one level being cut on the device while the previous one crosses the
// link.  `sink` receives (level, host pointer valid during the call) -- or the snapshots go to `direct` + level * npx
// when that is given.
template <typename Sink>
ws_status stream_snapshots(ws_ctx* ctx, HostRun& hr, const ws_config* cfg, uint64_t* direct, Sink sink) {
  ws_plan* p = hr.plan;
  const size_t npx = hr.npx;
  const uint32_t nlev = (uint32_t)cfg->max_water_level + 1u;
  // Caller memory that is page-locked takes the copies directly.  Pageable memory would make the driver
  // stage every copy itself (~20 GB/s measured): it goes through the ring and the worker threads instead.
  bool staged = true;
  if (direct) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, direct) == cudaSuccess && at.type == cudaMemoryTypeHost) staged = false;
    cudaGetLastError();
  }
  for (int i = 0; i < 2; ++i) {
    WS_TRY(grow(ctx, ctx->d_out[i], ctx->d_out_cap[i], npx));
    if (staged && !direct && i == 0 && ctx->h_pin_cap[i] < npx) {
      if (ctx->h_pin[i]) cudaFreeHost(ctx->h_pin[i]);
      ctx->h_pin[i] = nullptr; ctx->h_pin_cap[i] = 0;
      WS_CUDA(ctx, cudaMallocHost((void**)&ctx->h_pin[i], npx * 8));
      ctx->h_pin_cap[i] = npx;
    }
  }
  const ws_kind kind = (ws_kind)cfg->kind;
  if (staged) {
    // Pageable destination (an ordinary Vec<Array2<usize>>) or a hook: the snapshots cross the link as 32-bit words
    // through the ring and the workers widen them into place (the caller's array, or one page-locked image for the
    // hook) -- 4 B per pixel on the link and 12 B of host traffic instead of 8 and 24.  Level l + 1 is cut while
    // level l is on its way.
    WS_TRY(plan_snapshot32(p, kind, 0, reinterpret_cast<uint32_t*>(ctx->d_out[0])));
    for (uint32_t l = 0; l < nlev; ++l) {
      const int buf = l & 1;
      if (l + 1 < nlev) WS_TRY(plan_snapshot32(p, kind, (uint8_t)(l + 1), reinterpret_cast<uint32_t*>(ctx->d_out[buf ^ 1])));
      uint64_t* dst = direct ? direct + (size_t)l * npx : ctx->h_pin[0];
      WS_TRY(staged_d2h(ctx, ctx->d_out[buf], npx * 4, [=](const uint8_t* sp, size_t off, size_t len) {
        widen_labels_host(dst + off / 4, reinterpret_cast<const uint32_t*>(sp), len / 4);
      }));
      if (!direct) sink((uint8_t)l, ctx->h_pin[0]);
    }
    WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WS_OK;
  }
  for (uint32_t l = 0; l < nlev; ++l) {  // page-locked destination: straight over the link, two levels in flight
    const int buf = l & 1;
    if (l >= 2) WS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[buf], 0));
    WS_TRY(ws_plan_snapshot(p, kind, 0, (uint8_t)l, ctx->d_out[buf]));
    WS_CUDA(ctx, cudaEventRecord(ctx->ev_ready[buf], ctx->stream));
    WS_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_ready[buf], 0));
    WS_CUDA(ctx, cudaMemcpyAsync(direct + (size_t)l * npx, ctx->d_out[buf], npx * 8, cudaMemcpyDeviceToHost, ctx->copy_stream));
    WS_CUDA(ctx, cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream));
  }
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return WS_OK;
}

}  // namespace

static size_t dtype_size(int dtype) {
  switch (dtype) {
    case WS_F32: case WS_I32: case WS_F32_BE: case WS_I32_BE: return 4;
    case WS_F64: case WS_I64: case WS_F64_BE: case WS_I64_BE: return 8;
    case WS_U16: case WS_I16: case WS_I16_BE: return 2;
    case WS_U8: return 1;
  }
  return 0;
}

extern "C" ws_status ws_dev_pre_processor(ws_ctx* ctx, ws_dtype dtype, const void* d_data, size_t n, uint8_t max_value,
                                          uint8_t* d_out) {
  if (!ctx || (n && (!d_data || !d_out))) return WS_ERR_INVALID_ARG;
  if (dtype_size(dtype) == 0) return fail(ctx, WS_ERR_INVALID_ARG, "unknown dtype");
  // lib.rs:1143-1144: assert!(MAX < NEVER_FILL); assert!(MAX > ALWAYS_FILL)
  if (max_value >= WS_NEVER_FILL) return fail(ctx, WS_ERR_MAX_TOO_HIGH, ws_status_str(WS_ERR_MAX_TOO_HIGH));
  if (max_value <= WS_ALWAYS_FILL) return fail(ctx, WS_ERR_MAX_TOO_LOW, ws_status_str(WS_ERR_MAX_TOO_LOW));
  if (n == 0) return WS_OK;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->d_pp_scratch) WS_CUDA(ctx, cudaMalloc(&ctx->d_pp_scratch, pre_processor_scratch_bytes() + 64));
  double* minmax = (double*)((char*)ctx->d_pp_scratch + pre_processor_scratch_bytes());
  WS_CUDA(ctx, launch_pre_processor((int)dtype, d_data, n, max_value, ctx->d_pp_scratch, minmax, d_out, ctx->stream));
  return WS_OK;
}

extern "C" ws_status ws_pre_processor(ws_ctx* ctx, ws_dtype dtype, const void* data, size_t n, uint8_t max_value,
                                      uint8_t* out) {
  if (!ctx || (n && (!data || !out))) return WS_ERR_INVALID_ARG;
  const size_t es = dtype_size(dtype);
  if (es == 0) return fail(ctx, WS_ERR_INVALID_ARG, "unknown dtype");
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  void* d_in = nullptr;
  uint8_t* d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_in, std::max<size_t>(n * es, 16));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, std::max<size_t>(n, 16));
  ws_status st = WS_OK;
  if (e != cudaSuccess) st = cuda_fail(ctx, e, "ws_pre_processor: cudaMalloc");
  if (st == WS_OK && n) {
    e = cudaMemcpyAsync(d_in, data, n * es, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) st = cuda_fail(ctx, e, "ws_pre_processor: upload");
  }
  if (st == WS_OK) st = ws_dev_pre_processor(ctx, dtype, d_in, n, max_value, d_out);
  if (st == WS_OK && n) {
    e = cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = cuda_fail(ctx, e, "ws_pre_processor: download");
  }
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_in);
  cudaFree(d_out);
  return st;
}

// find_local_minima of n_img dense slices already on the device (ctx->d_img) -> library-owned host list
static ws_status minima_from_device(ws_ctx* ctx, ws_plan* p, size_t n_img, uint64_t** out_rc, uint64_t* out_offsets) {
  cudaStream_t s = ctx->stream;
  WS_TRY(grow(ctx, ctx->d_seed_off, ctx->d_seed_off_cap, n_img + 1));
  size_t total = 0;
  WS_TRY(ws_plan_find_local_minima(p, ctx->d_img, nullptr, 0, ctx->d_seed_off, &total));
  WS_TRY(grow(ctx, ctx->d_seeds, ctx->d_seeds_cap, 2 * total));
  if (total) WS_CUDA(ctx, launch_minima_write(ctx->d_img, p->d, p->chunk_counts, ctx->d_seeds, (uint32_t)total, s));
  std::vector<uint32_t> h(2 * std::max<size_t>(total, 1)), hoff(n_img + 1);
  if (total) WS_TRY(copy_d2h(ctx, h.data(), ctx->d_seeds, 2 * total * 4));
  WS_CUDA(ctx, cudaMemcpyAsync(hoff.data(), ctx->d_seed_off, (n_img + 1) * 4, cudaMemcpyDeviceToHost, s));
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  uint64_t* out = (uint64_t*)malloc(std::max<size_t>(2 * total, 1) * sizeof(uint64_t));
  if (!out) return fail(ctx, WS_ERR_INTERNAL, "host allocation failed");
  for (size_t i = 0; i < 2 * total; ++i) out[i] = h[i];
  for (size_t b = 0; b <= n_img; ++b) out_offsets[b] = hoff[b];
  *out_rc = out;
  return WS_OK;
}

extern "C" ws_status ws_find_local_minima_batch(ws_ctx* ctx, const uint8_t* imgs, size_t n_img, size_t rows,
                                                size_t cols, uint64_t** out_rc, uint64_t* out_offsets) {
  if (!ctx || !imgs || !out_rc || !out_offsets) return WS_ERR_INVALID_ARG;
  *out_rc = nullptr;
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_img == 0 || rows == 0 || cols == 0) return fail(ctx, WS_ERR_INVALID_ARG, "empty image");
  ws_plan* p = nullptr;
  WS_TRY(get_plan(ctx, n_img, rows, cols, &p));
  const size_t npx = n_img * rows * cols;
  WS_TRY(grow(ctx, ctx->d_img, ctx->d_img_cap, npx));
  WS_TRY(copy_h2d(ctx, ctx->d_img, imgs, npx));
  return minima_from_device(ctx, p, n_img, out_rc, out_offsets);
}

extern "C" ws_status ws_find_local_minima(ws_ctx* ctx, const ws_image* img, uint64_t** out_rc, size_t* out_n) {
  if (!ctx || !out_rc || !out_n) return WS_ERR_INVALID_ARG;
  *out_rc = nullptr;
  *out_n = 0;
  WS_TRY(check_image(ctx, img));
  WS_CUDA(ctx, cudaSetDevice(ctx->device));
  ws_plan* p = nullptr;
  WS_TRY(get_plan(ctx, 1, img->rows, img->cols, &p));
  WS_TRY(grow(ctx, ctx->d_img, ctx->d_img_cap, img->rows * img->cols));
  WS_TRY(upload_image(ctx, img, ctx->d_img));   // any strides: dense on the device (the workers gather, hostpipe.h)
  uint64_t off[2] = {0, 0};
  WS_TRY(minima_from_device(ctx, p, 1, out_rc, off));
  *out_n = (size_t)off[1];
  return WS_OK;
}

extern "C" ws_status ws_transform(ws_ctx* ctx, const ws_config* cfg, const ws_image* img, const uint64_t* seeds_rc,
                                  size_t nseeds, uint64_t* out_labels) {
  if (!ctx || !out_labels) return WS_ERR_INVALID_ARG;
  WS_TRY(check_cfg(ctx, cfg));
  WS_TRY(check_image(ctx, img));
  if (cfg->kind == WS_MERGING) {
    // lib.rs:1524-1536: image and seeds are ignored, the shape is the INPUT shape (no padding)
    const size_t rows = img->rows, cols = img->cols;
    memset(out_labels, 0, rows * cols * sizeof(uint64_t));
    if (rows >= 3 && cols >= 3)
      for (size_t r = 1; r + 1 < rows; ++r)
        for (size_t c = 1; c + 1 < cols; ++c) out_labels[r * cols + c] = 123;
    return WS_OK;
  }
  HostRun hr;
  WS_TRY(host_run(ctx, cfg, img, seeds_rc, nseeds, &hr));
  return download_labels_u64(ctx, hr.plan, hr.plan->fb.lab, hr.npx, out_labels);
}

extern "C" ws_status ws_transform_compact(ws_ctx* ctx, const ws_config* cfg, const ws_image* img,
                                          const uint64_t* seeds_rc, size_t nseeds, uint32_t* out_labels,
                                          uint8_t* out_level) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  HostRun hr;
  WS_TRY(host_run(ctx, cfg, img, seeds_rc, nseeds, &hr));
  if (out_labels) {
    if (hr.npx * 4 >= ((size_t)1 << 20) && !host_ptr_is_pinned(out_labels)) {
      // pageable: the label words cross as they are, the workers drop the marker bit out of the ring
      WS_TRY(staged_d2h(ctx, hr.plan->fb.lab, hr.npx * 4, [=](const uint8_t* sp, size_t off, size_t len) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(sp);
        uint32_t* dst = out_labels + off / 4;
        for (size_t i = 0; i < len / 4; ++i) dst[i] = src[i] & LAB_MASK;
      }));
    } else {
      WS_TRY(grow(ctx, ctx->d_out[0], ctx->d_out_cap[0], (hr.npx + 1) / 2));
      WS_CUDA(ctx, launch_strip_labels(hr.plan->fb.lab, hr.npx, (uint32_t*)ctx->d_out[0], ctx->stream));
      WS_CUDA(ctx, cudaMemcpyAsync(out_labels, ctx->d_out[0], hr.npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
      hr.plan->stats[4] += 1;
    }
  }
  if (out_level) WS_TRY(copy_d2h(ctx, out_level, hr.plan->fb.lvl, hr.npx));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return WS_OK;
}

extern "C" ws_status ws_transform_history(ws_ctx* ctx, const ws_config* cfg, const ws_image* img,
                                          const uint64_t* seeds_rc, size_t nseeds, uint8_t* out_levels,
                                          uint64_t* out_labels) {
  if (!ctx || !out_labels) return WS_ERR_INVALID_ARG;
  HostRun hr;
  WS_TRY(host_run(ctx, cfg, img, seeds_rc, nseeds, &hr));
  if (out_levels)
    for (uint32_t l = 0; l <= cfg->max_water_level; ++l) out_levels[l] = (uint8_t)l;
  return stream_snapshots(ctx, hr, cfg, out_labels, [](uint8_t, const uint64_t*) {});
}

extern "C" ws_status ws_transform_with_hook(ws_ctx* ctx, const ws_config* cfg, const ws_image* img,
                                            const uint64_t* seeds_rc, size_t nseeds, ws_level_hook hook, void* user) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  HostRun hr;
  WS_TRY(host_run(ctx, cfg, img, seeds_rc, nseeds, &hr));
  if (!hook) return WS_OK;  // no hook: the reference returns an empty Vec (lib.rs:1510, 1520)
  // host copies of what HookCtx exposes: the (padded) image and the (colour, (row, col)) list
  std::vector<uint8_t> h_img(hr.npx);
  WS_CUDA(ctx, cudaMemcpyAsync(h_img.data(), ctx->d_img, hr.npx, cudaMemcpyDeviceToHost, ctx->stream));
  WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const bool auto_seeds = nseeds == WS_SEEDS_AUTO;
  nseeds = hr.nseeds;
  std::vector<uint64_t> h_seeds(3 * std::max<size_t>(nseeds, 1));
  std::vector<uint32_t> found;
  if (auto_seeds && nseeds) {  // the list found on the device
    found.resize(2 * nseeds);
    WS_CUDA(ctx, cudaMemcpyAsync(found.data(), ctx->d_seeds, 2 * nseeds * 4, cudaMemcpyDeviceToHost, ctx->stream));
    WS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  for (size_t i = 0; i < nseeds; ++i) {
    h_seeds[3 * i] = i + 1;
    h_seeds[3 * i + 1] = auto_seeds ? found[2 * i] : seeds_rc[2 * i];
    h_seeds[3 * i + 2] = auto_seeds ? found[2 * i + 1] : seeds_rc[2 * i + 1];
  }
  ws_hook_ctx hc;
  hc.max_water_level = cfg->max_water_level;
  hc.image = h_img.data();
  hc.rows = hr.orows;
  hc.cols = hr.ocols;
  hc.seeds = h_seeds.data();
  hc.nseeds = nseeds;
  return stream_snapshots(ctx, hr, cfg, nullptr, [&](uint8_t level, const uint64_t* colours) {
    hc.water_level = level;
    hc.colours = colours;
    hook(user, &hc);
  });
}

extern "C" ws_status ws_transform_lake_counts(ws_ctx* ctx, const ws_config* cfg, const ws_image* img,
                                              const uint64_t* seeds_rc, size_t nseeds, uint64_t* out_lake_counts,
                                              uint64_t* out_uncoloured) {
  if (!ctx) return WS_ERR_INVALID_ARG;
  HostRun hr;
  WS_TRY(host_run(ctx, cfg, img, seeds_rc, nseeds, &hr));
  ws_plan* p = hr.plan;
  cudaStream_t s = ctx->stream;
  const uint32_t nlev = (uint32_t)cfg->max_water_level + 1u;
  std::vector<uint32_t> h_hist(256), h_counts(256);
  // level histogram -> uncoloured pixels per level
  WS_CUDA(ctx, launch_level_hist(p->fb.lvl, p->d, p->lvl_hist, s));
  WS_CUDA(ctx, cudaMemcpyAsync(h_hist.data(), p->lvl_hist, 256 * 4, cudaMemcpyDeviceToHost, s));
  if (cfg->kind == WS_MERGING)
    WS_CUDA(ctx, cudaMemcpyAsync(h_counts.data(), p->mb.counts, 256 * 4, cudaMemcpyDeviceToHost, s));
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  p->stats[4] += 1;
  if (cfg->kind == WS_SEGMENTING) {
    // no merges: every colour present on the canvas (counted by label_tile) is a lake at every level
    std::vector<uint32_t> nd(1);
    WS_CUDA(ctx, cudaMemcpyAsync(nd.data(), p->mb.ndistinct, 4, cudaMemcpyDeviceToHost, s));
    WS_CUDA(ctx, cudaStreamSynchronize(s));
    for (uint32_t l = 0; l < nlev; ++l) h_counts[l] = nd[0];
  }
  uint64_t coloured = 0;
  for (uint32_t l = 0; l < nlev; ++l) {
    coloured += h_hist[l];
    if (out_lake_counts) out_lake_counts[l] = h_counts[l];
    if (out_uncoloured) out_uncoloured[l] = (uint64_t)hr.npx - coloured;
  }
  return WS_OK;
}

// find_lake_sizes (lib.rs:629-635) for every level, on the device: d_sizes[l][c], c = 0..nseeds (row length
// ncol = nseeds + 1; column 0 = uncoloured pixels).  Labels never exceed nseeds, so the rest of the reference's
// (rows*cols + 1)-long rows is zero.  The caller frees *d_sizes_out with cudaFree.
static ws_status lake_sizes_device(ws_ctx* ctx, ws_plan* p, const ws_config* cfg, size_t npx, size_t ncol,
                                   uint64_t** d_sizes_out) {
  cudaStream_t s = ctx->stream;
  const uint32_t nlev = (uint32_t)cfg->max_water_level + 1u;
  uint32_t* d_cnt = nullptr;
  uint64_t* d_sizes = nullptr;
  uint64_t* d_scratch = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_cnt, ncol * nlev * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_sizes, ncol * nlev * 8);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_scratch, ncol * 8);
  auto body = [&]() -> ws_status {
    WS_CUDA(ctx, e);
    WS_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, ncol * nlev * 4, s));
    WS_CUDA(ctx, cudaMemsetAsync(d_sizes, 0, ncol * nlev * 8, s));
    WS_CUDA(ctx, launch_colour_level_count(p->fb.lab, p->fb.lvl, npx, (uint32_t)ncol, d_cnt, s));
    WS_CUDA(ctx, launch_sizes_cumulate(d_cnt, (uint32_t)ncol, nlev, npx, d_sizes, s));
    p->stats[4] += 3;
    if (cfg->kind == WS_MERGING) {
      for (uint32_t l = 0; l < nlev; ++l) {
        WS_TRY(plan_rep_table(p, l));
        WS_CUDA(ctx, launch_sizes_fold(d_sizes + (size_t)l * ncol, p->rep, (uint32_t)ncol, d_scratch, s));
        p->stats[4] += 1;
      }
    }
    return WS_OK;
  };
  const ws_status st = body();
  cudaStreamSynchronize(s);
  cudaFree(d_cnt);
  cudaFree(d_scratch);
  if (st != WS_OK) {
    cudaFree(d_sizes);
    d_sizes = nullptr;
  }
  *d_sizes_out = d_sizes;
  return st;
}

static ws_status lake_sizes_guard(ws_ctx* ctx, const ws_config* cfg, size_t nseeds) {
  const double cells = (double)(nseeds + 1) * ((double)cfg->max_water_level + 1.0);
  if (cells * 12.0 > 64e9) return fail(ctx, WS_ERR_TOO_LARGE, "lake sizes: (nseeds + 1) x levels too large");
  return WS_OK;
}

extern "C" ws_status ws_transform_to_list(ws_ctx* ctx, const ws_config* cfg, const ws_image* img,
                                          const uint64_t* seeds_rc, size_t nseeds, uint8_t* out_levels,
                                          uint64_t* out_sizes) {
  if (!ctx || !out_sizes) return WS_ERR_INVALID_ARG;
  WS_TRY(check_cfg(ctx, cfg));
  if (nseeds == WS_SEEDS_AUTO) return fail(ctx, WS_ERR_INVALID_ARG, "lake sizes are indexed by the caller's seed list");
  WS_TRY(lake_sizes_guard(ctx, cfg, nseeds));  // before any work is done
  HostRun hr;
  WS_TRY(host_run(ctx, cfg, img, seeds_rc, nseeds, &hr));
  ws_plan* p = hr.plan;
  cudaStream_t s = ctx->stream;
  const uint32_t nlev = (uint32_t)cfg->max_water_level + 1u;
  const size_t ncol = nseeds + 1;
  uint64_t* d_sizes = nullptr;
  WS_TRY(lake_sizes_device(ctx, p, cfg, hr.npx, ncol, &d_sizes));
  // rows of length npx+1 on the host (find_lake_sizes, lib.rs:630): zero, then the first entries.  With more
  // seeds than pixels (duplicate positions are legal) the colours that survive still fit the row: a colour
  // > npx on the canvas would make the reference index out of bounds (lib.rs:633) -- clamp the copied width.
  const size_t row = hr.npx + 1;
  const size_t width = std::min(ncol, row);
  memset(out_sizes, 0, (size_t)nlev * row * sizeof(uint64_t));
  cudaError_t e = cudaMemcpy2DAsync(out_sizes, row * 8, d_sizes, ncol * 8, width * 8, nlev, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_sizes);
  WS_CUDA(ctx, e);
  if (out_levels)
    for (uint32_t l = 0; l < nlev; ++l) out_levels[l] = (uint8_t)l;
  return WS_OK;
}

extern "C" ws_status ws_transform_lake_sizes_compact(ws_ctx* ctx, const ws_config* cfg, const ws_image* img,
                                                     const uint64_t* seeds_rc, size_t nseeds,
                                                     uint64_t* out_lake_counts, uint64_t* out_sizes) {
  if (!ctx || !out_sizes) return WS_ERR_INVALID_ARG;
  WS_TRY(check_cfg(ctx, cfg));
  if (nseeds == WS_SEEDS_AUTO) return fail(ctx, WS_ERR_INVALID_ARG, "lake sizes are indexed by the caller's seed list");
  WS_TRY(lake_sizes_guard(ctx, cfg, nseeds));
  HostRun hr;
  WS_TRY(host_run(ctx, cfg, img, seeds_rc, nseeds, &hr));
  cudaStream_t s = ctx->stream;
  const uint32_t nlev = (uint32_t)cfg->max_water_level + 1u;
  const size_t ncol = nseeds + 1;
  uint64_t* d_sizes = nullptr;
  WS_TRY(lake_sizes_device(ctx, hr.plan, cfg, hr.npx, ncol, &d_sizes));
  cudaError_t e = cudaMemcpyAsync(out_sizes, d_sizes, (size_t)nlev * ncol * 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_sizes);
  WS_CUDA(ctx, e);
  if (out_lake_counts)  // lakes = colours with at least one pixel
    for (uint32_t l = 0; l < nlev; ++l) {
      uint64_t n = 0;
      const uint64_t* r = out_sizes + (size_t)l * ncol;
      for (size_t c = 1; c < ncol; ++c) n += (r[c] != 0);
      out_lake_counts[l] = n;
    }
  return WS_OK;
}

extern "C" ws_status ws_transform_batch(ws_ctx* ctx, const ws_config* cfg, const uint8_t* imgs, size_t n_img,
                                        size_t rows, size_t cols, const uint64_t* seeds_rc,
                                        const uint64_t* seed_offsets, uint64_t* out_labels,
                                        uint64_t* out_lake_counts) {
  if (!ctx || !imgs || !seed_offsets) return WS_ERR_INVALID_ARG;
  if (n_img == 0 || rows == 0 || cols == 0) return fail(ctx, WS_ERR_INVALID_ARG, "empty image");
  WS_TRY(check_cfg(ctx, cfg));
  if (out_lake_counts && cfg->kind != WS_MERGING)
    return fail(ctx, WS_ERR_INVALID_ARG, "lake counts need kind = WS_MERGING");
  // seed_offsets[0..n_img]: starts at 0, ascending, total below 2^31 (colours are 31-bit on the device)
  if (seed_offsets[0] != 0) return fail(ctx, WS_ERR_INVALID_ARG, "seed_offsets[0] must be 0");
  for (size_t b = 0; b < n_img; ++b)
    if (seed_offsets[b] > seed_offsets[b + 1]) return fail(ctx, WS_ERR_INVALID_ARG, "seed_offsets must be ascending");
  if (seed_offsets[n_img] >= 0x7fffffffull) return fail(ctx, WS_ERR_TOO_LARGE, "more than 2^31 - 2 seeds");
  const size_t nseeds = (size_t)seed_offsets[n_img];
  HostRun hr;
  WS_TRY(host_run_batch(ctx, cfg, imgs, nullptr, n_img, rows, cols, seeds_rc, seed_offsets, nseeds, &hr));
  ws_plan* p = hr.plan;
  cudaStream_t s = ctx->stream;
  const uint32_t nlev = (uint32_t)cfg->max_water_level + 1u;
  if (out_labels) WS_TRY(download_labels_u64(ctx, p, p->fb.lab, n_img * hr.npx, out_labels));
  if (out_lake_counts) {
    std::vector<uint32_t> h(n_img * 256);
    WS_CUDA(ctx, cudaMemcpyAsync(h.data(), p->mb.counts, n_img * 256 * 4, cudaMemcpyDeviceToHost, s));
    WS_CUDA(ctx, cudaStreamSynchronize(s));
    for (size_t b = 0; b < n_img; ++b)
      for (uint32_t l = 0; l < nlev; ++l) out_lake_counts[b * nlev + l] = h[b * 256 + l];
  }
  WS_CUDA(ctx, cudaStreamSynchronize(s));
  return WS_OK;
}
