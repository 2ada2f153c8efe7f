// common.cuh -- shared definitions of the sm_100a watershed kernels.
//
// Arrival-time encoding: every pixel carries one u32
//     T = (level << 24) | hop
// = the water level at which the reference's loop colours it (lib.rs:1379 /
// 1689) and the index of the synchronous flood pass inside that level
// (lib.rs:1394 / 1704 'colouring_loop).  Seeds hold T = 0, pixels that are
// never coloured hold T >= T_INF.  The reference's nested loops are the unique
// fixed point of
//     T(p) = max(A(p), 1 + min over 4-neighbours q of T(q)),   A(p) = (img[p] << 24) | 1
// for interior pixels with img[p] <= max_water_level (A(p) = T_INF otherwise),
// so any relaxation order reaches the same result.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ws {

constexpr uint32_t T_INF = 0xFF000000u;     // level 255, hop 0: "never coloured"
constexpr uint32_t HOP_MASK = 0x00FFFFFFu;
constexpr uint32_t LAB_RESOLVED = 0x80000000u;  // lab word holds a final label, not a pixel index
constexpr uint32_t LAB_MASK = 0x7FFFFFFFu;

// Flood tile: TILE_W x TILE_H pixels per CTA; 8 consumer warps + 1 producer warp.
constexpr int TILE_W = 64;
constexpr int TILE_H = 32;
constexpr int ROWS_PER_THREAD = 8;
constexpr int FLOOD_CONSUMERS = TILE_W * TILE_H / ROWS_PER_THREAD;  // 256 threads iterate on the tile
constexpr int FLOOD_THREADS = FLOOD_CONSUMERS + 32;                 // + the producer warp
constexpr int SM_W = TILE_W + 3;   // working tile, odd row stride: column-wise and row-wise accesses conflict-free
constexpr int SM_H = TILE_H + 2;
constexpr int PIX_W = TILE_W + 4;  // byte row stride of the working image tile (17 words, odd)
// Padded global layouts so that every tile's box (with halo) is an in-bounds, 16-byte aligned
// set of rows for the bulk-copy engine:
//   arrival times: image (r, c) -> Tp[(r + 1) * t_pitch + c + T_PAD_L], t_pitch = tiles_x * TILE_W + 8
//   image bytes:   image (r, c) -> pix[r * pix_pitch + c],              pix_pitch = tiles_x * TILE_W
constexpr int T_PAD_L = 4;
constexpr int STG_W = TILE_W + 8;  // words per staged row: image columns c0-4 .. c0+67
constexpr int STG_H = TILE_H + 2;

struct ImageDims {
  int n_img;      // slices in the batch
  int rows, cols; // per slice
  int tiles_x, tiles_y;
  // Row-strip decomposition of one large field (n_img == 1): this plan holds rows
  // [row_offset, row_offset + rows) of a field of global_rows rows; its first / last row is a halo
  // copy of a neighbouring strip's row when halo_top / halo_bottom is set.  Plain plans: 0, rows, 0, 0.
  int row_offset, global_rows, halo_top, halo_bottom;
  __host__ __device__ bool is_halo_row(int r) const { return (halo_top && r == 0) || (halo_bottom && r == rows - 1); }
  // window centre of the WHOLE field (lib.rs:220, 411-414), in local coordinates
  __host__ __device__ bool is_centre(int r, int c) const {
    const int gr = r + row_offset;
    return gr >= 1 && gr <= global_rows - 2 && c >= 1 && c <= cols - 2;
  }
  __host__ __device__ size_t px_per_img() const { return (size_t)rows * (size_t)cols; }
  __host__ __device__ size_t px_total() const { return px_per_img() * (size_t)n_img; }
  __host__ __device__ int tiles_per_img() const { return tiles_x * tiles_y; }
  __host__ __device__ int tiles_total() const { return tiles_per_img() * n_img; }
  // padded planes
  __host__ __device__ int t_pitch() const { return tiles_x * TILE_W + 8; }
  __host__ __device__ int t_rows() const { return tiles_y * TILE_H + 2; }
  __host__ __device__ size_t t_plane() const { return (size_t)t_pitch() * t_rows(); }
  __host__ __device__ size_t t_index(int r, int c) const { return (size_t)(r + 1) * t_pitch() + c + T_PAD_L; }
  __host__ __device__ int pix_pitch() const { return tiles_x * TILE_W; }
  __host__ __device__ int pix_rows() const { return tiles_y * TILE_H; }
  __host__ __device__ size_t pix_plane() const { return (size_t)pix_pitch() * pix_rows(); }
};

// SMs of the device the library runs on (streaming grids are sized in multiples of it); 148 on a B200, set from
// the device attributes when a context is created.
int num_sms();
void set_num_sms(int n);

__device__ __forceinline__ uint32_t ld_cg(const uint32_t* p) { return __ldcg(p); }
// Polling load: relaxed, GPU scope, with a memory clobber.  __ldcg is an `asm volatile` WITHOUT the
// clobber, and nvcc hoists it out of a spin loop as loop-invariant (seen in SASS: one LDG followed by a
// counting loop that never reloads).  Everything that waits for another CTA's write uses this one.
__device__ __forceinline__ uint32_t ld_poll(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_cg(uint32_t* p, uint32_t v) { __stcg(p, v); }

__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }

// ---------------------------------------------------------------------------
// small PTX helpers (mbarrier + bulk async copy)
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
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0u;
}
// global -> shared bulk copy (16-byte aligned, multiple of 16 bytes), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace ws
