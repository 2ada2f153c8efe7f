// hostpipe.h -- host side of the reference-facing calls: moving PAGEABLE caller memory across the link.
//
// A Rust caller's `ArrayView2<u8>`, `&[(usize, usize)]` and `Array2<usize>` live in ordinary (pageable) memory.
// cudaMemcpy from / to such memory is staged by the driver through one internal buffer by one thread
// (~20 GB/s measured here, a third of the link).  This file is our own staging: a ring of page-locked slots
// and a small pool of worker threads that fill / drain a slot while the copy engine moves the previous one, and
// that do the format changes of the boundary on the way (usize pairs -> u32 pairs for the seeds, u32 label
// words -> usize labels for the result), so the link only carries the narrow forms.
// Plain C++ (no kernels); included by engine.cu only.
#pragma once

#include <cuda_runtime.h>
#include <immintrin.h>
#include <sched.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace ws {

// Persistent worker threads; run(n, f) calls f(0..n-1) on the workers and the caller and returns when all are done.
class HostPool {
 public:
  ~HostPool() { stop(); }
  int size() const { return (int)th_.size() + 1; }  // workers + the calling thread
  void start(int nthreads) {
    stop();
    quit_ = false;
    for (int i = 1; i < nthreads; ++i) th_.emplace_back([this] { worker(); });
  }
  void stop() {
    {
      std::lock_guard<std::mutex> g(m_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
    th_.clear();
  }
  void run(size_t n, const std::function<void(size_t)>& f) {
    if (n == 0) return;
    if (th_.empty() || n == 1) {
      for (size_t i = 0; i < n; ++i) f(i);
      return;
    }
    {
      std::lock_guard<std::mutex> g(m_);
      job_ = &f;
      njobs_ = n;
      left_.store(n, std::memory_order_relaxed);
      next_.store(0, std::memory_order_release);  // a worker that draws an index sees job_ / njobs_ of this job
      ++gen_;
    }
    cv_.notify_all();
    drain();
    std::unique_lock<std::mutex> g(m_);
    // (also wait for the workers to leave drain(): a late one must not meet the next job's counters)
    done_cv_.wait(g, [this] { return left_.load(std::memory_order_acquire) == 0 && active_ == 0; });
    job_ = nullptr;
  }

 private:
  void drain() {
    for (;;) {
      const size_t i = next_.fetch_add(1, std::memory_order_acq_rel);
      if (i >= njobs_) return;
      (*job_)(i);
      if (left_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> g(m_);
        done_cv_.notify_all();
      }
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return quit_ || gen_ != seen; });
        if (quit_) return;
        seen = gen_;
        ++active_;
      }
      drain();
      {
        std::lock_guard<std::mutex> g(m_);
        --active_;
      }
      done_cv_.notify_all();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(size_t)>* job_ = nullptr;
  size_t njobs_ = 0;
  std::atomic<size_t> next_{0}, left_{0};
  uint64_t gen_ = 0;
  int active_ = 0;
  bool quit_ = false;
};

// CPUs this process may run on (a rank of an 8-GPU job is usually confined to a share of the box).
inline int host_cpus_allowed() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int n = CPU_COUNT(&set);
    if (n > 0) return n;
  }
  const unsigned hw = std::thread::hardware_concurrency();
  return hw ? (int)hw : 1;
}

inline int host_threads_default() {
  if (const char* e = getenv("WS_HOST_THREADS")) {
    const int v = atoi(e);
    if (v >= 1) return std::min(v, 64);
  }
  return std::max(1, std::min(16, host_cpus_allowed()));
}

// true: page-locked (or managed / device-visible) memory the copy engine can address directly
inline bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes at;
  const cudaError_t e = cudaPointerGetAttributes(&at, p);
  cudaGetLastError();
  return e == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
}

// ---- the per-element format changes of the boundary --------------------------------------------------

// 32-byte non-temporal stores where the CPU has AVX2: 16 worker threads write 2.15 GB of labels in 18.3 ms instead
// of 21.5 ms with 16-byte stores on the pool's hosts (scripts/hostbench/widen_bench.cpp; 64-byte AVX-512 stores: 18.7)
__attribute__((target("avx2"))) inline size_t widen_labels_avx2(uint64_t* dst, const uint32_t* src, size_t n) {
  const __m256i mask = _mm256_set1_epi32(0x7FFFFFFF);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256i v = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)), mask);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), _mm256_cvtepu32_epi64(_mm256_castsi256_si128(v)));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 4), _mm256_cvtepu32_epi64(_mm256_extracti128_si256(v, 1)));
  }
  return i;
}

// n u32 label words (bit 31 = "resolved" marker of the device format) -> n usize labels
inline void widen_labels_host(uint64_t* dst, const uint32_t* src, size_t n) {
  static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("WS_HOST_NO_AVX2");   // (the switch: A/B timing)
  size_t i = 0;
  if (avx2) {
    while (i < n && ((uintptr_t)(dst + i) & 31u)) {
      dst[i] = src[i] & 0x7FFFFFFFu;
      ++i;
    }
    i += widen_labels_avx2(dst + i, src + i, n - i);
  }
  while (i < n && ((uintptr_t)(dst + i) & 15u)) {
    dst[i] = src[i] & 0x7FFFFFFFu;
    ++i;
  }
  const __m128i mask = _mm_set1_epi32(0x7FFFFFFF), zero = _mm_setzero_si128();
  // non-temporal stores: the destination is written once and not read here (no read-for-ownership traffic)
  for (; i + 4 <= n; i += 4) {
    const __m128i v = _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)), mask);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_unpacklo_epi32(v, zero));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 2), _mm_unpackhi_epi32(v, zero));
  }
  for (; i < n; ++i) dst[i] = src[i] & 0x7FFFFFFFu;
  _mm_sfence();
}

// n2 usize coordinates (row, col, row, col, ...) -> u32; anything outside the image becomes 0xFFFFFFFF so that
// seed_init flags it (the reference panics on an out-of-bounds seed, lib.rs:1366 / 1676).  `first` = index of
// src[0] in the whole list (parity decides row / column).
inline void narrow_seeds_host(uint32_t* dst, const uint64_t* src, size_t n2, size_t first, uint64_t rows, uint64_t cols) {
  // (a four-wide version with non-temporal stores into the ring slot measured the same on the pool's hosts: the
  // pass is bound by the 16 threads' reads of the caller's list)
  for (size_t i = 0; i < n2; ++i) {
    const uint64_t v = src[i];
    const uint64_t lim = ((first + i) & 1) ? cols : rows;
    dst[i] = v < lim ? (uint32_t)v : 0xFFFFFFFFu;
  }
}

}  // namespace ws
