// Host-side microbenchmark behind hostpipe.h's choice of store width: u32 label words -> usize labels with
// non-temporal stores of 16 / 32 / 64 bytes, T threads.   g++ -O2 -pthread -o widen_bench widen_bench.cpp
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void widen_sse2(uint64_t* dst, const uint32_t* src, size_t n) {
  const __m128i mask = _mm_set1_epi32(0x7FFFFFFF), zero = _mm_setzero_si128();
  for (size_t i = 0; i + 4 <= n; i += 4) {
    const __m128i v = _mm_and_si128(_mm_loadu_si128((const __m128i*)(src + i)), mask);
    _mm_stream_si128((__m128i*)(dst + i), _mm_unpacklo_epi32(v, zero));
    _mm_stream_si128((__m128i*)(dst + i + 2), _mm_unpackhi_epi32(v, zero));
  }
  _mm_sfence();
}
__attribute__((target("avx2"))) static void widen_avx2(uint64_t* dst, const uint32_t* src, size_t n) {
  const __m256i mask = _mm256_set1_epi32(0x7FFFFFFF);
  for (size_t i = 0; i + 8 <= n; i += 8) {
    const __m256i v = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(src + i)), mask);
    _mm256_stream_si256((__m256i*)(dst + i), _mm256_cvtepu32_epi64(_mm256_castsi256_si128(v)));
    _mm256_stream_si256((__m256i*)(dst + i + 4), _mm256_cvtepu32_epi64(_mm256_extracti128_si256(v, 1)));
  }
  _mm_sfence();
}
__attribute__((target("avx512f"))) static void widen_avx512(uint64_t* dst, const uint32_t* src, size_t n) {
  const __m256i mask = _mm256_set1_epi32(0x7FFFFFFF);
  for (size_t i = 0; i + 8 <= n; i += 8) {
    const __m256i v = _mm256_and_si256(_mm256_loadu_si256((const __m256i*)(src + i)), mask);
    _mm512_stream_si512((__m512i*)(dst + i), _mm512_cvtepu32_epi64(v));
  }
  _mm_sfence();
}
static void widen_plain(uint64_t* dst, const uint32_t* src, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = src[i] & 0x7FFFFFFFu;
}

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 16;
  const size_t n = (size_t)1 << 28;  // 268 M labels: 1 GB in, 2 GB out
  uint32_t* src = (uint32_t*)aligned_alloc(4096, n * 4);
  uint64_t* dst = (uint64_t*)aligned_alloc(4096, n * 8);
  memset(src, 1, n * 4);
  memset(dst, 0, n * 8);
  struct V { const char* name; void (*f)(uint64_t*, const uint32_t*, size_t); bool ok; };
  V vs[] = {{"plain stores", widen_plain, true},
            {"sse2 16 B nt", widen_sse2, true},
            {"avx2 32 B nt", widen_avx2, (bool)__builtin_cpu_supports("avx2")},
            {"avx512 64 B nt", widen_avx512, (bool)__builtin_cpu_supports("avx512f")}};
  for (auto& v : vs) {
    if (!v.ok) { printf("%-16s not supported\n", v.name); continue; }
    double best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      const size_t per = (n / T + 63) & ~(size_t)63;
      for (int t = 0; t < T; ++t) {
        const size_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
        th.emplace_back([=] { if (lo < hi) v.f(dst + lo, src + lo, hi - lo); });
      }
      for (auto& x : th) x.join();
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms < best) best = ms;
    }
    printf("%-16s %2d threads: %7.1f ms  (%.1f GB/s written, %.1f GB/s incl. the read)\n", v.name, T, best, n * 8 / best / 1e6, n * 12 / best / 1e6);
  }
  return 0;
}
