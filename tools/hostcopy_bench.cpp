// Host memory copy bandwidth on the GPU box: glibc memcpy vs SSE2 / AVX-512 streaming stores, 1..16 threads,
// 63 MB slots (one host-window slot of 1M envs). g++ -O2 -pthread -mavx512f tools/hostcopy_bench.cpp
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static void nt16(char* d, const char* s, size_t n) {
  for (size_t i = 0; i < n; i += 16) _mm_stream_si128((__m128i*)(d + i), _mm_load_si128((const __m128i*)(s + i)));
  _mm_sfence();
}
__attribute__((target("avx512f"))) static void nt64(char* d, const char* s, size_t n) {
  for (size_t i = 0; i < n; i += 64) _mm512_stream_si512((__m512i*)(d + i), _mm512_load_si512((const void*)(s + i)));
  _mm_sfence();
}
int main() {
  const size_t bytes = 63ull << 20;
  char *a, *b;
  posix_memalign((void**)&a, 4096, bytes);
  posix_memalign((void**)&b, 4096, bytes);
  memset(a, 1, bytes); memset(b, 2, bytes);
  for (int mode = 0; mode < 3; ++mode)
    for (int t : {1, 2, 4, 8, 12, 16}) {
      double best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        size_t chunk = (bytes / t + 4095) / 4096 * 4096;
        for (int i = 0; i < t; ++i) th.emplace_back([=] {
          size_t o = i * chunk; if (o >= bytes) return; size_t n = std::min(chunk, bytes - o);
          if (mode == 0) memcpy(b + o, a + o, n); else if (mode == 1) nt16(b + o, a + o, n); else nt64(b + o, a + o, n);
        });
        for (auto& x : th) x.join();
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
      }
      printf("%s threads %2d: %.2f ms  %.1f GB/s copied\n", mode == 0 ? "memcpy" : mode == 1 ? "nt-sse2" : "nt-avx512", t, best * 1e3, bytes / best / 1e9);
    }
  return 0;
}
