// Does pinning / memfd double mapping slow down CPU copies? nvcc -O2 -o /tmp/hcb2 tools/hostcopy_bench2.cu
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
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
static char* memfd_alias(size_t bytes, bool reg) {
  int fd = (int)syscall(SYS_memfd_create, "x", 0u);
  ftruncate(fd, bytes);
  char* span = (char*)mmap(nullptr, 2 * bytes, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  mmap(span, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED | MAP_POPULATE, fd, 0);
  mmap(span + bytes, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
  if (reg && cudaHostRegister(span, bytes, cudaHostRegisterPortable) != cudaSuccess) printf("register failed\n");
  return span;
}
static double run(char* b, const char* a, size_t bytes, int t, int mode) {
  double best = 1e9;
  for (int rep = 0; rep < 6; ++rep) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    size_t chunk = (bytes / t + 4095) / 4096 * 4096;
    for (int i = 0; i < t; ++i) th.emplace_back([=] {
      size_t o = i * chunk; if (o >= bytes) return; size_t n = std::min(chunk, bytes - o);
      if (mode == 0) memcpy(b + o, a + o, n); else nt16(b + o, a + o, n);
    });
    for (auto& x : th) x.join();
    best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  return best;
}
int main() {
  const size_t bytes = 60ull << 20;
  cudaFree(0);
  const char* names[4] = {"malloc", "memfd-alias", "memfd-alias+cudaHostRegister", "cudaHostAlloc"};
  for (int kind = 0; kind < 4; ++kind) {
    char *a, *b;
    if (kind == 0) { posix_memalign((void**)&a, 4096, bytes); posix_memalign((void**)&b, 4096, bytes); }
    else if (kind == 1) { a = memfd_alias(bytes, false); b = memfd_alias(bytes, false); }
    else if (kind == 2) { a = memfd_alias(bytes, true); b = memfd_alias(bytes, true); }
    else { cudaHostAlloc((void**)&a, bytes, cudaHostAllocPortable); cudaHostAlloc((void**)&b, bytes, cudaHostAllocPortable); }
    memset(a, 1, bytes); memset(b, 2, bytes);
    for (int t : {4, 8})
      for (int mode = 0; mode < 2; ++mode) {
        double s = run(b, a, bytes, t, mode);
        printf("%-30s %s threads %d: %.2f ms %.1f GB/s\n", names[kind], mode ? "nt-sse2" : "memcpy ", t, s * 1e3, bytes / s / 1e9);
      }
    // second mapping as the source (alias) for memfd kinds
    if (kind == 1 || kind == 2) { double s = run(b, a + bytes, bytes, 4, 1); printf("%-30s nt-sse2 threads 4 via alias src: %.2f ms\n", names[kind], s * 1e3); }
  }
  return 0;
}
