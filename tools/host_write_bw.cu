// Host-memory write bandwidth of the box, in the patterns the host-window path uses (VERDICT r1 "next" 6):
//   (a) device->host DMA from G GPUs at once into pinned host memory (what the frame download is),
//   (b) streaming stores from T host threads (what the carry-over into the second ring is),
//   (c) both at once.
// nvcc -O2 -std=c++17 -o /tmp/host_write_bw tools/host_write_bw.cu && /tmp/host_write_bw > gpurun_out/host_write_bw.json
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Gpu { int dev; char* d; char* h; cudaStream_t s; };

static void nt_fill(char* dst, const char* src, size_t n) {
  for (size_t i = 0; i < n; i += 16) _mm_stream_si128((__m128i*)(dst + i), _mm_load_si128((const __m128i*)(src + i)));
  _mm_sfence();
}

int main() {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  const size_t bytes = 256ull << 20;
  const int reps = 8;
  std::vector<Gpu> g(ndev);
  for (int i = 0; i < ndev; ++i) {
    cudaSetDevice(i);
    g[i].dev = i;
    cudaMalloc(&g[i].d, bytes);
    cudaMemset(g[i].d, 1, bytes);
    cudaHostAlloc(&g[i].h, bytes, cudaHostAllocPortable);
    memset(g[i].h, 0, bytes);
    cudaStreamCreateWithFlags(&g[i].s, cudaStreamNonBlocking);
  }
  const int hw = (int)std::thread::hardware_concurrency();
  char *ca, *cb;
  posix_memalign((void**)&ca, 4096, bytes);
  posix_memalign((void**)&cb, 4096, bytes);
  memset(ca, 3, bytes); memset(cb, 4, bytes);
  auto dma = [&](int G) {          // aggregate GB/s of G concurrent D2H streams
    for (int i = 0; i < G; ++i) { cudaSetDevice(i); cudaMemcpyAsync(g[i].h, g[i].d, bytes, cudaMemcpyDeviceToHost, g[i].s); }
    for (int i = 0; i < G; ++i) { cudaSetDevice(i); cudaStreamSynchronize(g[i].s); }
    double t0 = now();
    for (int r = 0; r < reps; ++r)
      for (int i = 0; i < G; ++i) { cudaSetDevice(i); cudaMemcpyAsync(g[i].h, g[i].d, bytes, cudaMemcpyDeviceToHost, g[i].s); }
    for (int i = 0; i < G; ++i) { cudaSetDevice(i); cudaStreamSynchronize(g[i].s); }
    return (double)G * reps * bytes / (now() - t0) / 1e9;
  };
  auto cpu = [&](int T, std::atomic<bool>* stop, double* gbs) {   // aggregate GB/s of T threads of streaming copies
    std::vector<std::thread> th;
    std::vector<size_t> done(T, 0);
    double t0 = now();
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        const size_t chunk = bytes / T / 4096 * 4096;
        int it = 0;
        do { nt_fill(cb + t * chunk, ca + t * chunk, chunk); done[t] += chunk; ++it; } while (stop ? !stop->load() : it < reps);
      });
    for (auto& x : th) x.join();
    size_t tot = 0;
    for (size_t d : done) tot += d;
    *gbs = tot / (now() - t0) / 1e9;
  };
  // (d) would sending 48 of each frame's 60 bytes help (the goal columns are constant within an episode)? A 2-D copy of
  // 48-byte runs at a 60-byte pitch on both sides, one GPU: payload GB/s against the contiguous copy of whole rows
  double c2d_payload = 0.0, c1d = 0.0;
  {
    cudaSetDevice(0);
    const size_t rows = bytes / 60;
    for (int pass = 0; pass < 2; ++pass) {
      double t0 = now();
      for (int r = 0; r < reps; ++r) {
        if (pass == 0) cudaMemcpy2DAsync(g[0].h, 60, g[0].d, 60, 48, rows, cudaMemcpyDeviceToHost, g[0].s);
        else cudaMemcpyAsync(g[0].h, g[0].d, rows * 60, cudaMemcpyDeviceToHost, g[0].s);
      }
      cudaStreamSynchronize(g[0].s);
      const double dt = now() - t0;
      if (pass == 0) c2d_payload = (double)reps * rows * 48 / dt / 1e9; else c1d = (double)reps * rows * 60 / dt / 1e9;
    }
  }
  printf("{\"gpus\": %d, \"host_threads\": %d, \"buffer_mb\": %zu,\n \"frames_48_of_60_bytes_2d_copy\": {\"payload_gbs\": %.1f, \"rows_per_s\": %.3e, "
         "\"contiguous_60_byte_rows_gbs\": %.1f, \"contiguous_rows_per_s\": %.3e},\n \"dma_d2h_gbs\": {",
         ndev, hw, bytes >> 20, c2d_payload, c2d_payload * 1e9 / 48, c1d, c1d * 1e9 / 60);
  bool first = true;
  for (int G = 1; G <= ndev; G *= 2) { printf("%s\"%d\": %.1f", first ? "" : ", ", G, dma(G)); first = false; }
  printf("},\n \"cpu_streaming_copy_gbs\": {");
  first = true;
  for (int T : {1, 2, 4, 8, 16, 32}) {
    if (T > hw) break;
    double v; cpu(T, nullptr, &v);
    printf("%s\"%d\": %.1f", first ? "" : ", ", T, v); first = false;
  }
  printf("},\n \"dma_all_gpus_with_cpu_copies\": {");
  first = true;
  for (int T : {4, 8, 16}) {
    if (T > hw) break;
    std::atomic<bool> stop{false};
    double cpu_gbs = 0;
    std::thread c([&] { cpu(T, &stop, &cpu_gbs); });
    double d = dma(ndev);
    stop.store(true);
    c.join();
    printf("%s\"%d_threads\": {\"dma_gbs\": %.1f, \"cpu_write_gbs\": %.1f, \"total_write_gbs\": %.1f}", first ? "" : ", ", T, d, cpu_gbs, d + cpu_gbs);
    first = false;
  }
  printf("}}\n");
  return 0;
}
