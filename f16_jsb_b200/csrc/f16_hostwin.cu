// f16_hostwin.cu - host-resident observation windows (include/f16_hostwin.h).
//
// Host code (CUDA runtime calls for pinning and DMA; the only kernel is a one-thread publish of the done count): the step kernel in its frame layout
// (f16_b200.cu, OBS_FRAME) emits 60 B per env-step; this file lands those frames in a slot-major ring of pinned
// host memory whose pages are mapped twice back to back, so that the reference's (N,10,15) stacked observation
// (jsbsim_gym/jsbsim_gym.py:150,235,263) is a strided view of the ring and never has to be assembled.
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/f16_hostwin.h"

extern "C" int f16_internal_fail(const char* msg);
extern "C" int f16_internal_set_obs_frame(f16_handle h, float* obs_frame);
extern "C" int f16_internal_frame_buffers(f16_handle h, int64_t* n, int* device, float** obs_frame, float** reward, uint8_t** done,
                                          uint8_t** truncated, float** actions_stage);

// The done count goes to mapped host memory by a store from the device, not by a copy: a 4-byte cudaMemcpyAsync
// would queue behind the frame downloads of the earlier pieces on the copy engine, and the host wants the
// count as soon as the last kernel is done so that the fix-ups run under the remaining DMA.
__global__ void f16_publish_count_kernel(const int32_t* __restrict__ count_dev, volatile int32_t* count_host) { *count_host = *count_dev; }
extern "C" void f16_internal_count_launch(void);

namespace {

constexpr int SLOTS = F16_HOSTWIN_SLOTS;               // 11: the ten rows of a window + the slot being written
constexpr int ROWS = F16_OBS_FRAMES;                   // 10
constexpr int FEAT = F16_OBS_FEATURES;                 // 15
constexpr size_t ROW_BYTES = FEAT * sizeof(float);     // 60
constexpr size_t PAGE = 4096;

int failf(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return f16_internal_fail(buf);
}

// ---- NUMA locality. On a multi-socket host every rank's frames arrive by DMA from ITS GPU's PCIe root and are then
// touched by this rank's fix-up and carry-over threads; if the ring pages or those threads sit on the other socket,
// every byte crosses the inter-socket link twice (measured on the 8-GPU box: the carry-over copy took 10 ms per step
// instead of 2 ms on the one-GPU box). The ring is therefore first-touched, and the worker threads are kept, on the CPUs
// of the NUMA node the GPU hangs off (/sys/bus/pci/devices/<bus id>/numa_node). F16_HOSTWIN_NUMA=0 turns it off, =2
// also leaves the calling thread on that node. Only affinity calls are used (mbind / set_mempolicy are filtered in most
// containers): memory follows the first touch.
struct NumaPlace {
  bool valid = false;
  int node = -1;
  cpu_set_t cpus;
};
NumaPlace numa_of_device(int device) {
  NumaPlace p;
  CPU_ZERO(&p.cpus);
  const char* e = getenv("F16_HOSTWIN_NUMA");
  if (e && atoi(e) == 0) return p;
  char bus[64] = {0};
  if (device < 0 || cudaDeviceGetPCIBusId(bus, (int)sizeof(bus) - 1, device) != cudaSuccess) { cudaGetLastError(); return p; }
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  char path[256];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE* f = fopen(path, "r");
  if (!f) return p;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  if (node < 0) return p;                       // single-node machine or no affinity information
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return p;
  char list[4096] = {0};
  const size_t got = fread(list, 1, sizeof(list) - 1, f);
  fclose(f);
  list[got] = 0;
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return p;
  int n_set = 0;
  for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {   // "0-31,64-95"
    int a = 0, b = 0;
    const int k = sscanf(tok, "%d-%d", &a, &b);
    if (k < 1) continue;
    if (k == 1) b = a;
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &p.cpus); ++n_set; }
  }
  if (n_set == 0) return p;                     // the cgroup gives this process none of that node's CPUs
  p.valid = true;
  p.node = node;
  return p;
}

struct Ring {
  char* base = nullptr;      // 2 * bytes of address space
  size_t pitch = 0, bytes = 0;
  bool aliased = false;      // second half maps the same pages (else: a mirror that every write repeats)
  bool registered = false;   // cudaHostRegister-ed (aliased) / cudaHostAlloc-ed (mirror)
  bool pinned_alloc = false;
  int fd = -1;
  char* dev_base = nullptr;  // device address of `base` when the ring is mapped into the GPU's address space (zero-copy frames)
};

struct Fix {
  int32_t env;
  float frame[FEAT];
};

// A few persistent host threads: they run the fix-ups of finished envs while the frame DMA is in flight and
// carry the newest slot over to the second ring in the background between two steps.
class Pool {
 public:
  explicit Pool(int n, const NumaPlace* place = nullptr) {
    for (int i = 0; i < n; ++i) {
      th_.emplace_back([this, i] { worker(i); });
      if (place && place->valid) pthread_setaffinity_np(th_.back().native_handle(), sizeof(place->cpus), &place->cpus);
    }
  }
  ~Pool() {
    {
      std::unique_lock<std::mutex> lk(m_);
      cv_done_.wait(lk, [this] { return remaining_ == 0; });
      stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  // every worker runs job(worker index, worker count); returns at once
  void launch(std::function<void(int, int)> job) {
    wait();
    if (th_.empty()) { job(0, 1); return; }
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = std::move(job);
      remaining_ = (int)th_.size();
      ++gen_;
    }
    cv_work_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [this] { return remaining_ == 0; });
  }
  // f(begin, end) over [0, n), split over the workers when there is enough work to pay for the hand-off
  template <class F>
  void parallel_for(int64_t n, int64_t grain, F f) {
    wait();
    if (n <= 0) return;
    if (th_.empty() || n < 2 * grain) { f((int64_t)0, n); return; }
    launch([=](int i, int k) {
      const int64_t chunk = (n + k - 1) / k, b = i * chunk, e = std::min(n, b + chunk);
      if (b < e) f(b, e);
    });
    wait();
  }

 private:
  void worker(int i) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void(int, int)> job;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_work_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        job = job_;
      }
      job(i, (int)th_.size());
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--remaining_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_work_, cv_done_;
  std::function<void(int, int)> job_;
  uint64_t gen_ = 0;
  int remaining_ = 0;
  bool stop_ = false;
};

}  // namespace

struct f16_hostwin {
  int64_t n = 0;
  int n_rings = 1, flags = 0;
  bool pin = false;
  Ring ring[2];
  int head = SLOTS - 1;          // slot holding the newest frame
  int64_t t = 0;                 // steps since the last reset (selects the ring and the scalar buffers)
  float* reward[2] = {nullptr, nullptr};
  uint8_t* done[2] = {nullptr, nullptr};
  uint8_t* trunc[2] = {nullptr, nullptr};
  float* actions[2] = {nullptr, nullptr};
  f16_done_record* records = nullptr;    // mapped pinned memory the step kernel appends to (PIN only)
  int32_t* count_dev = nullptr;
  int32_t* count_host = nullptr;
  cudaEvent_t ev = nullptr, fork = nullptr;
  int n_chunks = 1;                      // pieces one step is pipelined in (upload | kernel | download)
  int zero_copy = 1;                     // 1: steps of up to 32 768 envs let the kernel write its frames straight into the ring; 2: every step; 0: never
  cudaStream_t cs[F16_HOSTWIN_MAX_CHUNKS] = {};
  cudaEvent_t kdone[F16_HOSTWIN_MAX_CHUNKS] = {};
  std::vector<float> term[2];
  std::vector<Fix> pending;              // finished envs of the previous step, still to be applied to the other ring
  int device = -1;
  int numa_node = -1;                    // node the ring was first-touched on and the workers are kept on (-1: no placement)
  double phase_s[F16_HOSTWIN_PHASES] = {0, 0, 0, 0, 0, 0, 0, 0};   // accumulated wall time per phase of f16_hostwin_step
  int64_t phase_steps = 0;
  std::atomic<int64_t> copy_ns{0};   // duration of the last carry-over, launch to last worker done
  Pool* pool = nullptr;      // fix-ups of finished envs
  Pool* copier = nullptr;    // carry-over of the newest slot to the other ring

  float* row(int r, int slot, int64_t env) const { return (float*)(ring[r].base + (size_t)slot * ring[r].pitch + (size_t)env * ROW_BYTES); }
  void write_row(int r, int slot, int64_t env, const float* f) const {
    memcpy(row(r, slot, env), f, ROW_BYTES);
    if (!ring[r].aliased) memcpy(row(r, slot + SLOTS, env), f, ROW_BYTES);
  }
};

namespace {

void free_ring(Ring& g) {
  if (!g.base) return;
  if (g.pinned_alloc) cudaFreeHost(g.base);
  else if (g.fd >= 0 || g.aliased) {
    if (g.registered) cudaHostUnregister(g.base);
    munmap(g.base, 2 * g.bytes);
  } else free(g.base);
  if (g.fd >= 0) close(g.fd);
  g = Ring();
}

// The ring's pages mapped twice back to back: [base, base+bytes) and [base+bytes, base+2*bytes).
bool make_aliased(Ring& g, bool pin) {
  int fd = (int)syscall(SYS_memfd_create, "f16_hostwin", 0u);
  if (fd < 0) return false;
  if (ftruncate(fd, (off_t)g.bytes) != 0) { close(fd); return false; }
  void* span = mmap(nullptr, 2 * g.bytes, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (span == MAP_FAILED) { close(fd); return false; }
  void* a = mmap(span, g.bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED | MAP_POPULATE, fd, 0);
  void* b = mmap((char*)span + g.bytes, g.bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
  if (a == MAP_FAILED || b == MAP_FAILED) { munmap(span, 2 * g.bytes); close(fd); return false; }
  if (pin) {
    // DMA only ever targets the first mapping; the second is read by the CPU alone
    if (cudaHostRegister(span, g.bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
      cudaGetLastError();
      munmap(span, 2 * g.bytes);
      close(fd);
      return false;
    }
    g.registered = true;
    void* dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, span, 0) == cudaSuccess) g.dev_base = (char*)dev;
    else cudaGetLastError();
  }
  g.base = (char*)span;
  g.fd = fd;
  g.aliased = true;
  return true;
}

int make_ring(Ring& g, int64_t n, bool pin, bool allow_alias) {
  g.pitch = ((size_t)n * ROW_BYTES + PAGE - 1) / PAGE * PAGE;
  g.bytes = g.pitch * SLOTS;
  if (allow_alias && make_aliased(g, pin)) return 0;
  // mirror fallback: 22 real slots
  if (pin) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, 2 * g.bytes, cudaHostAllocPortable) != cudaSuccess) return failf("f16_hostwin: cudaHostAlloc of %zu bytes failed", 2 * g.bytes);
    g.base = (char*)p;
    g.pinned_alloc = true;
  } else {
    void* p = nullptr;
    if (posix_memalign(&p, PAGE, 2 * g.bytes) != 0) return failf("f16_hostwin: out of host memory (%zu bytes)", 2 * g.bytes);
    g.base = (char*)p;
  }
  memset(g.base, 0, 2 * g.bytes);
  g.aliased = false;
  return 0;
}

template <class T>
int host_array(T** out, size_t count, bool pin, unsigned pin_flags = cudaHostAllocPortable) {
  void* p = nullptr;
  if (pin) {
    if (cudaHostAlloc(&p, count * sizeof(T), pin_flags) != cudaSuccess) return failf("f16_hostwin: cudaHostAlloc of %zu bytes failed", count * sizeof(T));
  } else {
    p = calloc(count, sizeof(T));
    if (!p) return failf("f16_hostwin: out of host memory");
  }
  *out = (T*)p;
  return 0;
}
template <class T>
void host_free(T* p, bool pin) {
  if (!p) return;
  if (pin) cudaFreeHost(p);
  else free(p);
}

inline int first_slot_of(int head) { return (head + 2) % SLOTS; }            // = head - 9 (mod 11)
inline int slot_back(int head, int k) { return (head - k + 2 * SLOTS) % SLOTS; }  // the slot written k steps ago

// The fix-ups of a step, in two phases. `early` runs once the done records are known, while the newest slot
// (`head`) of the returned ring is still receiving its DMA and - with two rings - slot head-1 may still be
// receiving the previous step's carry-over: it only touches slots head-9 .. head-2. `late` runs after both
// have landed and handles slot head-1.
void fix_early(f16_hostwin* w, int ring_now, const f16_done_record* recs, int64_t n_done, float* term) {
  const int head = w->head;
  // (1) finished envs of the previous step, for the ring that was not returned then: its window now is
  //     slots head-9 .. head; slot head-1 holds (or is about to receive) their reset frame, the rest is history
  if (!w->pending.empty()) {
    const Fix* P = w->pending.data();
    w->pool->parallel_for((int64_t)w->pending.size(), 512, [=](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i)
        for (int k = 2; k <= ROWS - 1; ++k) w->write_row(ring_now, slot_back(head, k), P[i].env, P[i].frame);
    });
    w->pending.clear();
  }
  // (2) envs that finished in this step: terminal stack = the nine previous rows + the terminal frame
  //     (dummy_vec_env.py:68), then those nine rows become copies of the reset frame (jsbsim_gym.py:325-329)
  if (n_done > 0) {
    w->pool->parallel_for(n_done, 512, [=](int64_t b, int64_t e) {
      for (int64_t j = b; j < e; ++j) {
        const f16_done_record& rc = recs[j];
        float* tj = term + (size_t)j * ROWS * FEAT;
        if (j + 2 < e)   // the rows of an env are a slot pitch apart: one TLB and one cache miss each
          for (int k = 2; k <= ROWS - 1; ++k) __builtin_prefetch(w->row(ring_now, slot_back(head, k), recs[j + 2].env), 1);
        for (int k = ROWS - 1; k >= 2; --k) {
          const int slot = slot_back(head, k);
          memcpy(tj + (size_t)(ROWS - 1 - k) * FEAT, w->row(ring_now, slot, rc.env), ROW_BYTES);
          w->write_row(ring_now, slot, rc.env, rc.reset_frame);
        }
        memcpy(tj + (size_t)(ROWS - 1) * FEAT, rc.terminal_frame, ROW_BYTES);
      }
    });
    if (w->n_rings == 2) {
      w->pending.resize((size_t)n_done);
      for (int64_t j = 0; j < n_done; ++j) {
        w->pending[(size_t)j].env = recs[j].env;
        memcpy(w->pending[(size_t)j].frame, recs[j].reset_frame, ROW_BYTES);
      }
    }
  }
}
void fix_late(f16_hostwin* w, int ring_now, const f16_done_record* recs, int64_t n_done, float* term) {
  const int slot = slot_back(w->head, 1);
  w->copier->wait();   // the previous step's carry-over into this ring (slot head-1) is complete from here on
  if (n_done <= 0) return;
  w->pool->parallel_for(n_done, 1024, [=](int64_t b, int64_t e) {
    for (int64_t j = b; j < e; ++j) {
      if (j + 4 < e) __builtin_prefetch(w->row(ring_now, slot, recs[j + 4].env), 1);
      memcpy(term + ((size_t)j * ROWS + (ROWS - 2)) * FEAT, w->row(ring_now, slot, recs[j].env), ROW_BYTES);
      w->write_row(ring_now, slot, recs[j].env, recs[j].reset_frame);
    }
  });
}

// streaming copy (non-temporal stores: the destination slot is not read again before the next step)
void stream_copy(char* dst, const char* src, size_t len) {
#if defined(__SSE2__)
  if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
    const __m128i* s = (const __m128i*)src;
    __m128i* d = (__m128i*)dst;
    const size_t n16 = len / 16;
    for (size_t i = 0; i < n16; ++i) _mm_stream_si128(d + i, _mm_load_si128(s + i));
    _mm_sfence();
    if (len & 15) memcpy(dst + n16 * 16, src + n16 * 16, len & 15);
    return;
  }
#endif
  memcpy(dst, src, len);
}

// Two rings: the frames of this step were DMA-ed into the returned ring only; host threads carry them over to
// the other ring in the background, between this step and the next (which returns that ring).
void carry_over(f16_hostwin* w, int ring_src) {
  if (w->n_rings != 2 || (w->flags & F16_HOSTWIN_DMA_BOTH)) return;
  const int ring_dst = 1 - ring_src;
  const size_t bytes = (size_t)w->n * ROW_BYTES;
  const char* src = (const char*)w->row(ring_src, w->head, 0);
  char* dst = (char*)w->row(ring_dst, w->head, 0);
  char* dst2 = w->ring[ring_dst].aliased ? nullptr : (char*)w->row(ring_dst, w->head + SLOTS, 0);
  if (bytes < ((size_t)1 << 16) || w->copier->size() == 0) {    // below 64 KB the copy is cheaper than waking a thread
    memcpy(dst, src, bytes);
    if (dst2) memcpy(dst2, src, bytes);
    return;
  }
  const auto t0 = std::chrono::steady_clock::now();
  w->copy_ns.store(0);
  w->copier->launch([=](int i, int k) {
    const size_t chunk = ((bytes + (size_t)k - 1) / (size_t)k + PAGE - 1) / PAGE * PAGE, b = (size_t)i * chunk;
    if (b >= bytes) return;
    const size_t len = std::min(chunk, bytes - b);
    stream_copy(dst + b, src + b, len);
    if (dst2) stream_copy(dst2 + b, src + b, len);
    const int64_t ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    int64_t seen = w->copy_ns.load();
    while (ns > seen && !w->copy_ns.compare_exchange_weak(seen, ns)) {}
  });
}

void fill_result(f16_hostwin* w, int ring_now, int cur, int64_t n_done, const f16_done_record* recs, f16_hostwin_result* out) {
  if (!out) return;
  out->ring = ring_now;
  out->first_slot = first_slot_of(w->head);
  out->n_done = n_done;
  out->reward = w->reward[cur];
  out->done = w->done[cur];
  out->truncated = w->trunc[cur];
  out->records = recs;
  out->terminal_obs = w->term[cur].data();
}

}  // namespace

extern "C" {

int f16_hostwin_create(f16_hostwin_handle* out, int64_t n_envs, int n_rings, int flags) {
  if (!out) return failf("f16_hostwin_create: out is NULL");
  *out = nullptr;
  if (n_envs <= 0) return failf("f16_hostwin_create: n_envs must be positive (got %lld)", (long long)n_envs);
  if (n_rings != 1 && n_rings != 2) return failf("f16_hostwin_create: n_rings must be 1 or 2 (got %d)", n_rings);
  f16_hostwin* w = new (std::nothrow) f16_hostwin;
  if (!w) return failf("out of host memory");
  w->n = n_envs;
  w->n_rings = n_rings;
  w->flags = flags;
  w->pin = (flags & F16_HOSTWIN_PIN) != 0;
  if (w->pin) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      delete w;
      return failf("f16_hostwin_create: F16_HOSTWIN_PIN needs a CUDA device; there is no CPU fallback for the env itself");
    }
  }
  // NUMA placement (see NumaPlace): the calling thread moves to the GPU's node while the buffers are first touched
  NumaPlace place;
  cpu_set_t caller_cpus;
  bool moved = false;
  if (w->pin) {
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess) place = numa_of_device(dev);
    else cudaGetLastError();
    if (place.valid && sched_getaffinity(0, sizeof(caller_cpus), &caller_cpus) == 0)
      moved = sched_setaffinity(0, sizeof(place.cpus), &place.cpus) == 0;
  }
  w->numa_node = place.valid ? place.node : -1;
  {
    // worker threads per pool: up to four (F16_HOSTWIN_THREADS overrides). Measured on the 8-GPU box (32 vCPUs, eight
    // ranks): 1 / 2 / 4 threads per pool give 6.9e8 / 8.0e8 / 8.1e8 env-steps/s end to end - the carry-over there is bound
    // by the host's memory system (~90 GB/s of DMA + copy writes over all ranks), not by thread count.
    const unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max(1u, std::min(4u, hw > 2 ? (hw - 1) / 2 : 1u));
    if (const char* e = getenv("F16_HOSTWIN_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 16) nt = v; }
    w->pool = new (std::nothrow) Pool(nt, &place);
    w->copier = new (std::nothrow) Pool(n_rings == 2 ? nt : 0, &place);
    if (!w->pool || !w->copier) { delete w->pool; delete w->copier; delete w; return failf("out of host memory"); }
  }
  int rc = 0;
  for (int r = 0; r < n_rings && !rc; ++r) rc = make_ring(w->ring[r], n_envs, w->pin, !(flags & F16_HOSTWIN_NO_ALIAS));
  for (int b = 0; b < 2 && !rc; ++b) {
    rc = host_array(&w->reward[b], (size_t)n_envs, w->pin);
    if (!rc) rc = host_array(&w->done[b], (size_t)n_envs, w->pin);
    if (!rc) rc = host_array(&w->trunc[b], (size_t)n_envs, w->pin);
    if (!rc) rc = host_array(&w->actions[b], (size_t)n_envs * F16_ACTION_DIM, w->pin);
  }
  if (!rc && w->pin) {
    rc = host_array(&w->records, (size_t)n_envs, true, cudaHostAllocPortable | cudaHostAllocMapped);
    if (!rc) rc = host_array(&w->count_host, 16, true, cudaHostAllocPortable | cudaHostAllocMapped);
    if (!rc && cudaMalloc(&w->count_dev, sizeof(int32_t)) != cudaSuccess) rc = failf("f16_hostwin_create: cudaMalloc failed");
    if (!rc && cudaEventCreateWithFlags(&w->ev, cudaEventDisableTiming) != cudaSuccess) rc = failf("f16_hostwin_create: cudaEventCreate failed");
    if (!rc && cudaEventCreateWithFlags(&w->fork, cudaEventDisableTiming) != cudaSuccess) rc = failf("f16_hostwin_create: cudaEventCreate failed");
    for (int c = 0; c < F16_HOSTWIN_MAX_CHUNKS && !rc; ++c) {
      if (cudaStreamCreateWithFlags(&w->cs[c], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&w->kdone[c], cudaEventDisableTiming) != cudaSuccess)
        rc = failf("f16_hostwin_create: stream / event creation failed");
    }
    // a step is worth cutting in pieces once a piece still fills the GPU for a wave or two
    w->n_chunks = n_envs >= 524288 ? 4 : n_envs >= 131072 ? 2 : 1;
    if (const char* e = getenv("F16_HOSTWIN_ZEROCOPY")) w->zero_copy = atoi(e);
    if (const char* e = getenv("F16_HOSTWIN_CHUNKS")) {
      const int c = atoi(e);
      if (c >= 1 && c <= F16_HOSTWIN_MAX_CHUNKS) w->n_chunks = c;
    }
  }
  {
    const char* e = getenv("F16_HOSTWIN_NUMA");
    if (moved && !(e && atoi(e) == 2)) sched_setaffinity(0, sizeof(caller_cpus), &caller_cpus);
  }
  if (rc) { f16_hostwin_destroy(w); return rc; }
  *out = w;
  return 0;
}

int f16_hostwin_numa_node(f16_hostwin_handle w) { return w ? w->numa_node : -1; }

int f16_hostwin_detach(f16_hostwin_handle w, f16_handle env) {
  // the env's done list lives in this window's mapped / device memory: un-point it before the window goes away
  if (!w || !env) return failf("f16_hostwin_detach: NULL argument");
  return f16_set_done_list(env, nullptr, nullptr);
}

int f16_hostwin_destroy(f16_hostwin_handle w) {
  if (!w) return 0;
  delete w->copier;    // waits for a carry-over in flight
  delete w->pool;
  w->pool = w->copier = nullptr;
  for (int r = 0; r < 2; ++r) free_ring(w->ring[r]);
  for (int b = 0; b < 2; ++b) {
    host_free(w->reward[b], w->pin);
    host_free(w->done[b], w->pin);
    host_free(w->trunc[b], w->pin);
    host_free(w->actions[b], w->pin);
  }
  host_free(w->records, true && w->pin);
  host_free(w->count_host, w->pin);
  if (w->count_dev) cudaFree(w->count_dev);
  if (w->ev) cudaEventDestroy(w->ev);
  if (w->fork) cudaEventDestroy(w->fork);
  for (int c = 0; c < F16_HOSTWIN_MAX_CHUNKS; ++c) {
    if (w->cs[c]) cudaStreamDestroy(w->cs[c]);
    if (w->kdone[c]) cudaEventDestroy(w->kdone[c]);
  }
  delete w;
  return 0;
}

int f16_hostwin_layout(f16_hostwin_handle w, int ring, float** base, int64_t* slot_pitch_bytes, int32_t* n_slots, int32_t* aliased) {
  if (!w) return failf("f16_hostwin_layout: NULL handle");
  if (ring < 0 || ring >= w->n_rings) return failf("f16_hostwin_layout: ring %d out of range", ring);
  if (base) *base = (float*)w->ring[ring].base;
  if (slot_pitch_bytes) *slot_pitch_bytes = (int64_t)w->ring[ring].pitch;
  if (n_slots) *n_slots = 2 * SLOTS;
  if (aliased) *aliased = w->ring[ring].aliased ? 1 : 0;
  return 0;
}

int f16_hostwin_timing(f16_hostwin_handle w, double* seconds_per_step, int reset) {
  if (!w) return failf("f16_hostwin_timing: NULL handle");
  if (seconds_per_step)
    for (int i = 0; i < F16_HOSTWIN_PHASES; ++i) seconds_per_step[i] = w->phase_steps ? w->phase_s[i] / (double)w->phase_steps : 0.0;
  if (reset) {
    for (int i = 0; i < F16_HOSTWIN_PHASES; ++i) w->phase_s[i] = 0.0;
    w->phase_steps = 0;
  }
  return 0;
}

float* f16_hostwin_action_buffer(f16_hostwin_handle w, int which) { return (w && (which == 0 || which == 1)) ? w->actions[which] : nullptr; }

// Contiguous copy of the current window: dst[n][k][f] = ring[ring][first_slot + k][n][f], on the window's worker threads.
// This is what copy_obs=True (DummyVecEnv's behaviour: the caller owns the array) costs - 600 B read + 600 B written
// per env; NumPy's strided copy of the same view runs single-threaded over 60-byte pieces and is ~10x slower.
int f16_hostwin_gather(f16_hostwin_handle w, int ring, int first_slot, float* dst) {
  if (!w || !dst) return failf("f16_hostwin_gather: NULL argument");
  if (ring < 0 || ring >= w->n_rings || first_slot < 0 || first_slot > SLOTS) return failf("f16_hostwin_gather: bad ring / slot");
  const f16_hostwin* cw = w;
  w->pool->parallel_for(w->n, 4096, [=](int64_t b, int64_t e) {
    for (int64_t n = b; n < e; ++n) {
      float* d = dst + n * (ROWS * FEAT);
      for (int k = 0; k < ROWS; ++k) memcpy(d + k * FEAT, cw->row(ring, first_slot + k, n), ROW_BYTES);
    }
  });
  return 0;
}

int f16_hostwin_fill(f16_hostwin_handle w, const float* frames, f16_hostwin_result* out) {
  if (!w || !frames) return failf("f16_hostwin_fill: NULL argument");
  w->copier->wait();
  for (int r = 0; r < w->n_rings; ++r)
    for (int s = 0; s < (w->ring[r].aliased ? SLOTS : 2 * SLOTS); ++s) memcpy(w->row(r, s, 0), frames, (size_t)w->n * ROW_BYTES);
  w->head = SLOTS - 1;
  w->t = 0;
  w->pending.clear();
  fill_result(w, 0, 0, 0, nullptr, out);
  return 0;
}

int f16_hostwin_push(f16_hostwin_handle w, const float* frames, const float* reward, const uint8_t* done, const uint8_t* truncated,
                     const f16_done_record* records, int64_t n_done, f16_hostwin_result* out) {
  if (!w || !frames) return failf("f16_hostwin_push: NULL argument");
  if (n_done < 0 || n_done > w->n || (n_done && !records)) return failf("f16_hostwin_push: bad done list");
  for (int64_t j = 0; j < n_done; ++j)
    if (records[j].env < 0 || records[j].env >= w->n) return failf("f16_hostwin_push: record %lld names env %d", (long long)j, records[j].env);
  w->t += 1;
  w->head = (w->head + 1) % SLOTS;
  const int ring_now = w->n_rings == 2 ? (int)(w->t & 1) : 0, cur = (int)(w->t & 1);
  for (int r = 0; r < w->n_rings; ++r) {
    if (r != ring_now && !(w->flags & F16_HOSTWIN_DMA_BOTH)) continue;      // carried over below
    memcpy(w->row(r, w->head, 0), frames, (size_t)w->n * ROW_BYTES);
    if (!w->ring[r].aliased) memcpy(w->row(r, w->head + SLOTS, 0), frames, (size_t)w->n * ROW_BYTES);
  }
  if (reward) memcpy(w->reward[cur], reward, (size_t)w->n * sizeof(float));
  if (done) memcpy(w->done[cur], done, (size_t)w->n);
  if (truncated) memcpy(w->trunc[cur], truncated, (size_t)w->n);
  w->term[cur].resize((size_t)n_done * ROWS * FEAT);
  fix_early(w, ring_now, records, n_done, w->term[cur].data());
  fix_late(w, ring_now, records, n_done, w->term[cur].data());
  carry_over(w, ring_now);
  fill_result(w, ring_now, cur, n_done, records, out);
  return 0;
}

#define CUDA_OK(call)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) return failf("%s failed: %s", #call, cudaGetErrorString(_e));       \
  } while (0)

static int env_buffers(f16_hostwin* w, f16_handle env, float** obs_frame, float** reward, uint8_t** done, uint8_t** trunc, float** act_stage) {
  if (!w) return failf("f16_hostwin: NULL handle");
  if (!w->pin) return failf("f16_hostwin: this window was created without F16_HOSTWIN_PIN; only the host-only entry points work");
  int64_t n = 0;
  int dev = 0;
  int rc = f16_internal_frame_buffers(env, &n, &dev, obs_frame, reward, done, trunc, act_stage);
  if (rc) return rc;
  if (n != w->n) return failf("f16_hostwin: the window holds %lld envs, the env handle %lld", (long long)w->n, (long long)n);
  CUDA_OK(cudaSetDevice(dev));
  w->device = dev;
  return 0;
}

int f16_hostwin_reset(f16_hostwin_handle w, f16_handle env, void* stream, f16_hostwin_result* out) {
  float *obs_frame, *reward, *act_stage;
  uint8_t *done, *trunc;
  int rc = env_buffers(w, env, &obs_frame, &reward, &done, &trunc, &act_stage);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)w->n * ROW_BYTES;
  w->copier->wait();
  for (int r = 0; r < w->n_rings; ++r)
    for (int s = 0; s < (w->ring[r].aliased ? SLOTS : 2 * SLOTS); ++s)
      CUDA_OK(cudaMemcpyAsync(w->row(r, s, 0), obs_frame, bytes, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  w->head = SLOTS - 1;
  w->t = 0;
  w->pending.clear();
  fill_result(w, 0, 0, 0, nullptr, out);
  return 0;
}

int f16_hostwin_step(f16_hostwin_handle w, f16_handle env, const float* actions_host, int auto_reset, void* stream,
                     f16_hostwin_result* out) {
  if (!actions_host) return failf("f16_hostwin_step: actions_host is NULL");
  using clk = std::chrono::steady_clock;
  auto t_prev = clk::now();
  int phase = 0;
  auto lap = [&]() {
    const auto now = clk::now();
    if (w) w->phase_s[phase] += std::chrono::duration<double>(now - t_prev).count();
    ++phase;
    t_prev = now;
  };
  float *obs_frame, *reward, *act_stage;
  uint8_t *done, *trunc;
  int rc = env_buffers(w, env, &obs_frame, &reward, &done, &trunc, &act_stage);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  f16_done_record* recs_dev = nullptr;
  CUDA_OK(cudaHostGetDevicePointer((void**)&recs_dev, w->records, 0));
  int32_t* count_host_dev = nullptr;
  CUDA_OK(cudaHostGetDevicePointer((void**)&count_host_dev, w->count_host, 0));
  rc = f16_set_done_list(env, recs_dev, w->count_dev);
  if (rc) return rc;
  // The step counter and the ring head advance only once every launch and copy of the step has been enqueued: an
  // early return (a failed launch) leaves the window exactly where it was.
  const int64_t t_new = w->t + 1;
  const int head_new = (w->head + 1) % SLOTS;
  const int ring_now = w->n_rings == 2 ? (int)(t_new & 1) : 0, cur = (int)(t_new & 1);
  // zero-copy steps point the env's frame output into the mapped ring; whatever happens it points back at the device
  // buffer when this function returns
  struct ObsFrameGuard {
    f16_handle env; float* device_frames; bool armed;
    ~ObsFrameGuard() { if (armed) f16_internal_set_obs_frame(env, device_frames); }
  } guard = {env, obs_frame, false};
  // One step in n_chunks pieces, each on its own stream: piece c's actions go up while piece c-1 computes and
  // piece c-2's frames come down (the three run on different engines). The done count is read once every
  // kernel has finished, so the host can start on the finished envs while the frames are still in flight.
  const int C = (int)std::min<int64_t>(w->n_chunks, (w->n + 127) / 128);
  const int64_t per = ((w->n + C - 1) / C + 127) / 128 * 128;
  rc = f16_step_begin(env, stream);
  if (rc) return rc;
  // Zero-copy frames: the ring is mapped into the GPU's address space, so the step kernel can store its newest frames
  // (one contiguous 1 920-byte span per warp) straight into this step's slot over PCIe while it runs, instead of writing
  // them to HBM and handing them to the copy engine afterwards. That removes the copy's launch latency and the
  // kernel -> copy serialisation, which is what a step of a few thousand envs consists of; large batches are bound by the
  // link either way and keep the copy engine (its 256-byte transactions use the link better than a warp's stores).
  const bool zc = w->pin && w->ring[ring_now].aliased && w->ring[ring_now].dev_base && !(w->flags & F16_HOSTWIN_DMA_BOTH) &&
                  (w->zero_copy == 2 || (w->zero_copy == 1 && C == 1 && w->n <= 32768));   // measured: +9 % at 4 096 envs, +4 % at 16 384, -2 % at 65 536, -9 % at 1M
  if (zc) {
    rc = f16_internal_set_obs_frame(env, (float*)(w->ring[ring_now].dev_base + (size_t)head_new * w->ring[ring_now].pitch));
    if (rc) return rc;
    guard.armed = true;
  }
  if (C > 1) CUDA_OK(cudaEventRecord(w->fork, st));
  for (int c = 0; c < C; ++c) {
    const int64_t first = (int64_t)c * per;
    if (first >= w->n) break;
    const size_t cnt = (size_t)std::min<int64_t>(per, w->n - first);
    const cudaStream_t s = C > 1 ? w->cs[c] : st;
    if (C > 1) CUDA_OK(cudaStreamWaitEvent(s, w->fork, 0));
    CUDA_OK(cudaMemcpyAsync(act_stage + first * F16_ACTION_DIM, actions_host + first * F16_ACTION_DIM, cnt * F16_ACTION_DIM * sizeof(float),
                            cudaMemcpyHostToDevice, s));
    rc = f16_step_range(env, act_stage, auto_reset, first, (int64_t)cnt, (void*)s);
    if (rc) return rc;
    if (C > 1) CUDA_OK(cudaEventRecord(w->kdone[c], s));
    else {
      f16_publish_count_kernel<<<1, 1, 0, st>>>(w->count_dev, count_host_dev);
      f16_internal_count_launch();
      CUDA_OK(cudaEventRecord(w->ev, st));
    }
    CUDA_OK(cudaMemcpyAsync(w->done[cur] + first, done + first, cnt, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(w->trunc[cur] + first, trunc + first, cnt, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(w->reward[cur] + first, reward + first, cnt * sizeof(float), cudaMemcpyDeviceToHost, s));
    for (int r = 0; r < w->n_rings; ++r) {
      const int rr = (r == 0) ? ring_now : 1 - ring_now;      // the returned ring first
      if (rr != ring_now && !(w->flags & F16_HOSTWIN_DMA_BOTH)) continue;     // carried over by host threads after the sync
      if (zc) continue;                                                        // the kernel wrote them there itself
      CUDA_OK(cudaMemcpyAsync(w->row(rr, head_new, first), obs_frame + first * FEAT, cnt * ROW_BYTES, cudaMemcpyDeviceToHost, s));
      if (!w->ring[rr].aliased)
        CUDA_OK(cudaMemcpyAsync(w->row(rr, head_new + SLOTS, first), obs_frame + first * FEAT, cnt * ROW_BYTES, cudaMemcpyDeviceToHost, s));
    }
  }
  if (C > 1) {
    for (int c = 0; c < C; ++c)
      if ((int64_t)c * per < w->n) CUDA_OK(cudaStreamWaitEvent(st, w->kdone[c], 0));
    f16_publish_count_kernel<<<1, 1, 0, st>>>(w->count_dev, count_host_dev);
    f16_internal_count_launch();
    CUDA_OK(cudaEventRecord(w->ev, st));
  }
  if (zc) {
    guard.armed = false;
    rc = f16_internal_set_obs_frame(env, obs_frame);      // launches have their arguments: back to the device buffer
    if (rc) return rc;
  }
  w->t = t_new;                             // everything is enqueued: the step has happened
  w->head = head_new;
  lap();                                    // [0] enqueue: copies and the kernel launch
  CUDA_OK(cudaEventSynchronize(w->ev));     // kernel finished: the records it wrote to mapped host memory are complete
  const int64_t n_done = *w->count_host;
  if (n_done < 0 || n_done > w->n) return failf("f16_hostwin_step: done count %lld out of range", (long long)n_done);
  // without auto-reset a finished env keeps its history (its newest row is the terminal frame itself)
  const int64_t n_fix = auto_reset ? n_done : 0;
  lap();                                    // [1] wait for the actions' upload and the kernel
  w->term[cur].resize((size_t)n_fix * ROWS * FEAT);
  fix_early(w, ring_now, w->records, n_fix, w->term[cur].data());
  lap();                                    // [2] fix-ups of slots head-9 .. head-2 (under the frame DMA)
  if (C > 1)
    for (int c = 0; c < C; ++c) CUDA_OK(cudaStreamSynchronize(w->cs[c]));
  CUDA_OK(cudaStreamSynchronize(st));
  lap();                                    // [3] wait for the rest of the device->host copies
  w->copier->wait();
  w->phase_s[7] += 1e-9 * (double)w->copy_ns.load();   // [7] how long that carry-over took on the copier threads (not part of the step's wall time)
  lap();                                    // [4] wait for the previous step's carry-over
  fix_late(w, ring_now, w->records, n_fix, w->term[cur].data());
  lap();                                    // [5] fix-ups of slot head-1
  carry_over(w, ring_now);
  lap();                                    // [6] hand the newest slot to the copier threads
  w->phase_steps += 1;
  fill_result(w, ring_now, cur, n_done, w->records, out);
  return 0;
}

}  // extern "C"
