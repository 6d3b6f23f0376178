// f16_b200.cu - kernels and C ABI of the batched F-16 environment (include/f16_b200.h).
//
// One thread = one environment. A step launch loads the env's structure-of-arrays state (coalesced,
// field-major), runs 4 FDM frames in registers (f16_model.cuh), builds the float32 observation frame
// exactly as jsbsim_gym/jsbsim_gym.py:172-197 does, evaluates reward / termination / truncation
// (jsbsim_gym.py:237-261, 487-509), optionally auto-resets (dummy_vec_env.py:63-72), stores the state
// back and then each warp cooperatively shifts its 32 envs' (10,15) observation stacks - one
// contiguous 19 200-byte span of the obs tensor - with fully coalesced 128-byte transactions.
// The aero/engine tables (6 kB float / 12 kB double) are staged once per CTA into shared memory.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/f16_b200.h"
#include "f16_env.cuh"
#include "f16_host_setup.h"

using namespace f16;

// ------------------------------------------------------------------------------------ state layout
// Tiles of 32 envs (one warp), field-major inside the tile:
//   tile t: [K fields: NKF x 32 doubles][R fields: NRF x 32 R][E fields: NEF x 32 4-byte words]     (field lists: f16_env.cuh)
// A warp's access to one field is one contiguous 128-byte (float) / 256-byte (double) row, exactly as with a global
// field-major array, but every field now sits at a COMPILE-TIME offset from the lane's tile pointer: the 122 loads and
// stores of a step take their address from one base register + an immediate instead of a 64-bit multiply-add each
// (round 1: 133 IADD3 + 34 IMAD around the 41 loads of the R fields alone), and a tile's state is one contiguous
// 9 216-byte (FP32 mode) / 14 592-byte (FP64 mode) span of HBM.
template <typename R>
struct TileLayout {
  static constexpr size_t K_OFF = 0;
  static constexpr size_t R_OFF = K_OFF + (size_t)NKF * 32 * sizeof(K);
  static constexpr size_t E_OFF = R_OFF + (size_t)NRF * 32 * sizeof(R);
  static constexpr size_t TILE_BYTES = E_OFF + (size_t)NEF * 32 * 4;
};
static_assert(TileLayout<float>::TILE_BYTES % 256 == 0 && TileLayout<double>::TILE_BYTES % 256 == 0, "tiles stay 256-byte aligned");
struct Layout {
  int64_t n, tiles;
  size_t tile_bytes, e_off, total;
};
static Layout make_layout(int64_t n, int mode) {
  Layout L;
  L.n = n;
  L.tiles = (n + 31) / 32;
  L.tile_bytes = mode == F16_MODE_FP64 ? TileLayout<double>::TILE_BYTES : TileLayout<float>::TILE_BYTES;
  L.e_off = mode == F16_MODE_FP64 ? TileLayout<double>::E_OFF : TileLayout<float>::E_OFF;
  L.total = (size_t)L.tiles * L.tile_bytes;
  return L;
}

// the lane's view of its tile: field f of the env is k[f * 32], r[f * 32], e[f * 32]
template <typename R>
struct StatePtrs {
  K* k;
  R* r;
  uint32_t* e;
};
template <typename R>
__device__ __forceinline__ StatePtrs<R> state_ptrs(void* state, int64_t env) {
  char* b = (char*)state + (size_t)(env >> 5) * TileLayout<R>::TILE_BYTES;
  const int lane = (int)(env & 31);
  StatePtrs<R> p;
  p.k = (K*)(b + TileLayout<R>::K_OFF) + lane;
  p.r = (R*)(b + TileLayout<R>::R_OFF) + lane;
  p.e = (uint32_t*)(b + TileLayout<R>::E_OFF) + lane;
  return p;
}

// state and observations are touched exactly once per step: streaming (evict-first) accesses keep them
// from displacing the prefetched observation rows in L2 (F16_STREAMING)
#ifndef F16_STREAMING
#define F16_STREAMING 0   // measured slower on B200 (extra register pressure), kept as an option
#endif
#if F16_STREAMING
#define F16_LD(ptr) __ldcs(ptr)
#define F16_ST(ptr, v) __stcs(ptr, v)
#else
#define F16_LD(ptr) (*(ptr))
#define F16_ST(ptr, v) (*(ptr) = (v))
#endif
template <typename R>
__device__ __forceinline__ void load_veh(Veh<R>& s, const StatePtrs<R>& p) {
  int f = 0;
#define X(m) s.m = F16_LD(&p.k[(f++) * 32]);
  F16_KFIELDS(X)
#undef X
  f = 0;
#define X(m) s.m = F16_LD(&p.r[(f++) * 32]);
  F16_RFIELDS(X)
#undef X
}
template <typename R>
__device__ __forceinline__ void store_veh(const Veh<R>& s, const StatePtrs<R>& p) {
  int f = 0;
#define X(m) F16_ST(&p.k[(f++) * 32], s.m);
  F16_KFIELDS(X)
#undef X
  f = 0;
#define X(m) F16_ST(&p.r[(f++) * 32], s.m);
  F16_RFIELDS(X)
#undef X
}
template <typename R>
__device__ __forceinline__ void load_env(EnvScalars& es, const StatePtrs<R>& p) {
  es.gx = __uint_as_float(p.e[EF_GOAL_X * 32]);
  es.gy = __uint_as_float(p.e[EF_GOAL_Y * 32]);
  es.gz = __uint_as_float(p.e[EF_GOAL_Z * 32]);
  es.last_d = __uint_as_float(p.e[EF_LAST_DIST * 32]);
  es.step = (int32_t)p.e[EF_STEP * 32];
  es.ep_ret = __uint_as_float(p.e[EF_EP_RET * 32]);
  es.ep_len = (int32_t)p.e[EF_EP_LEN * 32];
  es.episodes = p.e[EF_EPISODES * 32];
}
template <typename R>
__device__ __forceinline__ void store_env(const EnvScalars& es, const StatePtrs<R>& p) {
  p.e[EF_GOAL_X * 32] = __float_as_uint(es.gx);
  p.e[EF_GOAL_Y * 32] = __float_as_uint(es.gy);
  p.e[EF_GOAL_Z * 32] = __float_as_uint(es.gz);
  p.e[EF_LAST_DIST * 32] = __float_as_uint(es.last_d);
  p.e[EF_STEP * 32] = (uint32_t)es.step;
  p.e[EF_EP_RET * 32] = __float_as_uint(es.ep_ret);
  p.e[EF_EP_LEN * 32] = (uint32_t)es.ep_len;
  p.e[EF_EPISODES * 32] = es.episodes;
}

// ------------------------------------------------------------------------------------ constant memory
__constant__ MassSetT<double> c_msets[MS_COUNT];
__constant__ MassSetT<float> c_msets_f[MS_COUNT];
template <typename R> __device__ __forceinline__ const MassSetT<R>* msets_for();
template <> __device__ __forceinline__ const MassSetT<double>* msets_for<double>() { return c_msets; }
template <> __device__ __forceinline__ const MassSetT<float>* msets_for<float>() { return c_msets_f; }
__constant__ double c_snapshot[F16_NUM_STATE_FIELDS];   // canonical post-reset FDM state
__constant__ double c_snapshot_props[12];               // the 12 STATE_FORMAT properties after reset

// ------------------------------------------------------------------------------------ kernels
#ifndef F16_BLOCK
#define F16_BLOCK 128            // reset / pack kernels, and the step kernels unless overridden below
#endif
// Step-kernel launch shapes (profiles/r2_*: A/B runs at 1M envs, steady state).
#ifndef F16_BLOCK_F32
#define F16_BLOCK_F32 128
#endif
#ifndef F16_MIN_BLOCKS_F32
#define F16_MIN_BLOCKS_F32 4   // CTAs per SM of the float step kernel: 4 x 128 threads -> 128 registers
#endif
#ifndef F16_BLOCK_F64
#define F16_BLOCK_F64 128
#endif
#ifndef F16_MIN_BLOCKS_F64
#define F16_MIN_BLOCKS_F64 1   // 255 registers
#endif
// All warps of a CTA start every FDM frame together (one CTA barrier per frame): they then run the same ~25 KB of
// straight-line code within a few hundred cycles of each other and share their instruction-cache fills.
#ifndef F16_FRAME_SYNC_F32
#define F16_FRAME_SYNC_F32 1
#endif
#ifndef F16_FRAME_SYNC_F64
#define F16_FRAME_SYNC_F64 1
#endif
constexpr int BLOCK = F16_BLOCK;
template <typename R> struct StepShape;
template <> struct StepShape<float> { static constexpr int BLK = F16_BLOCK_F32, MINB = F16_MIN_BLOCKS_F32; static constexpr bool SYNC = F16_FRAME_SYNC_F32 != 0; };
template <> struct StepShape<double> { static constexpr int BLK = F16_BLOCK_F64, MINB = F16_MIN_BLOCKS_F64; static constexpr bool SYNC = F16_FRAME_SYNC_F64 != 0; };

template <typename R>
__device__ __forceinline__ void stage_tables(Tables<R>* dst, const Tables<R>* __restrict__ src) {
  static_assert(sizeof(Tables<R>) % 16 == 0, "table image must be a multiple of 16 bytes");
  const int4* s4 = reinterpret_cast<const int4*>(src);
  int4* d4 = reinterpret_cast<int4*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(Tables<R>) / 16); i += blockDim.x) d4[i] = __ldg(s4 + i);
  __syncthreads();
}

struct StepArgs {
  void* state;
  const void* tables;
  const float* actions;
  float* obs;
  float* reward;
  uint8_t* done;
  uint8_t* truncated;
  float* terminal_obs;
  float* ep_return;
  int32_t* ep_len;
  double* stats;
  int64_t n;
  uint64_t seed;
  int64_t env_id_base;
  uint32_t step_counter;
  int auto_reset;
  int ring_slot;     // ring layout only: slot (0..9) this step writes; the window is slots slot+1 .. slot+10
  size_t ring_pitch; // ring layout only: floats between consecutive slots (= N x 15)
  f16_done_record* done_list;   // frame layout only: one record per env that finished this step (may be mapped host memory)
  int32_t* done_count;          // frame layout only: device counter of appended records
  int64_t tile0;                // first 32-env tile of this launch (f16_step_range); n is the end of the range
  // near-ground tiles first (ground-reaction builds, whole-batch steps only; see f16_step_kernel). Three buffers rotate:
  // this step reads `cur`, fills `next` for the following step and clears `clr`, which the step before this one read.
  const int32_t* hot_list_cur; const int32_t* hot_count_cur; const uint8_t* hot_flag_cur;
  int32_t* hot_list_next; int32_t* hot_count_next; uint8_t* hot_flag_next;
  int32_t* hot_count_clr; uint8_t* hot_flag_clr;
  int hot_cap;                  // tiles the early CTAs can take (= early CTAs x warps per CTA); 0: scheduling off
};

// Canonical post-reset snapshot (f16_env.cuh: compute_snapshot), one thread, always in double.
__global__ void f16_init_snapshot_kernel(const Tables<double>* __restrict__ gT, const double* __restrict__ ic_state, double* out) {
  __shared__ __align__(16) Tables<double> T;
  stage_tables(&T, gT);
  if (threadIdx.x != 0) return;
  compute_snapshot(T, c_msets, ic_state, out);
}

// Warp-cooperative write of 32 envs' observation stacks: one contiguous span of 32 x 150 floats in
// the obs tensor. Per env (warp-uniform control flow) the warp moves the nine surviving rows down by
// one row in place - four full 128-byte transactions plus a 22-lane tail, all loads before all stores -
// and appends the newest frame from shared memory. Two envs are in flight per iteration (10 loads).
// flags: bit0 = env exists, bit1 = fill all ten rows with the frame (reset), bit2 = also emit the
// shifted stack to terminal_obs with `tframe_s[l]` as its newest row.
__device__ __forceinline__ void warp_write_obs(float* __restrict__ obs, float* __restrict__ term_obs, int64_t env0,
                                               const float (*frame_s)[16], const float (*tframe_s)[16],
                                               const uint8_t* flags_s) {
  const int lane = threadIdx.x & 31;
  constexpr int PER_ENV = F16_OBS_FRAMES * F16_OBS_FEATURES;   // 150 floats
  constexpr int SHIFT = F16_OBS_FEATURES;                      // 15
  constexpr int KEEP = PER_ENV - SHIFT;                        // 135 floats survive a step
  const uint8_t my = flags_s[lane];
  const unsigned m_active = __ballot_sync(0xffffffffu, my & 1);
  const unsigned m_reset = __ballot_sync(0xffffffffu, my & 2);
  const unsigned m_term = term_obs ? __ballot_sync(0xffffffffu, my & 4) : 0u;
  const int j4 = 128 + lane;                                   // tail chunk index (valid while < 150)
  int col[5];                                                  // column of element (c*32 + lane) within its row
#pragma unroll
  for (int c = 0; c < 5; ++c) col[c] = (c * 32 + lane) % SHIFT;
#ifndef F16_OBS_INFLIGHT
#define F16_OBS_INFLIGHT 8     // envs whose rows are in flight per iteration (5 loads each); 8 measured best
#endif
  constexpr int NU = F16_OBS_INFLIGHT;
  constexpr unsigned GROUP = (NU == 32) ? 0xffffffffu : ((1u << NU) - 1u);
  float* const base = obs + env0 * PER_ENV;
#pragma unroll 1
  for (int l0 = 0; l0 < 32; l0 += NU) {
    float* const gb = base + l0 * PER_ENV + lane;              // this lane's first element of the group's first env
    const unsigned act = (m_active >> l0) & GROUP, special = ((m_reset | m_term) >> l0) & GROUP;
    float v[NU][5];
    if (act == GROUP && special == 0u) {
      // common case, straight-line: every env of the group exists and just shifts by one row
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const float* eb = gb + u * PER_ENV + SHIFT;
#pragma unroll
        for (int c = 0; c < 4; ++c) v[u][c] = F16_LD(&eb[c * 32]);
        v[u][4] = (j4 < KEEP) ? F16_LD(&eb[128]) : frame_s[l0 + u][(j4 - KEEP) & 15];
      }
      __syncwarp();   // every lane's loads precede any lane's stores of the same env (in-place shift)
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float* eb = gb + u * PER_ENV;
#pragma unroll
        for (int c = 0; c < 4; ++c) F16_ST(&eb[c * 32], v[u][c]);
        if (j4 < PER_ENV) F16_ST(&eb[128], v[u][4]);
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      if (!((act >> u) & 1)) continue;
      const float* eb = gb + u * PER_ENV + SHIFT;
#pragma unroll
      for (int c = 0; c < 4; ++c) v[u][c] = F16_LD(&eb[c * 32]);
      v[u][4] = (j4 < KEEP) ? F16_LD(&eb[128]) : 0.0f;
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int l = l0 + u;
      if (!((act >> u) & 1)) continue;
      float* eb = gb + u * PER_ENV;
      if ((m_term >> l) & 1) {
        float* tb = term_obs + (env0 + l) * PER_ENV + lane;
#pragma unroll
        for (int c = 0; c < 4; ++c) tb[c * 32] = v[u][c];
        if (j4 < PER_ENV) tb[128] = (j4 < KEEP) ? v[u][4] : tframe_s[l][j4 - KEEP];
      }
      if ((m_reset >> l) & 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) eb[c * 32] = frame_s[l][col[c]];
        if (j4 < PER_ENV) eb[128] = frame_s[l][col[4]];
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) eb[c * 32] = v[u][c];
        if (j4 < PER_ENV) eb[128] = (j4 < KEEP) ? v[u][4] : frame_s[l][j4 - KEEP];
      }
    }
  }
}

// Ring layout of the observations (f16_bind_ring; F16BatchedEnv's default): 20 slots of one frame per env, slot-major
// ring[slot][env][15]. Each step writes the newest frame of every env into slot `slot` and into the mirror slot
// `slot + 10`, so the chronological ten-frame window is always the contiguous slot range slot+1 .. slot+10 and the
// stacked observation is the strided view obs[n][k][f] = ring[slot + 1 + k][n][f] - nothing is ever shifted or read:
// 120 B written per env-step instead of 540 B read + 600 B written. Slot-major makes each of the two writes what the
// frame layout's single write is: one contiguous, fully coalesced 1 920-byte span per warp (an env-major ring put
// 60-byte rows at a 1 200-byte pitch: partial 32-byte sectors that L2 has to read back from HBM before merging,
// measured 0.358 ms per step of 1M envs against 0.286 ms for the frame layout), and a consumer kernel that walks the
// window reads (N,15) planes that are contiguous over the envs. A reset fills all 20 slots of the env with the reset
// frame; a finished env's terminal stack is gathered from the nine older slots first.
__device__ __forceinline__ void warp_write_ring(float* __restrict__ ring, size_t slot_pitch, float* __restrict__ term_obs, int64_t env0,
                                                int64_t n, int slot, const float (*frame_s)[16], const float (*tframe_s)[16],
                                                const uint8_t* flags_s) {
  const int lane = threadIdx.x & 31;
  constexpr int ROW = F16_OBS_FEATURES, RING_ROWS = 2 * F16_OBS_FRAMES;   // 15, 20
  const uint8_t my = flags_s[lane];
  const unsigned m_active = __ballot_sync(0xffffffffu, my & 1);
  const unsigned m_reset = __ballot_sync(0xffffffffu, my & 2) & m_active;
  const unsigned m_term = term_obs ? (__ballot_sync(0xffffffffu, my & 4) & m_active) : 0u;
  float* const base = ring + env0 * ROW;                     // this warp's 32 x 15 floats inside slot 0
  // terminal stacks first (they need the window as it was before this step's frame): rare
  if (m_term) {
    for (int l = 0; l < 32; ++l) {
      if (!((m_term >> l) & 1)) continue;
      float* tb = term_obs + (env0 + l) * (F16_OBS_FRAMES * ROW);
      for (int j = lane; j < F16_OBS_FRAMES * ROW; j += 32) {
        const int k = j / ROW, c = j - k * ROW;                 // row k of the stack: slots slot+1 .. slot+9 are the nine older frames
        tb[j] = (k < 9) ? base[(size_t)(slot + 1 + k) * slot_pitch + l * ROW + c] : tframe_s[l][c];
      }
    }
    __syncwarp();
  }
  // newest frame (the reset frame for an env that auto-reset) into slots `slot` and `slot + 10`: two coalesced spans
  const int live = (int)((n - env0) < 32 ? (n - env0) : 32);
  float* const s0 = base + (size_t)slot * slot_pitch;
  float* const s1 = base + (size_t)(slot + F16_OBS_FRAMES) * slot_pitch;
  int l = 0, c = lane;                                    // element i = it*32 + lane -> (env l, column c)
  while (c >= ROW) { c -= ROW; ++l; }
#pragma unroll
  for (int it = 0; it < ROW; ++it) {
    if (l < live) {
      const float v = frame_s[l][c];
      s0[it * 32 + lane] = v;
      s1[it * 32 + lane] = v;
    }
    c += 2; l += 2;                                       // 32 = 2 * 15 + 2
    if (c >= ROW) { c -= ROW; ++l; }
  }
  // reset envs: the other 18 slots = reset frame as well (rare)
  if (m_reset) {
    for (int lr = 0; lr < 32; ++lr) {
      if (!((m_reset >> lr) & 1)) continue;
      for (int j = lane; j < RING_ROWS * ROW; j += 32) {
        const int r = j / ROW, cc = j - r * ROW;
        base[(size_t)r * slot_pitch + lr * ROW + cc] = frame_s[lr][cc];
      }
    }
  }
}

// Frame layout of the observations (opt-in, f16_bind_frames): the device keeps no history at all. Each step
// writes the newest frame of every env to a compact (N,15) tensor - one contiguous 1 920-byte span per warp,
// fully coalesced - which is exactly what crosses PCIe into the host-side window ring (f16_hostwin.h); an env
// that finished appends its terminal and reset frames to the done list instead.
__device__ __forceinline__ void warp_write_frames(float* __restrict__ frames, int64_t env0, int64_t n, const float (*frame_s)[16]) {
  const int lane = threadIdx.x & 31;
  constexpr int ROW = F16_OBS_FEATURES;
  float* const base = frames + env0 * ROW;
  const int live = (int)((n - env0) < 32 ? (n - env0) : 32);
  int l = 0, c = lane;                                    // element i = it*32 + lane -> (env l, column c)
  while (c >= ROW) { c -= ROW; ++l; }
#pragma unroll
  for (int it = 0; it < ROW; ++it) {
    if (l < live) base[it * 32 + lane] = frame_s[l][c];
    c += 2; l += 2;                                       // 32 = 2 * 15 + 2
    if (c >= ROW) { c -= ROW; ++l; }
  }
}

// ---- TMA (bulk async copy) staging of the table image: one elected thread issues a single
// cp.async.bulk global -> shared::cta that completes on an mbarrier; nobody spends registers or issue
// slots on the copy and the first tile's state loads overlap with it.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}

// The env-step of one env with ground reactions (env_step_one<R, true>, f16_ground.cuh), from the state in HBM
// back to HBM. Cold and deliberately not inlined: it is compiled with its own register allocation, so the
// contact and friction code (double precision, ~4 kB of stack) costs the hot path one compare per frame.
struct GroundStepOut { int flags; float reward, ep_ret; int32_t ep_len; };
template <typename R>
__device__ __noinline__ void env_step_ground(const StatePtrs<R> sp, const Tables<R>* T, float4 action, uint64_t seed,
                                             uint64_t gid, int auto_reset, float* frame16, float* tframe16, GroundStepOut* out) {
  Veh<R> s;
  EnvScalars es;
  load_veh(s, sp);
  load_env(es, sp);
  const float act[4] = {action.x, action.y, action.z, action.w};
  float reward = 0.0f, ep_ret = 0.0f;
  int32_t ep_len = 0;
  out->flags = env_step_one<R, GROUND_FULL>(s, es, *T, msets_for<R>(), c_msets, c_snapshot, c_snapshot_props, act, seed, gid, auto_reset,
                                     frame16, tframe16, &reward, &ep_ret, &ep_len);
  out->reward = reward; out->ep_ret = ep_ret; out->ep_len = ep_len;
  store_veh(s, sp);
  store_env(es, sp);
}

#ifndef F16_PREFETCH_OBS
#define F16_PREFETCH_OBS 1
#endif
// Step kernel: one warp = one tile of 32 envs, one CTA = StepShape<R>::BLK / 32 tiles; the table image is staged once
// per CTA by one bulk copy. Every thread of the CTA runs the four frames - lanes past the last env redo env n-1 without
// storing anything - so the per-frame CTA barrier (StepShape<R>::SYNC) is reached by all of them.
enum { OBS_STACKED = 0, OBS_RING = 1, OBS_FRAME = 2 };
template <typename R, int OBS, bool GROUND>
__global__ void __launch_bounds__(StepShape<R>::BLK, StepShape<R>::MINB) f16_step_kernel(const StepArgs a) {
  constexpr int WARPS = StepShape<R>::BLK / 32;
  constexpr bool SYNC = StepShape<R>::SYNC;
  __shared__ __align__(128) Tables<R> T;
  __shared__ __align__(16) float frame_s[WARPS][32][16];
  __shared__ __align__(16) float tframe_s[WARPS][32][16];
  __shared__ uint8_t flags_s[WARPS][32];
  __shared__ __align__(8) uint64_t tbar;
  static_assert(sizeof(Tables<R>) % 16 == 0, "bulk copy size must be a multiple of 16 bytes");

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (a.n + 31) >> 5;
  // ---- which tile does this warp step?
  // Ground-reaction builds, whole-batch steps: an env whose contact points reach the ground redoes its step alone in its
  // lane from a cold double-precision copy of the step (env_step_ground, ~80 of 1M envs per launch at the steady-state
  // crash rate of random actions). Where that happens in one of the grid's last CTAs it IS the tail of the launch:
  // measured 85 us of 425 (float) / 130 us of 1 080 (double) per step of 1M envs. An env can only touch within a step if it
  // starts it within one step's sink of the ground, so every warp that ends a step with such an env puts its tile on a list,
  // and the next launch begins with `hot_cap / WARPS` early CTAs that take the listed tiles; the regular CTA of a listed tile
  // skips it (its slot then runs one warp short, which is why the criterion is kept tight: ~5 % of the tiles).
  // The redo latency is then hidden behind the rest of the grid.
  const bool HOT = GROUND && a.hot_cap > 0;
  const int early_ctas = HOT ? a.hot_cap / WARPS : 0;
  int64_t tile;
  bool tile_ok;
  int n_act;                                                                   // warps of this CTA that step a tile (barrier size)
  if (HOT && (int)blockIdx.x < early_ctas) {
    const int hot_n = min(*a.hot_count_cur, a.hot_cap);
    const int first = (int)blockIdx.x * WARPS;
    if (first >= hot_n) return;                                                // (uniform over the CTA)
    n_act = min(WARPS, hot_n - first);
    tile_ok = warp < n_act;
    tile = tile_ok ? a.hot_list_cur[first + warp] : 0;
  } else {
    const int64_t base = a.tile0 + ((int64_t)blockIdx.x - early_ctas) * WARPS;
    const int64_t t = base + lane;                                             // lanes 0 .. WARPS-1 look at the CTA's tiles
    const bool mine = lane < WARPS && t < n_tiles && !(HOT && a.hot_flag_cur[t]);
    const unsigned act = __ballot_sync(0xffffffffu, mine);
    n_act = __popc(act);
    tile = base + warp;
    tile_ok = (act >> warp) & 1;
    if (HOT && lane == 0 && tile < n_tiles) a.hot_flag_clr[tile] = 0;          // every tile has exactly one regular warp
    if (HOT && blockIdx.x == (unsigned)early_ctas && threadIdx.x == 0) *a.hot_count_clr = 0;
    if (n_act == 0) return;                                                    // (uniform over the CTA)
  }
  // table image: one bulk copy per CTA, issued before any warp leaves and only by CTAs that have work (a CTA must not
  // exit with a copy into its shared memory in flight)
  if (threadIdx.x == 0) mbar_init(&tbar, 1);
  __syncthreads();
  if (threadIdx.x == 0) tma_load_1d(&T, a.tables, (uint32_t)sizeof(Tables<R>), &tbar);
  if (!tile_ok) return;
  const int64_t env0 = tile << 5;
  const bool valid = env0 + lane < a.n;
  const int64_t e = valid ? env0 + lane : a.n - 1;                             // lanes past the last env redo env n-1, storing nothing
  int flags = 0;
  bool near_ground = false;
  float sink_fps = 0.0f;           // radial (vertical) speed after the step, ft/s, positive down
  {
    const StatePtrs<R> sp = state_ptrs<R>(a.state, e);
    Veh<R> s;
    EnvScalars es;
    load_veh(s, sp);
    load_env(es, sp);
    const uint64_t gid = (uint64_t)(a.env_id_base + e);
    float act[4];
    if (a.actions) {
      float4 v = reinterpret_cast<const float4*>(a.actions)[e];
      act[0] = v.x; act[1] = v.y; act[2] = v.z; act[3] = v.w;
    } else {
      sample_action(a.seed, gid, a.step_counter, act);
    }
    mbar_wait(&tbar, 0);
    float reward, ep_ret = 0.0f;
    int32_t ep_len = 0;
    // each lane prefetches lines lane, lane+32, ... of the warp's 150-line observation span (F16_PREFETCH_OBS)
    PrefetchHint pf = {nullptr, 0, 0};
#if F16_PREFETCH_OBS == 1
    if (OBS == OBS_STACKED) {
      pf.ptr = reinterpret_cast<const char*>(a.obs + env0 * (F16_OBS_FRAMES * F16_OBS_FEATURES)) + lane * 128;
      pf.count = lane < 22 ? 5 : 4;       // 150 lines of 128 bytes
      pf.stride = 32 * 128;
    }
#endif
    flags = env_step_one<R, GROUND ? GROUND_DETECT : GROUND_OFF, SYNC>(s, es, T, msets_for<R>(), c_msets, c_snapshot, c_snapshot_props, act, a.seed, gid, a.auto_reset,
                                   frame_s[warp][lane], tframe_s[warp][lane], &reward, &ep_ret, &ep_len, pf, n_act * 32);
    if (!valid) {
      flags = 0;
    } else {
#ifndef F16_T_SKIP_REDO
#define F16_T_SKIP_REDO 0          // timing experiment only: drop the cold redo (results are then wrong for touching envs)
#endif
      if (GROUND && (flags & STEP_NEAR_GROUND) && !F16_T_SKIP_REDO) {
        // a contact point reached the ground (last env-step of a crash): redo the step with the contact forces
        // from the state still in HBM; the cold copy stores the new state itself
        GroundStepOut go;
        if (a.stats) atomicAdd(a.stats + 7, 1.0);
        env_step_ground<R>(sp, &T, make_float4(act[0], act[1], act[2], act[3]), a.seed, gid, a.auto_reset,
                           frame_s[warp][lane], tframe_s[warp][lane], &go);
        flags = go.flags; reward = go.reward; ep_ret = go.ep_ret; ep_len = go.ep_len;
      } else {
        store_veh(s, sp);
        store_env(es, sp);
        // d|r|/dt = r.v / |r| (the earth-rotation part of the velocity is perpendicular to r); |r| = a to 0.4 %
        sink_fps = -(float)((s.ri[0] * s.vi[0] + s.ri[1] * s.vi[1] + s.ri[2] * s.vi[2]) * (1.0 / kEarthA));
      }
      a.reward[e] = reward;
      a.done[e] = (flags & STEP_DONE) ? 1 : 0;
      a.truncated[e] = (flags & STEP_TRUNCATED) ? 1 : 0;
      if (flags & STEP_DONE) {
        if (a.ep_return) a.ep_return[e] = ep_ret;
        if (a.ep_len) a.ep_len[e] = ep_len;
        if (a.stats) {
          atomicAdd(a.stats + 0, 1.0);
          atomicAdd(a.stats + 1, (double)ep_ret);
          atomicAdd(a.stats + 2, (double)ep_len);
          if (flags & STEP_CRASH) atomicAdd(a.stats + 3, 1.0);
          if (flags & STEP_GOAL) atomicAdd(a.stats + 4, 1.0);
          if (flags & STEP_TRUNCATED) atomicAdd(a.stats + 5, 1.0);
        }
      }
      if (OBS == OBS_FRAME && (flags & STEP_DONE) && a.done_list) {
        // done list: terminal frame (the env's newest row when it finished) and the frame the next episode starts from
        f16_done_record* rec = a.done_list + atomicAdd(a.done_count, 1);
        const float* fr = frame_s[warp][lane];
        const float* tf = (flags & STEP_TERMINAL) ? tframe_s[warp][lane] : fr;
        uint4* w = reinterpret_cast<uint4*>(rec);
        w[0] = make_uint4((uint32_t)e, (uint32_t)(((flags & STEP_TRUNCATED) ? 1 : 0) | ((flags & STEP_CRASH) ? 2 : 0) | ((flags & STEP_GOAL) ? 4 : 0)),
                          __float_as_uint(ep_ret), (uint32_t)ep_len);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t4 = *reinterpret_cast<const float4*>(tf + 4 * q), r4 = *reinterpret_cast<const float4*>(fr + 4 * q);
          w[1 + q] = make_uint4(__float_as_uint(t4.x), __float_as_uint(t4.y), __float_as_uint(t4.z), __float_as_uint(t4.w));
          w[5 + q] = make_uint4(__float_as_uint(r4.x), __float_as_uint(r4.y), __float_as_uint(r4.z), __float_as_uint(r4.w));
        }
      }
    }
  }
  if (HOT) {
    // may this env touch the ground within the NEXT step? newest frame's altitude (metres; the reset frame's 1 524 m for
    // an env that auto-reset) minus what it sinks in one env-step (1/30 s, 25 % margin) against the reach of the contact
    // points (24.5 ft) + the redo margin
    near_ground = valid && frame_s[warp][lane][2] * 3.2808399f - fmaxf(sink_fps, 0.0f) * (1.25f / 30.0f) < 40.0f;
    if (__any_sync(0xffffffffu, near_ground) && lane == 0) {
      const int idx = atomicAdd(a.hot_count_next, 1);
      if (idx < a.hot_cap) { a.hot_list_next[idx] = (int32_t)tile; a.hot_flag_next[tile] = 1; }
    }
  }
  flags_s[warp][lane] = (uint8_t)(flags & (STEP_ACTIVE | STEP_RESET | STEP_TERMINAL));
  __syncwarp();
  if (OBS == OBS_FRAME) warp_write_frames(a.obs, env0, a.n, frame_s[warp]);
  else if (OBS == OBS_RING) warp_write_ring(a.obs, a.ring_pitch, a.terminal_obs, env0, a.n, a.ring_slot, frame_s[warp], tframe_s[warp], flags_s[warp]);
  else warp_write_obs(a.obs, a.terminal_obs, env0, frame_s[warp], tframe_s[warp], flags_s[warp]);
}

struct ResetArgs {
  void* state;
  const uint8_t* mask;
  const float* goals;
  float* obs;
  int64_t n;
  uint64_t seed;
  int64_t env_id_base;
  int obs_rows;      // 10 (stacked layout), 20 (ring layout) or 1 (frame layout)
  size_t env_pitch, row_pitch;   // floats between consecutive envs / rows of one env (stacked: 150, 15; ring: 15, N x 15)
  const void* tables;            // carry-over reset only: the table image of the context's precision
  const float* last_actions;     // carry-over reset only: N x 4, the action of each env's last step (or NULL)
};

template <typename R, bool CARRY>
__global__ void __launch_bounds__(BLOCK) f16_reset_kernel(const ResetArgs a) {
  __shared__ __align__(16) unsigned char Tbuf[CARRY ? sizeof(Tables<R>) : 16];   // only the carry-over reset runs FDM frames
  Tables<R>* const T = reinterpret_cast<Tables<R>*>(Tbuf);
  if (CARRY) stage_tables(T, reinterpret_cast<const Tables<R>*>(a.tables));
  const int64_t e = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (e >= a.n) return;
  if (a.mask && !a.mask[e]) return;
  StatePtrs<R> sp = state_ptrs<R>(a.state, e);
  Veh<R> s;
  EnvScalars es;
  load_env(es, sp);
  float g[3];
  if (a.goals) { g[0] = a.goals[e * 3 + 0]; g[1] = a.goals[e * 3 + 1]; g[2] = a.goals[e * 3 + 2]; }
  else sample_goal(a.seed, (uint64_t)(a.env_id_base + e), (es.episodes & ~kEpisodeUsedBit) + 1, g);
  float fr[16];
  if (CARRY) {
    // JSBSimEnv.reset on an env object that already exists: run_ic() + set-running on top of whatever the last
    // episode left behind (env_carryover_reset_one, f16_env.cuh)
    load_veh(s, sp);
    float la[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (a.last_actions) { const float4 v = reinterpret_cast<const float4*>(a.last_actions)[e]; la[0] = v.x; la[1] = v.y; la[2] = v.z; la[3] = v.w; }
    env_carryover_reset_one<R>(s, es, *T, msets_for<R>(), c_snapshot, c_snapshot_props, g, la, fr);
  } else {
    es.episodes = (es.episodes & ~kEpisodeUsedBit) + 1;
    env_reset_one<R>(s, es, c_snapshot, c_snapshot_props, g, fr);
  }
  store_veh(s, sp);
  store_env(es, sp);
  float* ob = a.obs + (size_t)e * a.env_pitch;
  for (int r = 0; r < a.obs_rows; ++r)
    for (int c = 0; c < F16_OBS_FEATURES; ++c) ob[(size_t)r * a.row_pitch + c] = fr[c];
}

template <typename R>
__global__ void f16_pack_kernel(void* state, int64_t n, int64_t first, double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  StatePtrs<R> sp = state_ptrs<R>(state, first + i);
  Veh<R> s;
  load_veh(s, sp);
  double a[F16_NUM_STATE_FIELDS];
  veh_to_packed(s, a);
  for (int f = 0; f < F16_NUM_STATE_FIELDS; ++f) out[i * F16_NUM_STATE_FIELDS + f] = a[f];
}
template <typename R>
__global__ void f16_unpack_kernel(void* state, int64_t n, int64_t first, const double* in) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  StatePtrs<R> sp = state_ptrs<R>(state, first + i);
  double a[F16_NUM_STATE_FIELDS];
  for (int f = 0; f < F16_NUM_STATE_FIELDS; ++f) a[f] = in[i * F16_NUM_STATE_FIELDS + f];
  Veh<R> s;
  veh_from_packed(s, a);
  store_veh(s, sp);
}

// ==================================================================================== host side
namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};   // all handles and host threads share it

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return -1;
}
#define CUDA_OK(call)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess) return fail("%s failed: %s", #call, cudaGetErrorString(_e));      \
  } while (0)

}  // namespace

struct f16_ctx {
  int device = 0, mode = 0;
  Layout L;
  void* tables_dev = nullptr;        // Tables<double> or Tables<float>
  double* stats_dev = nullptr;
  double* scratch_dev = nullptr;     // F16_NUM_STATE_FIELDS + 12 doubles
  double snapshot[F16_NUM_STATE_FIELDS + 12];
  void* state = nullptr;
  float* obs = nullptr;
  float* reward = nullptr;
  uint8_t* done = nullptr;
  uint8_t* truncated = nullptr;
  float* terminal_obs = nullptr;
  float* ep_return = nullptr;
  int32_t* ep_len = nullptr;
  float* actions_stage = nullptr;    // device staging for f16_step_host
  int64_t env_id_base = 0;
  uint64_t seed = 0;
  uint32_t step_counter = 0;
  double env_steps = 0.0;            // host-side count for stats[6]
  int num_sms = 1, ctas_per_sm = 1;  // persistent grid of the step kernel
  int ring = 0, ring_head = 0;       // observation layout (OBS_STACKED / OBS_RING / OBS_FRAME); ring layout: next slot to write
  f16_done_record* done_list = nullptr;   // frame layout: where the step kernel appends finished envs
  int32_t* done_count = nullptr;
  int ground = 1;                         // ground reactions (f16_set_ground_reactions); default: on in FP64 mode, off in FP32 mode
  // near-ground tiles first (f16_step_kernel): three rotating {list, count, per-tile flag} buffers and the number of
  // whole-batch steps taken, which selects this step's roles
  int32_t* hot_list = nullptr;            // 3 x hot_cap
  int32_t* hot_count = nullptr;           // 3
  uint8_t* hot_flag = nullptr;            // 3 x tiles
  int hot_cap = 0;
  int64_t hot_phase = 0;
};

template <typename R>
static int upload_tables(f16_ctx* c) {
  Tables<R>* h = new (std::nothrow) Tables<R>;
  if (!h) return fail("out of host memory");
  host::build_tables<R>(h);
  cudaError_t e = cudaMalloc(&c->tables_dev, sizeof(Tables<R>));
  if (e == cudaSuccess) e = cudaMemcpy(c->tables_dev, h, sizeof(Tables<R>), cudaMemcpyHostToDevice);
  delete h;
  if (e != cudaSuccess) return fail("table upload failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" {

// shared with f16_rollout.cu
int f16_internal_fail(const char* msg) { return fail("%s", msg); }
void f16_internal_count_launch(void) { g_launches++; }

// shared with f16_hostwin.cu: the buffers a frame-layout env is bound to
int f16_internal_frame_buffers(f16_handle h, int64_t* n, int* device, float** obs_frame, float** reward, uint8_t** done,
                               uint8_t** truncated, float** actions_stage) {
  if (!h || !h->state) return fail("the env handle is NULL or not bound");
  if (h->ring != OBS_FRAME) return fail("the env is not bound in the frame layout (f16_bind_frames)");
  *n = h->L.n; *device = h->device; *obs_frame = h->obs; *reward = h->reward; *done = h->done; *truncated = h->truncated;
  *actions_stage = h->actions_stage;
  return 0;
}

// shared with f16_hostwin.cu: where the frame layout's step kernel stores the newest frames (device memory, or pinned
// host memory mapped into the device's address space)
int f16_internal_set_obs_frame(f16_handle h, float* obs_frame) {
  if (!h || !h->state || h->ring != OBS_FRAME) return fail("the env is not bound in the frame layout (f16_bind_frames)");
  if (!obs_frame) return fail("obs_frame is NULL");
  h->obs = obs_frame;
  return 0;
}

const char* f16_last_error(void) { return g_err.c_str(); }
const char* f16_version(void) { return "f16_b200 0.1 (sm_100a)"; }
int64_t f16_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int f16_num_state_fields(void) { return F16_NUM_STATE_FIELDS; }

static int create_impl(f16_ctx* c, int64_t n_envs, int device, int mode) {
  c->device = device;
  c->mode = mode;
  c->L = make_layout(n_envs, mode);
  c->ground = mode == F16_MODE_FP64 ? 1 : 0;
  // constants shared by every context on this device
  MassSetT<double> ms[MS_COUNT];
  MassSetT<float> msf[MS_COUNT];
  host::build_mass_sets(ms);
  for (int i = 0; i < MS_COUNT; ++i) host::convert_mass_set(ms[i], &msf[i]);
  CUDA_OK(cudaMemcpyToSymbol(c_msets, ms, sizeof(ms)));
  CUDA_OK(cudaMemcpyToSymbol(c_msets_f, msf, sizeof(msf)));
  int rc = mode == F16_MODE_FP64 ? upload_tables<double>(c) : upload_tables<float>(c);
  if (rc) return rc;
  CUDA_OK(cudaMalloc(&c->stats_dev, F16_NUM_STATS * sizeof(double)));
  CUDA_OK(cudaMemset(c->stats_dev, 0, F16_NUM_STATS * sizeof(double)));
  CUDA_OK(cudaMalloc(&c->scratch_dev, 2 * (F16_NUM_STATE_FIELDS + 12) * sizeof(double)));
  // canonical snapshot: always computed in double, with a double table image
  {
    Tables<double>* hT = new (std::nothrow) Tables<double>;
    if (!hT) return fail("out of host memory");
    host::build_tables<double>(hT);
    Tables<double>* dT = nullptr;
    CUDA_OK(cudaMalloc(&dT, sizeof(Tables<double>)));
    CUDA_OK(cudaMemcpy(dT, hT, sizeof(Tables<double>), cudaMemcpyHostToDevice));
    delete hT;
    double ic[F16_NUM_STATE_FIELDS];
    host::initial_condition(900.0, 5000.0, ic);
    double* d_ic = c->scratch_dev;
    double* d_out = c->scratch_dev + (F16_NUM_STATE_FIELDS + 12);
    CUDA_OK(cudaMemcpy(d_ic, ic, sizeof(ic), cudaMemcpyHostToDevice));
    f16_init_snapshot_kernel<<<1, 32>>>(dT, d_ic, d_out);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(c->snapshot, d_out, sizeof(c->snapshot), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaFree(dT));
    CUDA_OK(cudaMemcpyToSymbol(c_snapshot, c->snapshot, F16_NUM_STATE_FIELDS * sizeof(double)));
    CUDA_OK(cudaMemcpyToSymbol(c_snapshot_props, c->snapshot + F16_NUM_STATE_FIELDS, 12 * sizeof(double)));
  }
  CUDA_OK(cudaMalloc(&c->actions_stage, (size_t)n_envs * F16_ACTION_DIM * sizeof(float)));
  {
    // the early CTAs can take up to a quarter of the batch's tiles (random actions keep ~5 % of the envs below 120 ft)
    const int warps = (mode == F16_MODE_FP64 ? StepShape<double>::BLK : StepShape<float>::BLK) / 32;
    const int64_t quarter = (c->L.tiles / 4 + warps - 1) / warps * warps;
    c->hot_cap = (int)std::max<int64_t>(warps, std::min<int64_t>(quarter, 1 << 20));
    CUDA_OK(cudaMalloc(&c->hot_list, 3 * (size_t)c->hot_cap * sizeof(int32_t)));
    CUDA_OK(cudaMalloc(&c->hot_count, 3 * sizeof(int32_t)));
    CUDA_OK(cudaMalloc(&c->hot_flag, 3 * (size_t)c->L.tiles));
    CUDA_OK(cudaMemset(c->hot_count, 0, 3 * sizeof(int32_t)));
    CUDA_OK(cudaMemset(c->hot_flag, 0, 3 * (size_t)c->L.tiles));
  }

  CUDA_OK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  if (mode == F16_MODE_FP64) CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->ctas_per_sm, f16_step_kernel<double, OBS_RING, true>, StepShape<double>::BLK, 0));
  else CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->ctas_per_sm, f16_step_kernel<float, OBS_RING, false>, StepShape<float>::BLK, 0));
  if (c->ctas_per_sm < 1) c->ctas_per_sm = 1;
  return 0;
}

int f16_create(f16_handle* out, int64_t n_envs, int device, int mode) {
  if (!out) return fail("f16_create: out is NULL");
  *out = nullptr;
  if (n_envs <= 0) return fail("f16_create: n_envs must be positive (got %lld)", (long long)n_envs);
  if (mode != F16_MODE_FP64 && mode != F16_MODE_FP32) return fail("f16_create: unknown mode %d", mode);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("f16_create: no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail("f16_create: device %d out of range [0,%d)", device, ndev);
  CUDA_OK(cudaSetDevice(device));
  f16_ctx* c = new (std::nothrow) f16_ctx;
  if (!c) return fail("out of host memory");
  // everything that can fail after the context exists runs in create_impl, so that one f16_destroy releases whatever
  // had been allocated by then
  const int rc = create_impl(c, n_envs, device, mode);
  if (rc) { f16_destroy(c); return rc; }
  *out = c;
  return 0;
}

int f16_destroy(f16_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->tables_dev);
  cudaFree(h->stats_dev);
  cudaFree(h->scratch_dev);
  cudaFree(h->actions_stage);
  cudaFree(h->hot_list);
  cudaFree(h->hot_count);
  cudaFree(h->hot_flag);
  delete h;
  return 0;
}

size_t f16_state_bytes(f16_handle h) { return h ? h->L.total : 0; }

int f16_set_ground_reactions(f16_handle h, int on) {
  if (!h) return fail("f16_set_ground_reactions: NULL handle");
  h->ground = on ? 1 : 0;
  return 0;
}
int f16_get_ground_reactions(f16_handle h) { return h ? h->ground : -1; }

int f16_bind(f16_handle h, void* state, float* obs, float* reward, uint8_t* done, uint8_t* truncated, float* terminal_obs,
             float* ep_return, int32_t* ep_len) {
  if (!h) return fail("f16_bind: NULL handle");
  if (!state || !obs || !reward || !done || !truncated) return fail("f16_bind: state, obs, reward, done and truncated are required");
  if (((uintptr_t)state & 255) != 0) return fail("f16_bind: state must be 256-byte aligned");
  CUDA_OK(cudaSetDevice(h->device));
  h->state = state; h->obs = obs; h->reward = reward; h->done = done; h->truncated = truncated;
  h->terminal_obs = terminal_obs; h->ep_return = ep_return; h->ep_len = ep_len;
  h->ring = OBS_STACKED; h->ring_head = 0;
  h->done_list = nullptr; h->done_count = nullptr;
  CUDA_OK(cudaMemset(state, 0, h->L.total));
  return 0;
}

int f16_bind_ring(f16_handle h, void* state, float* obs_ring, float* reward, uint8_t* done, uint8_t* truncated, float* terminal_obs,
                  float* ep_return, int32_t* ep_len) {
  int rc = f16_bind(h, state, obs_ring, reward, done, truncated, terminal_obs, ep_return, ep_len);
  if (rc) return rc;
  h->ring = OBS_RING;
  h->ring_head = 0;
  return 0;
}

int f16_bind_frames(f16_handle h, void* state, float* obs_frame, float* reward, uint8_t* done, uint8_t* truncated,
                    f16_done_record* done_list, int32_t* done_count) {
  if ((done_list == nullptr) != (done_count == nullptr)) return fail("f16_bind_frames: done_list and done_count go together");
  int rc = f16_bind(h, state, obs_frame, reward, done, truncated, nullptr, nullptr, nullptr);
  if (rc) return rc;
  h->ring = OBS_FRAME;
  h->done_list = done_list;
  h->done_count = done_count;
  return 0;
}

int f16_set_done_list(f16_handle h, f16_done_record* done_list, int32_t* done_count) {
  if (!h) return fail("f16_set_done_list: NULL handle");
  if (h->ring != OBS_FRAME) return fail("f16_set_done_list: the env is not bound in the frame layout (f16_bind_frames)");
  if ((done_list == nullptr) != (done_count == nullptr)) return fail("f16_set_done_list: done_list and done_count go together");
  h->done_list = done_list;
  h->done_count = done_count;
  return 0;
}

int f16_obs_window(f16_handle h, int* first_row) {
  if (!h || !first_row) return fail("f16_obs_window: NULL argument");
  // stacked layout: rows 0..9; ring layout: rows slot+1 .. slot+10 of the slot written by the last step
  *first_row = h->ring == OBS_RING ? ((h->ring_head + F16_OBS_FRAMES - 1) % F16_OBS_FRAMES) + 1 : 0;
  return 0;
}

int f16_set_env_id_base(f16_handle h, int64_t base) {
  if (!h) return fail("NULL handle");
  h->env_id_base = base;
  return 0;
}

static int launch_reset(f16_handle h, const uint8_t* mask, const float* goals, uint64_t seed, bool carry, const float* last_actions,
                        void* stream) {
  CUDA_OK(cudaSetDevice(h->device));
  h->seed = seed;
  ResetArgs a;
  a.state = h->state; a.mask = mask; a.goals = goals; a.obs = h->obs;
  a.n = h->L.n;
  a.seed = seed; a.env_id_base = h->env_id_base;
  a.obs_rows = h->ring == OBS_RING ? 2 * F16_OBS_FRAMES : h->ring == OBS_FRAME ? 1 : F16_OBS_FRAMES;
  a.env_pitch = h->ring == OBS_STACKED ? (size_t)F16_OBS_FRAMES * F16_OBS_FEATURES : (size_t)F16_OBS_FEATURES;
  a.row_pitch = h->ring == OBS_RING ? (size_t)h->L.n * F16_OBS_FEATURES : (size_t)F16_OBS_FEATURES;
  a.tables = h->tables_dev; a.last_actions = last_actions;
  unsigned grid = (unsigned)((h->L.n + BLOCK - 1) / BLOCK);
  const cudaStream_t st = (cudaStream_t)stream;
  if (h->mode == F16_MODE_FP64) {
    if (carry) f16_reset_kernel<double, true><<<grid, BLOCK, 0, st>>>(a);
    else f16_reset_kernel<double, false><<<grid, BLOCK, 0, st>>>(a);
  } else {
    if (carry) f16_reset_kernel<float, true><<<grid, BLOCK, 0, st>>>(a);
    else f16_reset_kernel<float, false><<<grid, BLOCK, 0, st>>>(a);
  }
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int f16_reset(f16_handle h, const uint8_t* mask, const float* goals, uint64_t seed, void* stream) {
  if (!h) return fail("f16_reset: NULL handle");
  if (!h->state) return fail("f16_reset: call f16_bind first");
  return launch_reset(h, mask, goals, seed, false, nullptr, stream);
}

int f16_reset_carryover(f16_handle h, const uint8_t* mask, const float* goals, uint64_t seed, const float* last_actions, void* stream) {
  if (!h) return fail("f16_reset_carryover: NULL handle");
  if (!h->state) return fail("f16_reset_carryover: call f16_bind first");
  return launch_reset(h, mask, goals, seed, true, last_actions, stream);
}

// One launch of the step kernel over envs [first, first + count); first is a multiple of 32.
static int launch_step(f16_handle h, const float* actions, int auto_reset, int64_t first, int64_t count, uint32_t step_counter, void* stream) {
  if (auto_reset < 0 || auto_reset > F16_AUTO_RESET_CARRYOVER) return fail("f16_step: auto_reset must be 0, 1 or 2 (got %d)", auto_reset);
  if (auto_reset == F16_AUTO_RESET_CARRYOVER && !h->ground)
    return fail("f16_step: the carry-over reset (auto_reset = 2) is part of the reference-detail build of the step kernel; "
                "turn ground reactions on first (f16_set_ground_reactions)");
  StepArgs a;
  a.state = h->state; a.tables = h->tables_dev; a.actions = actions; a.obs = h->obs; a.reward = h->reward;
  a.done = h->done; a.truncated = h->truncated; a.terminal_obs = h->terminal_obs; a.ep_return = h->ep_return;
  a.ep_len = h->ep_len; a.stats = h->stats_dev; a.n = first + count;
  a.seed = h->seed; a.env_id_base = h->env_id_base; a.step_counter = step_counter; a.auto_reset = auto_reset;
  a.ring_slot = h->ring_head;
  a.ring_pitch = (size_t)h->L.n * F16_OBS_FEATURES;
  a.done_list = h->done_list; a.done_count = h->done_count;
  a.tile0 = first / 32;
  const int64_t tiles = (count + 31) / 32;
  const cudaStream_t st = (cudaStream_t)stream;
  // near-ground tiles first: whole-batch steps of the ground-reaction builds only (a step taken in pieces,
  // f16_step_range, keeps the plain tile order)
  const bool hot = h->ground && first == 0 && count == h->L.n;
  a.hot_cap = hot ? h->hot_cap : 0;
  if (hot) {
    const int cur = (int)(h->hot_phase % 3), nxt = (int)((h->hot_phase + 1) % 3), clr = (int)((h->hot_phase + 2) % 3);
    a.hot_list_cur = h->hot_list + (size_t)cur * h->hot_cap; a.hot_count_cur = h->hot_count + cur; a.hot_flag_cur = h->hot_flag + (size_t)cur * h->L.tiles;
    a.hot_list_next = h->hot_list + (size_t)nxt * h->hot_cap; a.hot_count_next = h->hot_count + nxt; a.hot_flag_next = h->hot_flag + (size_t)nxt * h->L.tiles;
    a.hot_count_clr = h->hot_count + clr; a.hot_flag_clr = h->hot_flag + (size_t)clr * h->L.tiles;
    h->hot_phase += 1;
  } else {
    a.hot_list_cur = nullptr; a.hot_count_cur = nullptr; a.hot_flag_cur = nullptr; a.hot_list_next = nullptr; a.hot_count_next = nullptr;
    a.hot_flag_next = nullptr; a.hot_count_clr = nullptr; a.hot_flag_clr = nullptr;
  }
#define F16_LAUNCH_STEP_G(R, G)                                                                             \
  do {                                                                                                    \
    constexpr int BLK = StepShape<R>::BLK;                                                                \
    const unsigned grid = (unsigned)((tiles + BLK / 32 - 1) / (BLK / 32) + (G ? a.hot_cap / (BLK / 32) : 0)); \
    if (h->ring == OBS_FRAME) f16_step_kernel<R, OBS_FRAME, G><<<grid, BLK, 0, st>>>(a);                  \
    else if (h->ring == OBS_RING) f16_step_kernel<R, OBS_RING, G><<<grid, BLK, 0, st>>>(a);               \
    else f16_step_kernel<R, OBS_STACKED, G><<<grid, BLK, 0, st>>>(a);                                     \
  } while (0)
#define F16_LAUNCH_STEP(R)                                                                                  \
  do {                                                                                                    \
    if (h->ground) F16_LAUNCH_STEP_G(R, true);                                                            \
    else F16_LAUNCH_STEP_G(R, false);                                                                     \
  } while (0)
  // One register budget per precision (StepShape): the float kernel runs at 4 CTAs x 128 threads per SM (128
  // registers) - with the ring layout that is the faster build at every batch size (profiles/r2_ab_fp32_variants.txt:
  // 0.3103 against 0.3127 ms per step of 1M envs for the 168-register build, and one wave instead of two at 65 536 envs)
  if (h->mode == F16_MODE_FP64) F16_LAUNCH_STEP(double);
  else F16_LAUNCH_STEP(float);
#undef F16_LAUNCH_STEP
#undef F16_LAUNCH_STEP_G
  g_launches++;
  h->env_steps += (double)count;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int f16_step(f16_handle h, const float* actions, int auto_reset, void* stream) {
  if (!h) return fail("f16_step: NULL handle");
  if (!h->state) return fail("f16_step: call f16_bind first");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->ring == OBS_FRAME && h->done_count) CUDA_OK(cudaMemsetAsync(h->done_count, 0, sizeof(int32_t), (cudaStream_t)stream));
  const uint32_t counter = h->step_counter++;
  int rc = launch_step(h, actions, auto_reset, 0, h->L.n, counter, stream);
  if (h->ring == OBS_RING) h->ring_head = (h->ring_head + 1) % F16_OBS_FRAMES;
  return rc;
}

int f16_step_begin(f16_handle h, void* stream) {
  if (!h) return fail("f16_step_begin: NULL handle");
  if (!h->state) return fail("f16_step_begin: call f16_bind first");
  if (h->ring != OBS_FRAME) return fail("f16_step_begin: only the frame layout can be stepped in pieces (f16_bind_frames)");
  CUDA_OK(cudaSetDevice(h->device));
  h->step_counter++;
  if (h->done_count) CUDA_OK(cudaMemsetAsync(h->done_count, 0, sizeof(int32_t), (cudaStream_t)stream));
  return 0;
}

int f16_step_range(f16_handle h, const float* actions, int auto_reset, int64_t first, int64_t count, void* stream) {
  if (!h) return fail("f16_step_range: NULL handle");
  if (!h->state) return fail("f16_step_range: call f16_bind first");
  if (h->ring != OBS_FRAME) return fail("f16_step_range: only the frame layout can be stepped in pieces (f16_bind_frames)");
  if (first < 0 || count <= 0 || first + count > h->L.n || (first & 31) != 0)
    return fail("f16_step_range: bad range [%lld, %lld) (first must be a multiple of 32)", (long long)first, (long long)(first + count));
  CUDA_OK(cudaSetDevice(h->device));
  return launch_step(h, actions, auto_reset, first, count, h->step_counter - 1, stream);
}

int f16_step_host(f16_handle h, const float* actions_host, int auto_reset, float* obs_host, float* reward_host,
                  uint8_t* done_host, uint8_t* truncated_host, void* stream) {
  if (!h) return fail("f16_step_host: NULL handle");
  if (!actions_host) return fail("f16_step_host: actions_host is NULL");
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)h->L.n;
  CUDA_OK(cudaMemcpyAsync(h->actions_stage, actions_host, n * F16_ACTION_DIM * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = f16_step(h, h->actions_stage, auto_reset, stream);
  if (rc) return rc;
  if (obs_host && h->ring == OBS_FRAME) return fail("f16_step_host: the frame layout keeps no stacked observations on the device; use f16_hostwin_step");
  if (obs_host) {
    const size_t stack_bytes = F16_OBS_FRAMES * F16_OBS_FEATURES * sizeof(float);
    if (h->ring == OBS_RING) {
      // the window is ten (N,15) planes of the slot-major ring: one strided device->host copy per row of the stacks
      int first = 0;
      f16_obs_window(h, &first);
      const size_t row_bytes = F16_OBS_FEATURES * sizeof(float);
      for (int k = 0; k < F16_OBS_FRAMES; ++k)
        CUDA_OK(cudaMemcpy2DAsync(obs_host + (size_t)k * F16_OBS_FEATURES, stack_bytes, h->obs + (size_t)(first + k) * n * F16_OBS_FEATURES,
                                  row_bytes, row_bytes, n, cudaMemcpyDeviceToHost, st));
    } else {
      CUDA_OK(cudaMemcpyAsync(obs_host, h->obs, n * stack_bytes, cudaMemcpyDeviceToHost, st));
    }
  }
  if (reward_host) CUDA_OK(cudaMemcpyAsync(reward_host, h->reward, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (done_host) CUDA_OK(cudaMemcpyAsync(done_host, h->done, n, cudaMemcpyDeviceToHost, st));
  if (truncated_host) CUDA_OK(cudaMemcpyAsync(truncated_host, h->truncated, n, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

static int pack_range(f16_handle h, int64_t first, int64_t count, double* dev_out, cudaStream_t st) {
  unsigned grid = (unsigned)((count + 127) / 128);
  if (h->mode == F16_MODE_FP64) f16_pack_kernel<double><<<grid, 128, 0, st>>>(h->state, count, first, dev_out);
  else f16_pack_kernel<float><<<grid, 128, 0, st>>>(h->state, count, first, dev_out);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}
static int unpack_range(f16_handle h, int64_t first, int64_t count, const double* dev_in, cudaStream_t st) {
  unsigned grid = (unsigned)((count + 127) / 128);
  if (h->mode == F16_MODE_FP64) f16_unpack_kernel<double><<<grid, 128, 0, st>>>(h->state, count, first, dev_in);
  else f16_unpack_kernel<float><<<grid, 128, 0, st>>>(h->state, count, first, dev_in);
  g_launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int f16_get_state(f16_handle h, int64_t env, double* out, int n) {
  if (!h || !h->state) return fail("f16_get_state: handle not bound");
  if (n != F16_NUM_STATE_FIELDS) return fail("f16_get_state: n must be %d", (int)F16_NUM_STATE_FIELDS);
  if (env < 0 || env >= h->L.n) return fail("f16_get_state: env %lld out of range", (long long)env);
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaDeviceSynchronize());
  int rc = pack_range(h, env, 1, h->scratch_dev, 0);
  if (rc) return rc;
  CUDA_OK(cudaMemcpy(out, h->scratch_dev, n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}
int f16_set_state(f16_handle h, int64_t env, const double* in, int n) {
  if (!h || !h->state) return fail("f16_set_state: handle not bound");
  if (n != F16_NUM_STATE_FIELDS) return fail("f16_set_state: n must be %d", (int)F16_NUM_STATE_FIELDS);
  if (env < 0 || env >= h->L.n) return fail("f16_set_state: env %lld out of range", (long long)env);
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(h->scratch_dev, in, n * sizeof(double), cudaMemcpyHostToDevice));
  int rc = unpack_range(h, env, 1, h->scratch_dev, 0);
  if (rc) return rc;
  CUDA_OK(cudaDeviceSynchronize());
  return 0;
}
int f16_pack_states(f16_handle h, double* packed_dev, void* stream) {
  if (!h || !h->state) return fail("f16_pack_states: handle not bound");
  CUDA_OK(cudaSetDevice(h->device));
  return pack_range(h, 0, h->L.n, packed_dev, (cudaStream_t)stream);
}
int f16_unpack_states(f16_handle h, const double* packed_dev, void* stream) {
  if (!h || !h->state) return fail("f16_unpack_states: handle not bound");
  CUDA_OK(cudaSetDevice(h->device));
  return unpack_range(h, 0, h->L.n, packed_dev, (cudaStream_t)stream);
}
int f16_set_env_step(f16_handle h, int64_t env, int32_t current_step) {
  if (!h || !h->state) return fail("f16_set_env_step: handle not bound");
  if (env < 0 || env >= h->L.n) return fail("f16_set_env_step: env out of range");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaDeviceSynchronize());
  char* p = (char*)h->state + (size_t)(env >> 5) * h->L.tile_bytes + h->L.e_off + ((size_t)EF_STEP * 32 + (size_t)(env & 31)) * 4;
  CUDA_OK(cudaMemcpy(p, &current_step, 4, cudaMemcpyHostToDevice));
  return 0;
}

int f16_get_snapshot(f16_handle h, double* state_out, double* props12_out) {
  if (!h) return fail("NULL handle");
  if (state_out) memcpy(state_out, h->snapshot, F16_NUM_STATE_FIELDS * sizeof(double));
  if (props12_out) memcpy(props12_out, h->snapshot + F16_NUM_STATE_FIELDS, 12 * sizeof(double));
  return 0;
}

int f16_get_stats(f16_handle h, double* out8, int reset, void* stream) {
  if (!h) return fail("NULL handle");
  CUDA_OK(cudaSetDevice(h->device));
  CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  if (out8) {
    CUDA_OK(cudaMemcpy(out8, h->stats_dev, F16_NUM_STATS * sizeof(double), cudaMemcpyDeviceToHost));
    out8[6] = h->env_steps;
  }
  if (reset) {
    CUDA_OK(cudaMemset(h->stats_dev, 0, F16_NUM_STATS * sizeof(double)));
    h->env_steps = 0.0;
  }
  return 0;
}
int f16_stats_device_ptr(f16_handle h, double** out) {
  if (!h || !out) return fail("NULL argument");
  *out = h->stats_dev;
  return 0;
}

}  // extern "C"
