// f16_lma_wgrad_tc.cu - weight and bias gradients of the policy's Linear layers on the tensor cores
// (include/f16_lma.h): dW[out][in] = sum_rows dY[row][out] X[row][in], db[out] = sum_rows dY[row][out].
//
// After the forward and the input gradient moved to tcgen05 (f16_lma_linear.cu) the weight gradients were the largest
// item of an AM-PPO update step (24 %): a contraction over 10^5..10^6 rows with a 32..128 x 32..160 result, bound by
// reading X and dY once (60-110 us at the HBM rate where the FP32 slab kernel of f16_lma_wgrad.cu needs 160-190).
// Here D[n][k] = sum_m dY[m][n] X[m][k] is one UMMA accumulator, 128 lanes (n, the live ones are n < out) x `in` columns:
//   * both operands are "MN-major" for this product (the reduction index m is the row of the row-major activations), and
//     a 16-row x 32-float piece of an activation, its 32-byte pieces swizzled inside each four-row atom, IS the canonical
//     MN-major operand block (SWIZZLE_128B_BASE32B, the only layout in which the tensor core transposes 32-bit
//     operands), read through a descriptor with the transpose bits set. A chunk = 16 rows: out/32 blocks of dY and
//     in/32 blocks of X, TF32 head and remainder each
//     (the reference computes in FP32: three MMAs per 8-row k-step, dYl Xh + dYh Xl + dYh Xh; what is dropped is 2^-22).
//   * the tensor core's FP32 accumulation truncates, and here the chain is long (thousands of k-steps per CTA), so the
//     accumulator is read out and restarted every 64 rows (24 MMAs): the partial results are summed in FP32 with
//     round-to-nearest in a per-thread row of shared memory, and meet the other CTAs in float atomics on the
//     zero-initialised output at the end (as the slab kernel does). Measured max error / max |dW| over the policy's
//     shapes (tools/diag_wgrad_error.py): 2.1e-6 with windows of 256 rows, 0.6-0.7e-6 with 64 - what the FP32 slab
//     kernel has (0.4-1.5e-6) - for 3 % of the time.
//   * roles as in f16_lma_linear.cu: TMA producer (two tensor maps, one box of 16-64 rows x all features each per chunk),
//     two converter groups (split into head / remainder, swizzled stores, column sums of dY for the bias gradient on
//     the way), one MMA warp (uniform descriptors, one elected lane), four read-out warps (tcgen05.ld 32x32b; a thread
//     = an output feature n), two accumulators alternating in tensor memory.
// Built for out in {32, 64, 96, 128} and in a multiple of 32 up to 160 whose partial sums fit shared memory; the caller
// keeps the slab kernel for the rest (17 input features, the 4- and 1-wide heads, 160 -> 128).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/f16_lma.h"
#include "f16_tc_common.cuh"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
using namespace f16tc;
constexpr int LOADERS = 128, GROUPS = 2, THREADS = GROUPS * LOADERS + 128 + 32 + 32;
// warps: 0-7 converters (two groups of four; groups of eight were measured and are slower: the conversion is bound by
// shared-memory bandwidth - TMA writes, converter reads and stores, and the MMAs' operand reads add up to ~140 KB per
// 20 KB chunk - not by instruction issue), 8-11 read-out, 12 MMA issue, 13 TMA producer
constexpr int W_EPI = GROUPS * LOADERS / 32, W_MMA = W_EPI + 4, W_TMA = W_MMA + 1;
constexpr int FLUSH_ROWS = 64;                          // the accumulator is read out and restarted every 64 rows (24 MMAs)
constexpr int MAX_RAW_STAGES = 8, MAX_A_STAGES = 2 * GROUPS;
constexpr uint32_t BAR_BYTES = 256;
constexpr size_t SMEM_LIMIT = 227 * 1024;

struct WgArgs {
  float* dw; float* db;
  int64_t rows, chunks, chunks_per_cta;
  int k, n;                // in / out features
  int fm, fn;              // features on the M side (accumulator lanes) and on the N side (accumulator columns) of the product
  int spg;                 // operand stages per converter group (1 or 2)
  int raw_stages;          // chunks in the TMA ring (2..8)
  int flush_chunks;        // the accumulator is read out and restarted every flush_chunks chunks
  int swap;                // 0: D[n][k] = dY^T X (M side = dY); 1: D[k][n] = X^T dY (M side = X), chosen when in > out: the M side
                           // is always padded to 128 lanes, so the wider operand goes there
  uint32_t tmem_cols;
  long long* dbg;          // optional per-role cycle counters of CTA 0 (tuning)
};

// MN-major operand descriptor. 32-bit operands can only be transposed in the SWIZZLE_128B_BASE32B layout (type 1): rows of
// 128 B = 32 MN-elements, atoms of four rows, the 32-byte pieces of a row XOR-ed with the row index mod 4. Start address
// >> 4; leading byte offset = distance between 32-element blocks along MN (one BLOCK); stride byte offset = distance
// between four-row atoms along K (512 B); descriptor version 1
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t block_bytes) {
  const uint32_t hi = 32u | (1u << 14) | (1u << 29);
  return ((uint64_t)hi << 32) | (uint64_t)(((saddr >> 4) & 0x3FFFu) | ((block_bytes >> 4) << 16));
}
// kind::tf32, D = F32, A = B = TF32, both MN-major (bits 15, 16), N = N-side features, M = 128
__device__ __forceinline__ uint32_t wg_idesc(int fn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(fn >> 3) << 17) | (8u << 24);
}

// ROWS = rows per chunk (32 or 64: the per-chunk hand-offs cost ~650 cycles whatever the chunk holds, so a chunk should be
// 16-20 KB; the host picks the largest that fits shared memory)
template <int ROWS>
__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const WgArgs a, const __grid_constant__ CUtensorMap tmap_m,
                                                              const __grid_constant__ CUtensorMap tmap_n) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw0 = smem_u32(smem_raw);
  const uint32_t base = (raw0 + 1023u) & ~1023u;
  uint8_t* const sm = smem_raw + (base - raw0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t BLOCK = ROWS * 128;                         // ROWS rows x 32 floats
  const int FLUSH_CHUNKS = a.flush_chunks;
  const int A_STAGES = GROUPS * a.spg, RAW_STAGES = a.raw_stages;
  const int nbm = a.fm >> 5, nbn = a.fn >> 5, nb = nbm + nbn;    // 32-float blocks of the M side, of the N side, per chunk
  // shared memory: [operand stages, two per converter group: Mh | Ml | Nh | Nl blocks] [TMA ring: per slot M blocks | N blocks]
  // [partial sums: fm rows of fn + 4 floats] [bias sums: n floats] [barriers]
  const uint32_t STAGE = 2u * (uint32_t)nb * BLOCK, RAW_SLOT = (uint32_t)nb * BLOCK;
  const uint32_t OFF_A = 0, OFF_RAW = (uint32_t)A_STAGES * STAGE, OFF_PART = OFF_RAW + (uint32_t)RAW_STAGES * RAW_SLOT;
  const uint32_t PROW = (uint32_t)(a.fn + 4) * 4u;               // bytes per partial row (16 B of padding: conflict-free)
  const uint32_t OFF_DB = OFF_PART + (uint32_t)a.fm * PROW, OFF_BAR = (OFF_DB + (uint32_t)a.n * 4u + 15u) & ~15u;
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if ((base - raw0) + OFF_BAR + BAR_BYTES > dyn) __trap();
  }
  // barriers: full[4], empty[4] (operand stages), acc_full[2], acc_empty[2], raw_full[8], raw_empty[8]; tensor-memory address
  const uint32_t sBar = base + OFF_BAR;
  const uint32_t bar_full = sBar, bar_empty = sBar + 32, bar_acc_full = sBar + 64, bar_acc_empty = sBar + 80, bar_raw_full = sBar + 96,
                 bar_raw_empty = sBar + 160, tmem_holder = sBar + 224;
  volatile uint32_t* const tmem_holder_p = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 224);

  if (tid == 0) {
    for (int s = 0; s < MAX_A_STAGES; ++s) { mbar_init(bar_full + 8 * s, LOADERS); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_acc_full + 8 * s, 1); mbar_init(bar_acc_empty + 8 * s, 128); }
    for (int s = 0; s < MAX_RAW_STAGES; ++s) { mbar_init(bar_raw_full + 8 * s, 1); mbar_init(bar_raw_empty + 8 * s, LOADERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_holder), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // partial sums and bias sums start at zero; the operand stages too (the M-side descriptor always spans four blocks: what
  // it reads past fm/32 blocks only reaches accumulator lanes that are never read)
  for (uint32_t o = tid * 16u; o < OFF_PART; o += THREADS * 16u) *reinterpret_cast<float4*>(sm + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t o = OFF_PART + tid * 4u; o < OFF_BAR; o += THREADS * 4u) *reinterpret_cast<float*>(sm + o) = 0.f;
  tc_fence_before();
  fence_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder_p;

  // this CTA's slab of chunks, cut in windows of FLUSH_CHUNKS
  const int64_t c_begin = (int64_t)blockIdx.x * a.chunks_per_cta;
  int64_t c_end = c_begin + a.chunks_per_cta;
  if (c_end > a.chunks) c_end = a.chunks;
  const int64_t items = c_end > c_begin ? c_end - c_begin : 0;
  const int64_t windows = (items + FLUSH_CHUNKS - 1) / FLUSH_CHUNKS;

  if (warp < W_EPI) {
    // ======================================================================== converters
    const int g = warp >> 2, t = tid & (LOADERS - 1);
    const int64_t my_items = items > g ? (items - g + 1) / 2 : 0;
    const uint32_t m = (uint32_t)(t >> 3), p = (uint32_t)(t & 7);             // rows m + 16 r of the chunk, 16-byte piece of the block row
    const uint32_t dst_off = m * 128u + ((((p >> 1) ^ m) & 3u) << 5) + ((p & 1u) << 4);
    // the TMA ring holds the two activations row-major: [ROWS][fm] then [ROWS][fn] floats (one wide box each: TMA moves a
    // box row by row, and 128-byte rows - one 32-float block per box - could not keep up with HBM)
    const uint32_t pitch_m = (uint32_t)a.fm * 4u, pitch_n = (uint32_t)a.fn * 4u, raw_n = (uint32_t)ROWS * pitch_m;
    constexpr int RPT = ROWS / 16;                                            // rows per thread per block
    constexpr int MB_M = ROWS == 64 ? 2 : 4, MB_N = ROWS == 64 ? 2 : 5;          // blocks per side this instantiation is built for
    const bool dy_m = a.swap == 0;                                            // dY is the M side
    float bsum[4][4];                                                         // column sums of dY: block b, columns 4 p ..
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int i = 0; i < 4; ++i) bsum[b][i] = 0.f;
    long long t_raw = 0, t_empty = 0, t_conv = 0, t_fence = 0;
    for (int64_t u = 0; u < my_items; ++u) {
      const long long c0 = clock64();
      const int64_t item = 2 * u + g;
      const uint32_t slot = (uint32_t)(item % RAW_STAGES), raw_use = (uint32_t)(item / RAW_STAGES);
      const uint32_t st = a.spg == 2 ? 2u * (uint32_t)g + (uint32_t)(u & 1) : (uint32_t)g;
      const uint32_t use = a.spg == 2 ? (uint32_t)(u >> 1) : (uint32_t)u;
      const uint8_t* const src = sm + OFF_RAW + slot * RAW_SLOT;
      uint8_t* const stage = sm + OFF_A + st * STAGE;
      mbar_wait(bar_raw_full + 8 * slot, raw_use & 1u);
      const long long c1 = clock64();
      // stage: Mh blocks | Ml blocks | Nh blocks | Nl blocks. All loads of the chunk first (independent, so their latencies
      // overlap: a converter warp is alone on its scheduler), then split and store.
      float4 vm[MB_M][RPT], vn[MB_N][RPT];
#pragma unroll
      for (int b = 0; b < MB_M; ++b)
#pragma unroll
        for (int r = 0; r < RPT; ++r)
          if (b < nbm) vm[b][r] = *reinterpret_cast<const float4*>(src + (m + 16u * r) * pitch_m + (uint32_t)b * 128u + p * 16u);
#pragma unroll
      for (int b = 0; b < MB_N; ++b)
#pragma unroll
        for (int r = 0; r < RPT; ++r)
          if (b < nbn) vn[b][r] = *reinterpret_cast<const float4*>(src + raw_n + (m + 16u * r) * pitch_n + (uint32_t)b * 128u + p * 16u);
      const long long c2 = clock64();
      if (use > 0) mbar_wait(bar_empty + 8 * st, (use - 1) & 1u);             // the MMAs that read this stage are done
      const long long c3 = clock64();
#pragma unroll
      for (int b = 0; b < MB_M; ++b) {
        if (b < nbm) {
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            const float4 v = vm[b][r];
            if (dy_m) { bsum[b][0] += v.x; bsum[b][1] += v.y; bsum[b][2] += v.z; bsum[b][3] += v.w; }
            float4 h, l;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(stage + (uint32_t)b * BLOCK + dst_off + 2048u * r) = h;
            *reinterpret_cast<float4*>(stage + (uint32_t)(nbm + b) * BLOCK + dst_off + 2048u * r) = l;
          }
        }
      }
#pragma unroll
      for (int b = 0; b < MB_N; ++b) {
        if (b < nbn) {
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            const float4 v = vn[b][r];
            if (!dy_m && b < 4) { bsum[b & 3][0] += v.x; bsum[b & 3][1] += v.y; bsum[b & 3][2] += v.z; bsum[b & 3][3] += v.w; }
            float4 h, l;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(stage + (uint32_t)(2 * nbm + b) * BLOCK + dst_off + 2048u * r) = h;
            *reinterpret_cast<float4*>(stage + (uint32_t)(2 * nbm + nbn + b) * BLOCK + dst_off + 2048u * r) = l;
          }
        }
      }
      const long long c4 = clock64();
      fence_async_smem();
      mbar_arrive(bar_full + 8 * st);
      mbar_arrive(bar_raw_empty + 8 * slot);
      const long long c5 = clock64();
      t_raw += c1 - c0; t_empty += c3 - c2; t_conv += (c2 - c1) + (c4 - c3); t_fence += c5 - c4;
    }
    if (a.dbg && blockIdx.x == 0 && t == 0) {
      a.dbg[8 * g + 0] = t_raw; a.dbg[8 * g + 1] = t_empty; a.dbg[8 * g + 2] = t_conv; a.dbg[8 * g + 3] = t_fence; a.dbg[8 * g + 4] = my_items;
    }
    if (a.db) {
      float* const dbs = reinterpret_cast<float*>(sm + OFF_DB);
      const int nby = a.n >> 5;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (b < nby)
#pragma unroll
          for (int i = 0; i < 4; ++i) atomicAdd(dbs + 32 * b + 4 * (int)p + i, bsum[b][i]);
    }
  } else if (warp == W_TMA) {
    // ======================================================================== TMA producer (one thread)
    if (lane == 0) {
      long long t_wait = 0;
      const long long p0 = clock64();
      for (int64_t item = 0; item < items; ++item) {
        const uint32_t slot = (uint32_t)(item % RAW_STAGES), raw_use = (uint32_t)(item / RAW_STAGES);
        const long long q0 = clock64();
        if (raw_use > 0) mbar_wait(bar_raw_empty + 8 * slot, (raw_use - 1) & 1u);
        t_wait += clock64() - q0;
        const int row0 = (int)((c_begin + item) * ROWS);
        const uint32_t dst = base + OFF_RAW + slot * RAW_SLOT, bar = bar_raw_full + 8 * slot;
        mbar_expect_tx(bar, RAW_SLOT);
        tma_load_2d(dst, &tmap_m, bar, 0, row0);
        tma_load_2d(dst + (uint32_t)nbm * BLOCK, &tmap_n, bar, 0, row0);
      }
      if (a.dbg && blockIdx.x == 0) { a.dbg[16] = t_wait; a.dbg[17] = clock64() - p0; a.dbg[18] = items; }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ======================================================================== MMA issue (whole warp, one elected lane issues)
    const uint32_t idesc = wg_idesc(a.fn);
    int64_t item = 0;
    long long t_acc = 0, t_full = 0, t_issue = 0;
    const long long m0 = clock64();
    for (int64_t w = 0; w < windows; ++w) {
      const uint32_t acc = (uint32_t)(w & 1), acc_use = (uint32_t)(w >> 1);
      const long long w0 = clock64();
      if (acc_use > 0) mbar_wait(bar_acc_empty + 8 * acc, (acc_use - 1) & 1u);
      t_acc += clock64() - w0;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)a.fn;
      const int64_t w_end = (w + 1) * FLUSH_CHUNKS < items ? (w + 1) * FLUSH_CHUNKS : items;
      for (int first = 1; item < w_end; ++item, first = 0) {
        const int64_t u = item >> 1;                             // chunk u of converter group item & 1
        const uint32_t st = a.spg == 2 ? 2u * (uint32_t)(item & 1) + (uint32_t)(u & 1) : (uint32_t)(item & 1);
        const long long f0 = clock64();
        mbar_wait(bar_full + 8 * st, (uint32_t)(a.spg == 2 ? (u >> 1) : u) & 1u);
        const long long f1 = clock64();
        t_full += f1 - f0;
        tc_fence_after();
        const uint32_t s0 = base + OFF_A + st * STAGE;
        const uint64_t mh = mn_desc(s0, BLOCK), ml = mn_desc(s0 + (uint32_t)nbm * BLOCK, BLOCK);
        const uint64_t nh = mn_desc(s0 + 2u * (uint32_t)nbm * BLOCK, BLOCK), nl = mn_desc(s0 + (2u * (uint32_t)nbm + (uint32_t)nbn) * BLOCK, BLOCK);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < ROWS / 8; ++ks) {
            const uint64_t dk = (uint64_t)(64 * ks);             // the next 8 rows: 1024 B = 64 units of the address field
            umma_tf32(tmem_d, ml + dk, nh + dk, idesc, (first && ks == 0) ? 0u : 1u);
            umma_tf32(tmem_d, mh + dk, nl + dk, idesc, 1u);
            umma_tf32(tmem_d, mh + dk, nh + dk, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * st);
          if (item == w_end - 1) umma_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
        t_issue += clock64() - f1;
      }
    }
    if (a.dbg && blockIdx.x == 0 && lane == 0) { a.dbg[24] = t_acc; a.dbg[25] = t_full; a.dbg[26] = t_issue; a.dbg[27] = clock64() - m0; }
  } else {
    // ======================================================================== read-out (warps 8-11: thread = M-side feature i)
    const int i = (warp - W_EPI) * 32 + lane;
    const bool live_warp = (warp - W_EPI) * 32 < a.fm;               // fm is a multiple of 32: a warp is all live or all idle
    float* const prow = reinterpret_cast<float*>(sm + OFF_PART + (uint32_t)(live_warp ? i : 0) * PROW);
    for (int64_t w = 0; w < windows; ++w) {
      const uint32_t acc = (uint32_t)(w & 1), acc_use = (uint32_t)(w >> 1);
      mbar_wait(bar_acc_full + 8 * acc, acc_use & 1u);
      tc_fence_after();
      if (live_warp) {
        for (int c0 = 0; c0 < a.fn; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + acc * (uint32_t)a.fn + ((uint32_t)((warp - W_EPI) * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 s = *reinterpret_cast<float4*>(prow + c0 + 4 * q);
            s.x += v[4 * q]; s.y += v[4 * q + 1]; s.z += v[4 * q + 2]; s.w += v[4 * q + 3];
            *reinterpret_cast<float4*>(prow + c0 + 4 * q) = s;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_acc_empty + 8 * acc);
    }
    if (live_warp && windows > 0) {
      if (a.swap) for (int j = 0; j < a.fn; ++j) atomicAdd(a.dw + (size_t)j * a.k + i, prow[j]);       // D[k][n] -> dW[n][k]
      else        for (int j = 0; j < a.fn; ++j) atomicAdd(a.dw + (size_t)i * a.k + j, prow[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (a.db && tid < a.n && items > 0) atomicAdd(a.db + tid, reinterpret_cast<float*>(sm + OFF_DB)[tid]);
  if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
}

bool wg_swap(int k, int n) { return k > n && k <= 128; }
struct WgPlan { int rows, spg, raw; size_t smem; };
// the largest chunk, then the deepest rings, that fit shared memory (with 1 KB of alignment slack)
bool wg_plan(int k, int n, WgPlan* plan) {
  const size_t nb = (size_t)(k + n) / 32;
  const size_t fm = wg_swap(k, n) ? k : n, fn = wg_swap(k, n) ? n : k;
  const size_t fixed = fm * (fn + 4) * 4 + (size_t)n * 4 + 16 + BAR_BYTES + 1024;
  static const int opts[6][2] = {{2, 4}, {2, 3}, {1, 4}, {2, 2}, {1, 3}, {1, 2}};
  if (const char* ov = getenv("F16_WG_PLAN")) {                 // tuning override: "rows,stages per group,ring depth"
    int r = 0, sp = 0, rw = 0;
    if (sscanf(ov, "%d,%d,%d", &r, &sp, &rw) == 3 && (r == 32 || (r == 64 && fm <= 64 && fn <= 64)) && (sp == 1 || sp == 2) && rw >= 2 && rw <= MAX_RAW_STAGES) {
      const size_t need = (size_t)GROUPS * sp * 2 * nb * r * 128 + (size_t)rw * nb * r * 128 + fixed;
      if (need <= SMEM_LIMIT) { plan->rows = r; plan->spg = sp; plan->raw = rw; plan->smem = need; return true; }
    }
  }
  for (int rows = 64; rows >= 32; rows /= 2) {
    if ((size_t)rows * nb * 128 > 24 * 1024 && rows > 32) continue;   // a chunk of 16-24 KB is enough
    if (rows == 64 && (fm > 64 || fn > 64)) continue;               // the 64-row kernel is built for two blocks per side
    for (const auto& o : opts) {
      const size_t need = (size_t)GROUPS * o[0] * 2 * nb * rows * 128 + (size_t)o[1] * nb * rows * 128 + fixed;
      if (need <= SMEM_LIMIT) { plan->rows = rows; plan->spg = o[0]; plan->raw = o[1]; plan->smem = need; return true; }
    }
  }
  return false;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool make_map(EncodeFn encode, CUtensorMap* map, const float* p, int64_t rows, int cols, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows}, estride[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

extern "C" int f16_lma_linear_wgrad_tc_supported(int in_features, int out_features) {
  if (out_features < 32 || out_features > 128 || out_features % 32) return 0;
  if (in_features < 32 || in_features > 160 || in_features % 32) return 0;
  WgPlan plan;
  return wg_plan(in_features, out_features, &plan) ? 1 : 0;
}

extern "C" int f16_lma_linear_wgrad_tc(int64_t rows, int in_features, int out_features, const float* x, const float* dy, float* dweight,
                                       float* dbias, void* stream) {
  if (rows <= 0) return f16_internal_fail("f16_lma_linear_wgrad_tc: rows must be positive");
  if (!x || !dy || !dweight) return f16_internal_fail("f16_lma_linear_wgrad_tc: NULL pointer");
  if (!f16_lma_linear_wgrad_tc_supported(in_features, out_features))
    return f16_internal_fail("f16_lma_linear_wgrad_tc: unsupported shape (out features 32..128 and in features 32..160, multiples of 32, partial sums must fit shared memory)");
  if ((((uintptr_t)x) | ((uintptr_t)dy)) & 15) return f16_internal_fail("f16_lma_linear_wgrad_tc: x and dy must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dweight, 0, (size_t)in_features * out_features * sizeof(float), st) != cudaSuccess)
    return f16_internal_fail("f16_lma_linear_wgrad_tc: memset failed");
  if (dbias && cudaMemsetAsync(dbias, 0, (size_t)out_features * sizeof(float), st) != cudaSuccess)
    return f16_internal_fail("f16_lma_linear_wgrad_tc: memset failed");
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess)
      return f16_internal_fail("f16_lma_linear_wgrad_tc: the driver does not export cuTensorMapEncodeTiled");
    encode = (EncodeFn)fn;
  }
  WgPlan plan;
  wg_plan(in_features, out_features, &plan);
  alignas(64) CUtensorMap map_x, map_dy;
  if (!make_map(encode, &map_x, x, rows, in_features, plan.rows) || !make_map(encode, &map_dy, dy, rows, out_features, plan.rows))
    return f16_internal_fail("f16_lma_linear_wgrad_tc: cuTensorMapEncodeTiled failed");
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  }
  WgArgs a;
  a.dw = dweight; a.db = dbias;
  a.rows = rows; a.chunks = (rows + plan.rows - 1) / plan.rows;
  a.k = in_features; a.n = out_features;
  a.spg = plan.spg; a.raw_stages = plan.raw;
  a.dbg = nullptr;
  if (const char* d = getenv("F16_WG_DBG")) a.dbg = (long long*)strtoull(d, nullptr, 0);
  // contiguous slabs, a whole number of read-out windows per CTA
  int flush_rows = FLUSH_ROWS;
  if (const char* f = getenv("F16_WG_FLUSH")) { const int v = atoi(f); if (v >= plan.rows && v <= 4096 && v % plan.rows == 0) flush_rows = v; }   // tuning override
  const int flush_chunks = flush_rows / plan.rows;
  a.flush_chunks = flush_chunks;
  int64_t per = (a.chunks + sms - 1) / sms;
  per = (per + flush_chunks - 1) / flush_chunks * flush_chunks;
  a.chunks_per_cta = per;
  const int64_t grid = (a.chunks + per - 1) / per;
  a.swap = wg_swap(in_features, out_features) ? 1 : 0;
  a.fm = a.swap ? in_features : out_features;
  a.fn = a.swap ? out_features : in_features;
  a.tmem_cols = 32;
  while ((int)a.tmem_cols < 2 * a.fn) a.tmem_cols *= 2;
  auto kern = plan.rows == 64 ? wgrad_tc_kernel<64> : wgrad_tc_kernel<32>;
  static bool attr_done[2] = {false, false};
  const int ki = plan.rows == 64 ? 1 : 0;
  if (!attr_done[ki]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT) != cudaSuccess)
      return f16_internal_fail("f16_lma_linear_wgrad_tc: cannot raise the shared-memory limit");
    attr_done[ki] = true;
  }
  kern<<<(unsigned)grid, THREADS, plan.smem, st>>>(a, a.swap ? map_x : map_dy, a.swap ? map_dy : map_x);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
