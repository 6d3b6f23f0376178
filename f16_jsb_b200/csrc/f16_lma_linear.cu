// f16_lma_linear.cu - y = x W^T + b for the policy's tall-skinny Linear layers on the 5th-generation tensor cores
// (include/f16_lma.h).
//
// The layers of the reference's LMA policy (jsbsim_gym/LMA_features.py:221-279,315-385; SB3 MlpExtractor heads) have
// 17..160 input and 32..160 output features and 10^5..10^6 rows per AM-PPO minibatch: every one of them moves a few
// hundred MB through a contraction of at most 128 x 160. The library's FP32 SIMT kernels need 100-220 us per layer for
// that (forward + input-gradient GEMMs = 41 % of an update step); the bytes alone take 20-65 us at the
// HBM rate, and the arithmetic intensity (8-25 flop per byte) is beyond the FP32 FMA pipe at that rate but a few per
// cent of the tensor pipe. So: tcgen05.mma kind::tf32 with FP32 accumulation in tensor memory, and - because the
// reference computes in FP32 - the operands split into a TF32 head and a TF32 remainder (x = xh + xl, W = Wh + Wl),
// three MMAs per k-step (xl Wh + xh Wl + xh Wh): what is dropped is xl Wl (2^-22 relative; a fourth MMA for it was
// measured: no visible change in the error, 3 % slower because every MMA re-reads its operands from shared memory, which
// is what bounds the kernel), the rounding of the remainders (2^-23 of each operand) and the tensor core's truncating
// FP32 accumulation, i.e. the result is FP32-accurate to a few ulp of the sum's magnitude, like a reordered FP32 sum
// (tests/test_gpu_linear_tc.py).
//
// One CTA per SM, 448 threads, persistent over 128-row tiles of x (UMMA M = 128, N = out features, K = 8 per
// instruction); the roles meet only in mbarriers:
//   * warp 13, one thread, TMA producer: x as a 2-D tensor map, one box = 128 rows x 32 floats (a "chunk"; rows past the
//     end read as zeros) into a four-deep ring of row-major chunks (cp.async.bulk.tensor, complete_tx on `raw_full`).
//   * warps 0-7, two converter groups of 128 threads; group g takes the chunks with index = g mod 2 and owns operand
//     stage g: read the chunk (128-bit, conflict-free), split into xh / xl, store both into the stage in the canonical
//     K-major SWIZZLE_128B operand layout (rows of 32 floats = 128 B, 16-byte pieces XOR-ed with the row index inside
//     each 8-row atom), fence.proxy.async, arrive on the stage's `full` and on the chunk's `raw_empty`. Two groups
//     because one warp per scheduler cannot issue the conversion at the rate HBM delivers chunks. W (both parts) is
//     written once per CTA in the same layout. (The 17-feature input layer has rows that are not multiples of 16 B, so
//     no tensor map: its tiles are contiguous and the converters load them themselves, one tile ahead.)
//   * warp 12, MMA issue: the whole warp runs the loop so that descriptors and barrier addresses stay in uniform
//     registers (issued from one diverged lane, every operand went through R2UR and the issue loop itself - 125 cycles
//     per MMA - was the bottleneck of the input-heavy layers); one elected lane issues the MMAs of a chunk, commits
//     them to the stage's `empty` barrier and, after the last chunk of a tile, to the accumulator's `acc_full`. Two
//     accumulators alternate in tensor memory, so the MMAs of tile t+1 run while tile t is being read out.
//   * warps 8-11, epilogue: tcgen05.ld 32x32b (a thread = a row, 32 columns at a time), bias, a padded per-warp
//     staging tile in shared memory (conflict-free both ways), then stores of whole 128-byte lines (eight lanes per
//     row); arrive on `acc_empty` as soon as the accumulator has been read.
// Every mbarrier wait is bounded (trap after ~2 s) so a protocol error cannot hang the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/f16_lma.h"
#include "f16_tc_common.cuh"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
using namespace f16tc;
constexpr int TILE_M = 128, LOADERS = 128, GROUPS = 2, THREADS = GROUPS * LOADERS + 128 + 32 + 32;
constexpr int A_STAGES = GROUPS;                        // operand stages: one per converter group
constexpr int RAW_STAGES = 4;                           // chunks in the TMA ring
constexpr uint32_t RAW_STAGE = TILE_M * 128;            // 16 KB: 128 rows x 32 floats, row-major
constexpr uint32_t HALF_STAGE = TILE_M * 128;           // one operand part of one stage
constexpr uint32_t STAGE = 2 * HALF_STAGE;              // xh | xl
constexpr uint32_t EPI_ROW = 128 + 16;                  // 32 floats + 16 B of padding: conflict-free 128-bit accesses
constexpr uint32_t EPI_WARP = 32 * EPI_ROW;
// shared memory: [operand stages (1024-aligned)] [TMA ring] [W, both parts (1024-aligned)] [epilogue staging] [barriers]
constexpr uint32_t OFF_A = 0, OFF_RAW = A_STAGES * STAGE, OFF_B = OFF_RAW + RAW_STAGES * RAW_STAGE;
constexpr uint32_t BAR_BYTES = 192;
constexpr uint32_t TAIL_BYTES = 4 * EPI_WARP + BAR_BYTES;
static_assert(OFF_B % 1024 == 0, "W atoms need 1024-byte alignment");

struct LinArgs {
  const float* x; const float* w; const float* bias; float* y;
  int64_t rows, tiles;
  int kg;          // features per row of x (= row pitch of x and w)
  int kp;          // kg rounded up to a multiple of 8 (MMA k-steps); the padding columns are zero
  int n;           // out features of this launch = UMMA N, a multiple of 32
  int ldy;         // row pitch of y in floats (>= n: a launch may cover a group of the layer's output columns)
  uint32_t tmem_cols;
};

// shared-memory operand descriptor, K-major, SWIZZLE_128B: start address >> 4 and leading byte offset 1 (unused with a
// swizzle) in the low word; stride byte offset 1024 B (8 rows x 128 B) >> 4, descriptor version 1 (sm_100) and layout
// type 2 in the high word
constexpr uint32_t DESC_HI = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return ((uint64_t)DESC_HI << 32) | (uint64_t)(((saddr >> 4) & 0x3FFFu) | (1u << 16));
}
// instruction descriptor, kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
// byte offset of element (row, k) of a [rows][32-float] K-atom in the SWIZZLE_128B layout
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t k) { return row * 128u + ((((k >> 2) ^ row) & 7u) << 4) + ((k & 3u) << 2); }

// VEC: kg is a multiple of 32 and x is 16-byte aligned: chunks come through the tensor map, converter t takes the 16-byte
// piece t & 7 of rows (t >> 3) + 16 i. Otherwise (kg <= 32: the 17 input features) a chunk is the tile's 128 x kg
// contiguous floats, converter t takes elements t + 128 i.
template <bool VEC>
__global__ void __launch_bounds__(THREADS, 1) linear_tc_kernel(const LinArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw0 = smem_u32(smem_raw);
  const uint32_t base = (raw0 + 1023u) & ~1023u;                 // operand atoms need 1024-byte alignment
  uint8_t* const sm = smem_raw + (base - raw0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nka = (a.kp + 31) >> 5;                              // K atoms (chunks) per tile
  const uint32_t b_half = (uint32_t)nka * (uint32_t)a.n * 128u;  // one part of W
  const uint32_t OFF_EPI = OFF_B + 2u * b_half, OFF_BAR = OFF_EPI + 4 * EPI_WARP;
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if ((base - raw0) + OFF_BAR + BAR_BYTES > dyn) __trap();     // the launcher's slack did not cover the alignment
  }
  // barriers: full[2], empty[2] (operand stages), acc_full[2], acc_empty[2], raw_full[4], raw_empty[4]; then the
  // tensor-memory address
  const uint32_t sBar = base + OFF_BAR;
  const uint32_t bar_full = sBar, bar_empty = sBar + 16, bar_acc_full = sBar + 32, bar_acc_empty = sBar + 48, bar_raw_full = sBar + 64,
                 bar_raw_empty = sBar + 96, tmem_holder = sBar + 128;
  volatile uint32_t* const tmem_holder_p = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + 128);

  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(bar_full + 8 * s, LOADERS); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_acc_full + 8 * s, 1); mbar_init(bar_acc_empty + 8 * s, 128); }
    for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(bar_raw_full + 8 * s, 1); mbar_init(bar_raw_empty + 8 * s, LOADERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_holder), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand stages once (padding columns of a narrow chunk stay zero), then W in operand layout
  for (uint32_t o = tid * 16u; o < A_STAGES * STAGE; o += THREADS * 16u) *reinterpret_cast<float4*>(sm + OFF_A + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = tid; e < a.n * nka * 32; e += THREADS) {
    const int row = e / (nka * 32), k = e - row * (nka * 32);
    const float v = k < a.kg ? __ldg(a.w + (size_t)row * a.kg + k) : 0.f;
    float hi, lo;
    split_tf32(v, hi, lo);
    const uint32_t off = OFF_B + (uint32_t)(k >> 5) * ((uint32_t)a.n * 128u) + sw128((uint32_t)row, (uint32_t)(k & 31));
    *reinterpret_cast<float*>(sm + off) = hi;
    *reinterpret_cast<float*>(sm + b_half + off) = lo;
  }
  tc_fence_before();
  fence_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder_p;

  // items of this CTA: (tile, chunk) pairs in order; item i belongs to converter group i % 2 = operand stage i % 2
  int64_t my_tiles = 0;
  if ((int64_t)blockIdx.x < a.tiles) my_tiles = (a.tiles - 1 - blockIdx.x) / gridDim.x + 1;
  const int64_t items = my_tiles * nka;

  if (warp < 8) {
    // ======================================================================== converters
    const int g = warp >> 2, t = tid & (LOADERS - 1);
    const int64_t my_items = items > g ? (items - g + 1) / 2 : 0;
    uint8_t* const hi_p = sm + OFF_A + (uint32_t)g * STAGE;
    uint8_t* const lo_p = hi_p + HALF_STAGE;
    if (VEC) {
      for (int64_t u = 0; u < my_items; ++u) {
        const int64_t item = 2 * u + g;
        const uint32_t slot = (uint32_t)(item % RAW_STAGES), raw_use = (uint32_t)(item / RAW_STAGES);
        const uint8_t* const src = sm + OFF_RAW + slot * RAW_STAGE;
        mbar_wait(bar_raw_full + 8 * slot, raw_use & 1u);                    // the TMA box has landed
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(src + (uint32_t)((t >> 3) + 16 * i) * 128u + (uint32_t)(t & 7) * 16u);
        if (u > 0) mbar_wait(bar_empty + 8 * g, (uint32_t)(u - 1) & 1u);     // the MMAs that read this stage are done
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t m = (uint32_t)(t >> 3) + 16u * i;
          const uint32_t off = m * 128u + ((((uint32_t)t ^ m) & 7u) << 4);
          float4 h, l;
          split_tf32(v[i].x, h.x, l.x); split_tf32(v[i].y, h.y, l.y); split_tf32(v[i].z, h.z, l.z); split_tf32(v[i].w, h.w, l.w);
          *reinterpret_cast<float4*>(hi_p + off) = h;
          *reinterpret_cast<float4*>(lo_p + off) = l;
        }
        fence_async_smem();
        mbar_arrive(bar_full + 8 * g);
        mbar_arrive(bar_raw_empty + 8 * slot);                   // after the stores that consumed the loaded values
      }
    } else {
      const uint32_t magic = ((1u << 20) + (uint32_t)a.kg - 1u) / (uint32_t)a.kg;             // e / kg = (e * magic) >> 20 for e < 4096, kg <= 32
      float nxt[32];
      auto load = [&](int64_t u) {
        const int64_t tile = blockIdx.x + (2 * u + g) * (int64_t)gridDim.x;                  // nka = 1: the CTA's item index = its tile index
        const int64_t f0 = tile * TILE_M * (int64_t)a.kg + t, f1 = a.rows * (int64_t)a.kg;
#pragma unroll
        for (int i = 0; i < 32; ++i) nxt[i] = (i < a.kg && f0 + LOADERS * i < f1) ? __ldg(a.x + f0 + LOADERS * i) : 0.f;
      };
      if (my_items > 0) load(0);
      for (int64_t u = 0; u < my_items; ++u) {
        float cur[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
        if (u + 1 < my_items) load(u + 1);                       // in flight while this tile is converted
        if (u > 0) mbar_wait(bar_empty + 8 * g, (uint32_t)(u - 1) & 1u);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i < a.kg) {
            const uint32_t e = (uint32_t)t + (uint32_t)LOADERS * i, m = (e * magic) >> 20, k = e - m * (uint32_t)a.kg;
            float h, l;
            split_tf32(cur[i], h, l);
            const uint32_t off = sw128(m, k);
            *reinterpret_cast<float*>(hi_p + off) = h;
            *reinterpret_cast<float*>(lo_p + off) = l;
          }
        }
        fence_async_smem();
        mbar_arrive(bar_full + 8 * g);
      }
    }
  } else if (warp == 13) {
    // ======================================================================== TMA producer (one thread)
    if (VEC && lane == 0) {
      for (int64_t item = 0; item < items; ++item) {
        const uint32_t slot = (uint32_t)(item % RAW_STAGES), raw_use = (uint32_t)(item / RAW_STAGES);
        if (raw_use > 0) mbar_wait(bar_raw_empty + 8 * slot, (raw_use - 1) & 1u);
        const int64_t tl = item / nka;
        const int j = (int)(item - tl * nka);
        const int64_t tile = blockIdx.x + tl * gridDim.x;
        mbar_expect_tx(bar_raw_full + 8 * slot, RAW_STAGE);
        tma_load_2d(base + OFF_RAW + slot * RAW_STAGE, &tmap, bar_raw_full + 8 * slot, 32 * j, (int)(tile * TILE_M));
      }
    }
    __syncwarp();
  } else if (warp == 12) {
    // ======================================================================== MMA issue (whole warp, one elected lane issues)
    const uint32_t idesc = umma_idesc(a.n);
    const uint32_t b_atom = (uint32_t)a.n * 128u;
    int64_t item = 0;
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const uint32_t acc = (uint32_t)(tl & 1), acc_use = (uint32_t)(tl >> 1);
      if (acc_use > 0) mbar_wait(bar_acc_empty + 8 * acc, (acc_use - 1) & 1u);     // the epilogue has read this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)a.n;
      for (int j = 0; j < nka; ++j, ++item) {
        const uint32_t st = (uint32_t)(item & 1);                // = the converter group of this item
        mbar_wait(bar_full + 8 * st, (uint32_t)(item >> 1) & 1u);
        tc_fence_after();
        const int ksteps = ((a.kp - 32 * j) < 32 ? (a.kp - 32 * j) : 32) >> 3;
        const uint64_t dah0 = umma_desc(base + OFF_A + st * STAGE), dal0 = umma_desc(base + OFF_A + st * STAGE + HALF_STAGE);
        const uint64_t dbh0 = umma_desc(base + OFF_B + (uint32_t)j * b_atom), dbl0 = umma_desc(base + OFF_B + (uint32_t)j * b_atom + b_half);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < ksteps) {
              const uint64_t dk = (uint64_t)(2 * ks);            // 32 B along K = 2 units of the address field
#ifndef F16_LIN_MMAS
#define F16_LIN_MMAS 3
#endif
#if F16_LIN_MMAS == 4
              umma_tf32(tmem_d, dal0 + dk, dbl0 + dk, idesc, (j | ks) ? 1u : 0u);      // smallest terms first
              umma_tf32(tmem_d, dal0 + dk, dbh0 + dk, idesc, 1u);
#else
              umma_tf32(tmem_d, dal0 + dk, dbh0 + dk, idesc, (j | ks) ? 1u : 0u);
#endif
              umma_tf32(tmem_d, dah0 + dk, dbl0 + dk, idesc, 1u);
              umma_tf32(tmem_d, dah0 + dk, dbh0 + dk, idesc, 1u);
            }
          }
          umma_commit(bar_empty + 8 * st);
          if (j == nka - 1) umma_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================================================================== epilogue (warps 8-11 = accumulator rows 32 (warp - 8) ..)
    const int we = warp - 8;
    uint8_t* const stage = sm + OFF_EPI + (uint32_t)we * EPI_WARP;
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const uint32_t acc = (uint32_t)(tl & 1), acc_use = (uint32_t)(tl >> 1);
      const int64_t tile = blockIdx.x + tl * gridDim.x;
      const int64_t row0 = tile * TILE_M + we * 32;
      mbar_wait(bar_acc_full + 8 * acc, acc_use & 1u);
      tc_fence_after();
      for (int c0 = 0; c0 < a.n; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + acc * (uint32_t)a.n + ((uint32_t)(we * 32) << 16) + (uint32_t)c0, v);
        if (c0 + 32 >= a.n) {                                    // the accumulator has been read: hand it back
          tc_fence_before();
          mbar_arrive(bar_acc_empty + 8 * acc);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (a.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c0) + q);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          *reinterpret_cast<float4*>(stage + (uint32_t)lane * EPI_ROW + 16u * q) = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = (lane >> 3) + 4 * i;
          const float4 o = *reinterpret_cast<const float4*>(stage + (uint32_t)r * EPI_ROW + 16u * (lane & 7));
          if (row0 + r < a.rows) *reinterpret_cast<float4*>(a.y + (size_t)(row0 + r) * a.ldy + c0 + 4 * (lane & 7)) = o;
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
}

constexpr size_t SMEM_LIMIT = 227 * 1024;
size_t smem_needed(int kp, int n) { return OFF_B + 2 * (size_t)((kp + 31) / 32) * n * 128 + TAIL_BYTES; }
}  // namespace

// the widest group of output columns (a multiple of 32) whose W, in both parts, fits next to the rings
static int max_group(int in_features) {
  const int kp = (in_features + 7) / 8 * 8;
  const size_t per32 = 2 * (size_t)((kp + 31) / 32) * 32 * 128;
  const size_t room = SMEM_LIMIT - OFF_B - TAIL_BYTES;
  int g = (int)(room / per32) * 32;
  return g > 256 ? 256 : g;
}

extern "C" int f16_lma_linear_supported(int in_features, int out_features) {
  if (out_features < 32 || out_features % 32) return 0;
  if (in_features <= 0) return 0;
  if (in_features > 32 && in_features % 32) return 0;
  if (max_group(in_features) < 32) return 0;
  // layers wider than one group are computed in column groups (x is read once per group): at most four
  return (out_features + max_group(in_features) - 1) / max_group(in_features) <= 4 ? 1 : 0;
}

static int launch_group(int64_t rows, int in_features, int n, int ldy, const float* x, const float* weight, const float* bias, float* y,
                        cudaStream_t stream) {
  LinArgs a;
  a.x = x; a.w = weight; a.bias = bias; a.y = y;
  a.rows = rows; a.tiles = (rows + TILE_M - 1) / TILE_M;
  a.kg = in_features; a.kp = (in_features + 7) / 8 * 8; a.n = n; a.ldy = ldy;
  a.tmem_cols = 32;
  while ((int)a.tmem_cols < 2 * n) a.tmem_cols *= 2;              // two accumulators
  const bool vec = in_features % 32 == 0;
  // 1 KB of slack for the 1024-byte alignment of the operand atoms; the largest shapes get what is left below the
  // per-block limit (the kernel checks that its carve-up fits and traps otherwise)
  size_t smem = smem_needed(a.kp, a.n) + 1024;
  if (smem > SMEM_LIMIT) smem = SMEM_LIMIT;
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (vec) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess)
        return f16_internal_fail("f16_lma_linear_forward: the driver does not export cuTensorMapEncodeTiled");
      encode = (EncodeFn)fn;
    }
    // x as a 2-D tensor (features innermost), boxes of 32 features x 128 rows; rows past the end read as zeros
    const cuuint64_t gdim[2] = {(cuuint64_t)in_features, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)in_features * sizeof(float)};
    const cuuint32_t box[2] = {32, (cuuint32_t)TILE_M}, estride[2] = {1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return f16_internal_fail("f16_lma_linear_forward: cuTensorMapEncodeTiled failed");
  }
  auto kern = vec ? linear_tc_kernel<true> : linear_tc_kernel<false>;
  static bool attr_done[2] = {false, false};
  if (!attr_done[vec]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT) != cudaSuccess)
      return f16_internal_fail("f16_lma_linear_forward: cannot raise the shared-memory limit");
    attr_done[vec] = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  }
  int64_t grid = sms;                                             // one persistent CTA per SM
  if (grid > a.tiles) grid = a.tiles;
  kern<<<(unsigned)grid, THREADS, smem, stream>>>(a, tmap);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}

extern "C" int f16_lma_linear_forward(int64_t rows, int in_features, int out_features, const float* x, const float* weight,
                                      const float* bias, float* y, void* stream) {
  if (rows <= 0) return f16_internal_fail("f16_lma_linear_forward: rows must be positive");
  if (!x || !weight || !y) return f16_internal_fail("f16_lma_linear_forward: NULL pointer");
  if (!f16_lma_linear_supported(in_features, out_features))
    return f16_internal_fail("f16_lma_linear_forward: unsupported shape (out features a multiple of 32; in features <= 32 or a multiple of 32; W must fit shared memory in at most four column groups)");
  if ((((uintptr_t)y) & 15) || (((uintptr_t)x) & 15) || (((uintptr_t)weight) & 15) || (bias && (((uintptr_t)bias) & 15)))
    return f16_internal_fail("f16_lma_linear_forward: x, weight, y and bias must be 16-byte aligned");
  // column groups: as even as the 32-column granularity allows
  const int gmax = max_group(in_features);
  const int groups = (out_features + gmax - 1) / gmax;
  const int per = ((out_features / 32 + groups - 1) / groups) * 32;
  for (int c0 = 0; c0 < out_features; c0 += per) {
    const int n = out_features - c0 < per ? out_features - c0 : per;
    const int rc = launch_group(rows, in_features, n, out_features, x, weight + (size_t)c0 * in_features, bias ? bias + c0 : nullptr, y + c0,
                                (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}
