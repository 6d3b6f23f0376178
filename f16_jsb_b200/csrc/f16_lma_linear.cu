// f16_lma_linear.cu - y = x W^T + b for the policy's tall-skinny Linear layers on the 5th-generation tensor cores
// (include/f16_lma.h).
//
// The layers of the reference's LMA policy (jsbsim_gym/LMA_features.py:221-279,315-385; SB3 MlpExtractor heads) have
// 17..160 input and 32..160 output features and 10^5..10^6 rows per AM-PPO minibatch: every one of them moves a few
// hundred MB through a contraction of at most 128 x 160. The library's FP32 SIMT kernels need 100-220 us per layer for
// that (forward + input-gradient GEMMs = 41 % of an update step); the bytes alone take 20-65 us at the
// HBM rate, and the arithmetic intensity (8-25 flop per byte) is beyond the FP32 FMA pipe at that rate but a few per
// cent of the tensor pipe. So: tcgen05.mma kind::tf32 with FP32 accumulation in tensor memory, and - because the
// reference computes in FP32 - the operands split into a TF32 head and an FP32 remainder (x = xh + xl, W = Wh + Wl),
// four MMAs per k-step (xh Wh + xl Wh + xh Wl + xl Wl; the tensor pipe is idle anyway): what is dropped is the rounding
// of the remainders to TF32 (2^-23 of each operand), i.e. the result is FP32-accurate to a few ulp like a reordered
// FP32 sum (tests/test_gpu_linear_tc.py).
//
// One CTA = 128 threads = one 128-row tile of x at a time (UMMA M = 128, N = out features, K = 8 per instruction):
//   * W (both parts) is written once per CTA into shared memory in the canonical K-major SWIZZLE_128B operand layout
//     (rows of 32 floats = 128 B, 16-byte chunks XOR-ed with the row index inside each 8-row atom);
//   * x streams in chunks of 128 rows x 32 floats: coalesced 128-bit loads into registers (issued one chunk ahead, so
//     they are in flight during the MMAs and the epilogue of the chunk before), split into xh / xl, stored into a
//     two-stage ring in the same operand layout; fence.proxy.async + CTA barrier; one thread issues the MMAs and
//     commits them to the stage's mbarrier (which frees the stage) and, after the last chunk of a tile, to the
//     accumulator's mbarrier;
//   * epilogue: every warp reads its 32 accumulator rows from tensor memory (tcgen05.ld 32x32b), adds the bias and
//     stores its rows.
// CTAs are persistent over tiles (grid = min(tiles, resident CTAs)); two to three are resident per SM, so one CTA's
// epilogue overlaps the others' loads. Every mbarrier wait is bounded (trap after ~2 s) so a logic error cannot hang
// the GPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_lma.h"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int THREADS = 128, TILE_M = 128;
constexpr uint32_t HALF_STAGE = TILE_M * 128;          // one operand part of one stage: 128 rows x 128 B
constexpr uint32_t STAGE = 2 * HALF_STAGE;             // xh | xl
constexpr uint32_t A_BYTES = 2 * STAGE;                // two stages

struct LinArgs {
  const float* x; const float* w; const float* bias; float* y;
  int64_t rows, tiles;
  int kg;          // features per row of x (= row pitch of x and w)
  int kp;          // kg rounded up to a multiple of 8 (MMA k-steps); the padding columns are zero
  int n;           // out features = UMMA N
  uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();      // ~2 s: a protocol error must not hang the GPU
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory operand descriptor, K-major, SWIZZLE_128B: start address >> 4, leading byte offset 1 (unused with a
// swizzle), stride byte offset 1024 B (8 rows x 128 B) >> 4, descriptor version 1 (sm_100), layout type 2
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3, M >> 4
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// TF32 head and TF32 remainder, both rounded to nearest (ties away): 10 explicit mantissa bits, the 13 low bits zero, so
// the tensor core reads exactly these values whether it truncates or rounds; v - hi - lo is at most 2^-23 |v|
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
  lo = __uint_as_float((__float_as_uint(v - hi) + 0x1000u) & 0xFFFFE000u);
}
// byte offset of element (row, k) of a [rows][32-float] K-atom in the SWIZZLE_128B layout
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t k) { return row * 128u + ((((k >> 2) ^ row) & 7u) << 4) + ((k & 3u) << 2); }

// VEC: kg is a multiple of 32 and x is 16-byte aligned: a chunk is 128 rows x 8 float4, thread t takes chunk column
// t & 7 of rows (t >> 3) + 16 i. Otherwise (kg <= 32: the 17 input features) a chunk is the tile's 128 x kg contiguous
// floats, thread t takes elements t + 128 i.
template <bool VEC>
__global__ void __launch_bounds__(THREADS) linear_tc_kernel(const LinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // operand atoms need 1024-byte alignment
  uint8_t* const sm = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nka = (a.kp + 31) >> 5;                              // K atoms (chunks) per tile
  const uint32_t b_half = (uint32_t)nka * (uint32_t)a.n * 128u;  // one part of W
  const uint32_t sA = base, sB = base + A_BYTES, sBar = sB + 2u * b_half;
  uint8_t* const pA = sm;
  uint8_t* const pB = sm + A_BYTES;
  const uint32_t bar_stage0 = sBar, bar_stage1 = sBar + 8, bar_acc = sBar + 16, tmem_holder = sBar + 24;
  volatile uint32_t* const tmem_holder_p = reinterpret_cast<volatile uint32_t*>(sm + A_BYTES + 2u * b_half + 24);

  if (tid == 0) {
    mbar_init(bar_stage0, 1); mbar_init(bar_stage1, 1); mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_holder), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero both stages once (padding columns and the unused tail of a narrow chunk stay zero), then W in operand layout
  for (uint32_t o = tid * 16u; o < A_BYTES; o += THREADS * 16u) *reinterpret_cast<float4*>(pA + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = tid; e < a.n * nka * 32; e += THREADS) {
    const int row = e / (nka * 32), k = e - row * (nka * 32);
    const float v = k < a.kg ? a.w[(size_t)row * a.kg + k] : 0.f;
    float hi, lo;
    split_tf32(v, hi, lo);
    const uint32_t off = (uint32_t)(k >> 5) * ((uint32_t)a.n * 128u) + sw128((uint32_t)row, (uint32_t)(k & 31));
    *reinterpret_cast<float*>(pB + off) = hi;
    *reinterpret_cast<float*>(pB + b_half + off) = lo;
  }
  tc_fence_before();
  fence_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_holder_p;
  const uint32_t idesc = umma_idesc(a.n);

  float4 pre[8];                                                 // the chunk in flight (VEC) / up to 32 scalars
  float* const pre_s = reinterpret_cast<float*>(pre);
  auto prefetch = [&](int64_t tile, int j) {
    const int64_t r0 = tile * TILE_M;
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = r0 + (tid >> 3) + 16 * i;
        pre[i] = r < a.rows ? __ldg(reinterpret_cast<const float4*>(a.x + (size_t)r * a.kg + 32 * j) + (tid & 7)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const int64_t e0 = r0 * a.kg, e1 = a.rows * (int64_t)a.kg;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int64_t e = e0 + tid + 128 * i;
        pre_s[i] = (i < a.kg && e < e1) ? __ldg(a.x + e) : 0.f;
      }
    }
  };
  auto stage_store = [&](int st) {
    uint8_t* const hi_p = pA + (uint32_t)st * STAGE;
    uint8_t* const lo_p = hi_p + HALF_STAGE;
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t m = (uint32_t)(tid >> 3) + 16u * i;
        const uint32_t off = m * 128u + ((((uint32_t)tid ^ m) & 7u) << 4);
        float4 h, l;
        split_tf32(pre[i].x, h.x, l.x); split_tf32(pre[i].y, h.y, l.y); split_tf32(pre[i].z, h.z, l.z); split_tf32(pre[i].w, h.w, l.w);
        *reinterpret_cast<float4*>(hi_p + off) = h;
        *reinterpret_cast<float4*>(lo_p + off) = l;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < a.kg) {
          const uint32_t e = (uint32_t)tid + 128u * i, m = e / (uint32_t)a.kg, k = e - m * (uint32_t)a.kg;
          float h, l;
          split_tf32(pre_s[i], h, l);
          const uint32_t off = sw128(m, k);
          *reinterpret_cast<float*>(hi_p + off) = h;
          *reinterpret_cast<float*>(lo_p + off) = l;
        }
      }
    }
  };

  uint32_t it = 0, tile_iter = 0;
  if ((int64_t)blockIdx.x < a.tiles) prefetch(blockIdx.x, 0);
  for (int64_t tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, ++tile_iter) {
    for (int j = 0; j < nka; ++j, ++it) {
      const int st = (int)(it & 1u);
      const uint32_t use = it >> 1;                              // how often this stage has been filled before
      if (use > 0) mbar_wait(st ? bar_stage1 : bar_stage0, (use - 1) & 1u);   // the MMAs that read it are done
      stage_store(st);
      fence_async_smem();
      tc_fence_before();                                         // the previous tile's tcgen05.ld before the barrier
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const int ksteps = ((a.kp - 32 * j) < 32 ? (a.kp - 32 * j) : 32) >> 3;
        const uint32_t a_hi = sA + (uint32_t)st * STAGE, a_lo = a_hi + HALF_STAGE;
        const uint32_t b_hi = sB + (uint32_t)j * ((uint32_t)a.n * 128u), b_lo = b_hi + b_half;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t dah = umma_desc(a_hi + 32u * ks), dal = umma_desc(a_lo + 32u * ks);
          const uint64_t dbh = umma_desc(b_hi + 32u * ks), dbl = umma_desc(b_lo + 32u * ks);
          umma_tf32(tmem_d, dah, dbh, idesc, (j | ks) ? 1u : 0u);
          umma_tf32(tmem_d, dal, dbh, idesc, 1u);
          umma_tf32(tmem_d, dah, dbl, idesc, 1u);
          umma_tf32(tmem_d, dal, dbl, idesc, 1u);
        }
        umma_commit(st ? bar_stage1 : bar_stage0);
        if (j == nka - 1) umma_commit(bar_acc);
      }
      __syncwarp();                                              // warp 0 converges again before any .sync.aligned instruction
      // next chunk's loads go out now: in flight during the MMAs and the epilogue
      {
        int jn = j + 1;
        int64_t tn = tile;
        if (jn == nka) { jn = 0; tn = tile + gridDim.x; }
        if (tn < a.tiles) prefetch(tn, jn);
      }
      if (j == nka - 1) {
        mbar_wait(bar_acc, tile_iter & 1u);
        tc_fence_after();
        const int64_t row = tile * TILE_M + warp * 32 + lane;
        float* const yr = a.y + (size_t)row * a.n;
        for (int c0 = 0; c0 < a.n; c0 += 16) {
          float v[16];
          tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
          if (row < a.rows) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              if (a.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + c0) + q);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              *reinterpret_cast<float4*>(yr + c0 + 4 * q) = o;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(a.tmem_cols) : "memory");
}

size_t smem_bytes(int kp, int n) { return 1024 + A_BYTES + 2 * (size_t)((kp + 31) / 32) * n * 128 + 64; }
}  // namespace

extern "C" int f16_lma_linear_supported(int in_features, int out_features) {
  if (out_features < 16 || out_features > 256 || out_features % 16) return 0;
  if (in_features <= 0) return 0;
  if (in_features > 32 && in_features % 32) return 0;
  const int kp = (in_features + 7) / 8 * 8;
  return smem_bytes(kp, out_features) <= 227 * 1024 ? 1 : 0;
}

extern "C" int f16_lma_linear_forward(int64_t rows, int in_features, int out_features, const float* x, const float* weight,
                                      const float* bias, float* y, void* stream) {
  if (rows <= 0) return f16_internal_fail("f16_lma_linear_forward: rows must be positive");
  if (!x || !weight || !y) return f16_internal_fail("f16_lma_linear_forward: NULL pointer");
  if (!f16_lma_linear_supported(in_features, out_features))
    return f16_internal_fail("f16_lma_linear_forward: unsupported shape (out features a multiple of 16 up to 256; in features <= 32 or a multiple of 32)");
  if ((((uintptr_t)y) & 15) || (bias && (((uintptr_t)bias) & 15))) return f16_internal_fail("f16_lma_linear_forward: y and bias must be 16-byte aligned");
  LinArgs a;
  a.x = x; a.w = weight; a.bias = bias; a.y = y;
  a.rows = rows; a.tiles = (rows + TILE_M - 1) / TILE_M;
  a.kg = in_features; a.kp = (in_features + 7) / 8 * 8; a.n = out_features;
  a.tmem_cols = 32;
  while ((int)a.tmem_cols < out_features) a.tmem_cols *= 2;
  const bool vec = in_features % 32 == 0 && (((uintptr_t)x) & 15) == 0;
  if (!vec && in_features > 32) return f16_internal_fail("f16_lma_linear_forward: x must be 16-byte aligned");
  const size_t smem = smem_bytes(a.kp, a.n);
  auto kern = vec ? linear_tc_kernel<true> : linear_tc_kernel<false>;
  static bool attr_done[2] = {false, false};
  if (!attr_done[vec]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return f16_internal_fail("f16_lma_linear_forward: cannot raise the shared-memory limit");
    attr_done[vec] = true;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem) != cudaSuccess || per_sm < 1)
    return f16_internal_fail("f16_lma_linear_forward: the kernel does not fit an SM");
  const int by_tmem = 512 / (int)a.tmem_cols;                     // tensor memory: 512 columns per SM
  if (per_sm > by_tmem) per_sm = by_tmem;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > a.tiles) grid = a.tiles;
  kern<<<(unsigned)grid, THREADS, smem, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
