// f16_lma_wgrad.cu - weight and bias gradients of the policy's small Linear layers (include/f16_lma.h).
//
// dW[out][in] = sum_rows dY[row][out] * X[row][in], db[out] = sum_rows dY[row][out] with out, in <= 160 and 10^5..10^6
// rows: a GEMM whose reduction axis is the batch. Library SGEMM kernels give such a shape to a handful of CTAs
// (measured: 480 us for 128 x 32 over 655 360 rows, 0.9 TB/s; one third of an AM-PPO update) and the bias gradient
// is a second pass over dY. Here the rows are cut in slabs, one CTA per (slab, 16*TO x 16*TI output tile): the
// slab streams through shared memory in chunks of 32 rows (cp.async, two stages), every thread keeps a TO x TI
// register tile, the first 16*TO threads also the column sums of dY, and the slabs meet in float atomics on the
// zero-initialised outputs. FP32 FMA, no tensor
// cores: the reference computes in FP32 and the kernel is bound by reading X and dY once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_lma.h"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int THREADS = 256, RC = 32;

template <int N> struct Vec;
template <> struct Vec<1> { static __device__ __forceinline__ void ld(float* d, const float* s) { d[0] = s[0]; } };
template <> struct Vec<2> { static __device__ __forceinline__ void ld(float* d, const float* s) { const float2 v = *reinterpret_cast<const float2*>(s); d[0] = v.x; d[1] = v.y; } };
template <> struct Vec<4> { static __device__ __forceinline__ void ld(float* d, const float* s) { const float4 v = *reinterpret_cast<const float4*>(s); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; } };
template <> struct Vec<8> { static __device__ __forceinline__ void ld(float* d, const float* s) { Vec<4>::ld(d, s); Vec<4>::ld(d + 4, s + 4); } };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool live) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = live ? 16 : 0;                      // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// VEC: rows of x and dy are whole float4s (in, out multiples of 4, 16-byte aligned bases): chunks arrive by cp.async
// into a two-stage ring, so the next 32 rows are in flight while the current ones are multiplied. Otherwise (17 input
// features, a single output) a plain load / store staging with one stage.
template <int TO, int TI, bool VEC>
__global__ void __launch_bounds__(THREADS) linear_wgrad_kernel(int64_t rows, int in, int out, const float* __restrict__ x,
                                                               const float* __restrict__ dy, float* __restrict__ dw, float* __restrict__ db) {
  constexpr int OT = 16 * TO, IT = 16 * TI, STAGES = VEC ? 2 : 1;
  __shared__ __align__(16) float Xs[STAGES][RC][IT];
  __shared__ __align__(16) float Ys[STAGES][RC][OT];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int o0 = blockIdx.y * OT, i0 = blockIdx.z * IT;
  const int64_t per = ((rows + gridDim.x - 1) / gridDim.x + RC - 1) / RC * RC;
  const int64_t r0 = (int64_t)blockIdx.x * per;
  const int64_t r1 = r0 + per < rows ? r0 + per : rows;
  if (r0 >= r1) return;
  float acc[TO][TI];
#pragma unroll
  for (int j = 0; j < TO; ++j)
#pragma unroll
    for (int k = 0; k < TI; ++k) acc[j][k] = 0.f;
  float accb = 0.f;                                   // column sum of dy for output o0 + threadIdx.x (threads < OT)

  auto stage_in = [&](int st, int64_t rb) {
    if (VEC) {
      constexpr int XV = IT / 4, YV = OT / 4;         // float4s per staged row
      for (int e = threadIdx.x; e < RC * XV; e += THREADS) {
        const int r = e / XV, c = (e % XV) * 4;
        const int64_t row = rb + r;
        const bool live = row < r1 && i0 + c < in;
        cp_async16(&Xs[st][r][c], live ? x + row * in + i0 + c : x, live);
      }
      for (int e = threadIdx.x; e < RC * YV; e += THREADS) {
        const int r = e / YV, c = (e % YV) * 4;
        const int64_t row = rb + r;
        const bool live = row < r1 && o0 + c < out;
        cp_async16(&Ys[st][r][c], live ? dy + row * out + o0 + c : dy, live);
      }
      cp_async_commit();
    } else {
      for (int e = threadIdx.x; e < RC * IT; e += THREADS) {
        const int r = e / IT, c = e % IT;
        const int64_t row = rb + r;
        Xs[st][r][c] = (row < r1 && i0 + c < in) ? x[row * in + i0 + c] : 0.f;
      }
      for (int e = threadIdx.x; e < RC * OT; e += THREADS) {
        const int r = e / OT, c = e % OT;
        const int64_t row = rb + r;
        Ys[st][r][c] = (row < r1 && o0 + c < out) ? dy[row * out + o0 + c] : 0.f;
      }
    }
  };

  int st = 0;
  stage_in(0, r0);
  for (int64_t rb = r0; rb < r1; rb += RC) {
    if (VEC) {
      if (rb + RC < r1) { stage_in(st ^ 1, rb + RC); cp_async_wait<1>(); }
      else cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < RC; ++r) {
      float a[TO], b[TI];
      Vec<TO>::ld(a, &Ys[st][r][ty * TO]);
      Vec<TI>::ld(b, &Xs[st][r][tx * TI]);
#pragma unroll
      for (int j = 0; j < TO; ++j)
#pragma unroll
        for (int k = 0; k < TI; ++k) acc[j][k] = fmaf(a[j], b[k], acc[j][k]);
    }
    if (db && blockIdx.z == 0 && threadIdx.x < OT) {   // whole warps (OT is a multiple of 16; the odd half-warp idles)
#pragma unroll 8
      for (int r = 0; r < RC; ++r) accb += Ys[st][r][threadIdx.x];
    }
    __syncthreads();
    if (VEC) st ^= 1;
    else if (rb + RC < r1) stage_in(0, rb + RC);
  }
#pragma unroll
  for (int j = 0; j < TO; ++j) {
    const int o = o0 + ty * TO + j;
    if (o >= out) continue;
#pragma unroll
    for (int k = 0; k < TI; ++k) {
      const int i = i0 + tx * TI + k;
      if (i < in) atomicAdd(dw + (size_t)o * in + i, acc[j][k]);
    }
  }
  if (db && blockIdx.z == 0 && threadIdx.x < OT && o0 + (int)threadIdx.x < out) atomicAdd(db + o0 + threadIdx.x, accb);
}

template <int TO, int TI>
void launch(int64_t rows, int in, int out, const float* x, const float* dy, float* dw, float* db, cudaStream_t st) {
  constexpr int OT = 16 * TO, IT = 16 * TI;
  const unsigned ty = (unsigned)((out + OT - 1) / OT), tz = (unsigned)((in + IT - 1) / IT);
  // slabs: enough CTAs to fill the machine a few times over, at least 8 chunks of rows each
  int64_t slabs = (148 * 8) / (int64_t)(ty * tz);
  const int64_t max_slabs = (rows + 8 * RC - 1) / (8 * RC);
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  const bool vec = in % 4 == 0 && out % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)dy)) & 15) == 0;
  if (vec) linear_wgrad_kernel<TO, TI, true><<<dim3((unsigned)slabs, ty, tz), THREADS, 0, st>>>(rows, in, out, x, dy, dw, db);
  else linear_wgrad_kernel<TO, TI, false><<<dim3((unsigned)slabs, ty, tz), THREADS, 0, st>>>(rows, in, out, x, dy, dw, db);
}
}  // namespace

extern "C" int f16_lma_linear_wgrad(int64_t rows, int in_features, int out_features, const float* x, const float* dy, float* dweight,
                                    float* dbias, void* stream) {
  if (rows <= 0 || in_features <= 0 || out_features <= 0) return f16_internal_fail("f16_lma_linear_wgrad: rows, in_features and out_features must be positive");
  if (!x || !dy || !dweight) return f16_internal_fail("f16_lma_linear_wgrad: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dweight, 0, (size_t)in_features * out_features * sizeof(float), st) != cudaSuccess)
    return f16_internal_fail("f16_lma_linear_wgrad: memset failed");
  if (dbias && cudaMemsetAsync(dbias, 0, (size_t)out_features * sizeof(float), st) != cudaSuccess)
    return f16_internal_fail("f16_lma_linear_wgrad: memset failed");
  // register tile by shape: wide outputs take 8 per thread, narrow inputs 1 or 2
  const bool wide_out = out_features > 64, tiny_out = out_features <= 16;
  const int ti = in_features <= 16 ? 1 : in_features <= 32 ? 2 : 4;
#define F16_WGRAD(TO)                                                                                        \
  do {                                                                                                      \
    if (ti == 1) launch<TO, 1>(rows, in_features, out_features, x, dy, dweight, dbias, st);                 \
    else if (ti == 2) launch<TO, 2>(rows, in_features, out_features, x, dy, dweight, dbias, st);            \
    else launch<TO, 4>(rows, in_features, out_features, x, dy, dweight, dbias, st);                         \
  } while (0)
  if (wide_out) F16_WGRAD(8);
  else if (tiny_out) F16_WGRAD(1);
  else F16_WGRAD(4);
#undef F16_WGRAD
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
