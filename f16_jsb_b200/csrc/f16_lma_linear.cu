// f16_lma_linear.cu - forward of the policy's small Linear layers with the bias fused (include/f16_lma.h).
//
// y[row][o] = sum_k x[row][k] * w[o][k] + b[o] for 10^5..10^6 rows and K, O <= 160: a tall-skinny SGEMM. The library
// picks SIMT kernels plus a separate bias pass for these shapes (measured: 16 layers = 3.0 ms of an 11 ms AM-PPO
// update step, 0.9 ms of it the bias kernels). One CTA takes 16*TR rows x 16*TO outputs; x and w stream through
// shared memory in chunks of 32 along K (x transposed on the way, so that a thread's TR rows are one vector load);
// every thread keeps a TR x TO register tile and adds the bias when it writes. FP32 FMA: the reference computes
// in FP32, so no tensor cores here.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_lma.h"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int THREADS = 256, KC = 32;

template <int N> struct V;
template <> struct V<4> {
  static __device__ __forceinline__ void ld(float* d, const float* s) { const float4 v = *reinterpret_cast<const float4*>(s); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
};
template <> struct V<8> {
  static __device__ __forceinline__ void ld(float* d, const float* s) { V<4>::ld(d, s); V<4>::ld(d + 4, s + 4); }
};

template <int TR, int TO>
__global__ void __launch_bounds__(THREADS) linear_fwd_kernel(int64_t rows, int K, int O, const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y) {
  constexpr int RT = 16 * TR, OT = 16 * TO, XP = RT + 4;      // padded pitch of the transposed x chunk (keeps 16-byte alignment)
  __shared__ __align__(16) float Xs[KC][XP];
  __shared__ __align__(16) float Ws[KC][OT];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * RT;
  const int o0 = blockIdx.y * OT;
  float acc[TR][TO];
#pragma unroll
  for (int i = 0; i < TR; ++i)
#pragma unroll
    for (int j = 0; j < TO; ++j) acc[i][j] = 0.f;
  for (int kc = 0; kc < K; kc += KC) {
    // x chunk: RT rows x KC columns, read along K (coalesced), stored transposed
#pragma unroll 4
    for (int e = threadIdx.x; e < RT * KC; e += THREADS) {
      const int r = e / KC, c = e % KC;
      const int64_t row = row0 + r;
      Xs[c][r] = (row < rows && kc + c < K) ? x[row * K + kc + c] : 0.f;
    }
    // w chunk: OT outputs x KC columns of w[o][k], read along K, stored as Ws[k][o]
#pragma unroll 4
    for (int e = threadIdx.x; e < OT * KC; e += THREADS) {
      const int o = e / KC, c = e % KC;
      Ws[c][o] = (o0 + o < O && kc + c < K) ? w[(size_t)(o0 + o) * K + kc + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
      float a[TR], b[TO];
      V<TR>::ld(a, &Xs[k][ty * TR]);
      V<TO>::ld(b, &Ws[k][tx * TO]);
#pragma unroll
      for (int i = 0; i < TR; ++i)
#pragma unroll
        for (int j = 0; j < TO; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float bj[TO];
#pragma unroll
  for (int j = 0; j < TO; ++j) { const int o = o0 + tx * TO + j; bj[j] = (bias && o < O) ? bias[o] : 0.f; }
  const bool vec = (O % 4) == 0;                       // o0 and tx * TO are multiples of 4: whole float4s when O is
#pragma unroll
  for (int i = 0; i < TR; ++i) {
    const int64_t row = row0 + ty * TR + i;
    if (row >= rows) continue;
    float* dst = y + row * O + o0 + tx * TO;
#pragma unroll
    for (int j = 0; j < TO; j += 4) {
      const int o = o0 + tx * TO + j;
      if (vec && o + 3 < O) {
        *reinterpret_cast<float4*>(dst + j) = make_float4(acc[i][j] + bj[j], acc[i][j + 1] + bj[j + 1], acc[i][j + 2] + bj[j + 2], acc[i][j + 3] + bj[j + 3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (o + q < O) dst[j + q] = acc[i][j + q] + bj[j + q];
      }
    }
  }
}

template <int TR, int TO>
void launch(int64_t rows, int K, int O, const float* x, const float* w, const float* b, float* y, cudaStream_t st) {
  constexpr int RT = 16 * TR, OT = 16 * TO;
  const dim3 grid((unsigned)((rows + RT - 1) / RT), (unsigned)((O + OT - 1) / OT));
  linear_fwd_kernel<TR, TO><<<grid, THREADS, 0, st>>>(rows, K, O, x, w, b, y);
}
}  // namespace

extern "C" int f16_lma_linear_forward(int64_t rows, int in_features, int out_features, const float* x, const float* weight, const float* bias,
                                      float* y, void* stream) {
  if (rows <= 0 || in_features <= 0 || out_features <= 0) return f16_internal_fail("f16_lma_linear_forward: rows, in_features and out_features must be positive");
  if (!x || !weight || !y) return f16_internal_fail("f16_lma_linear_forward: NULL pointer");
  if ((((uintptr_t)y) & 15) != 0) return f16_internal_fail("f16_lma_linear_forward: y must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (out_features > 64) launch<8, 8>(rows, in_features, out_features, x, weight, bias, y, st);       // 128 rows x 128 outputs per CTA
  else launch<8, 4>(rows, in_features, out_features, x, weight, bias, y, st);                          // 128 rows x 64 outputs
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
