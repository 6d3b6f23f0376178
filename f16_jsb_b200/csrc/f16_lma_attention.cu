// f16_lma_attention.cu - attention over the LMA extractor's five latent tokens (include/f16_lma.h): one thread
// per (sample, head), everything in registers, HBM-bound (160 floats in + 40 out per thread forward).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_lma.h"
#include "f16_model.cuh"     // philox4x32_10

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int T = 5, DH = 8;      // 25 sixteen-bit dropout draws = 13 words = 4 Philox blocks

struct Drop {
  uint32_t thr;      // an attention weight is dropped when its 16-bit draw < thr
  float scale;       // 65536 / (65536 - thr)
};
__host__ Drop make_drop(float p) {
  Drop d;
  long t = lroundf(p * 65536.0f);
  d.thr = (uint32_t)(t < 0 ? 0 : (t > 65535 ? 65535 : t));
  d.scale = 65536.0f / (65536.0f - (float)d.thr);
  return d;
}
// keep factors (0 or scale) of the 25 weights of one (sample, head): Philox4x32-10 keyed by the seed,
// counter = (sample*H + head, block)
__device__ __forceinline__ void keep_factors(uint64_t seed, uint64_t bh, Drop d, float (*m)[T]) {
  if (d.thr == 0) {
#pragma unroll
    for (int i = 0; i < T; ++i)
#pragma unroll
      for (int j = 0; j < T; ++j) m[i][j] = 1.0f;
    return;
  }
  uint32_t w[16];
#pragma unroll
  for (int blk = 0; blk < 4; ++blk)
    f16::philox4x32_10((uint32_t)bh, (uint32_t)(bh >> 32), (uint32_t)blk, 0x1A77u, (uint32_t)seed, (uint32_t)(seed >> 32), &w[4 * blk]);
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const int e = i * T + j;
      const uint32_t r = (w[e >> 1] >> (16 * (e & 1))) & 0xffffu;
      m[i][j] = r < d.thr ? 0.0f : d.scale;
    }
}
__device__ __forceinline__ void load8(float* dst, const float* src) {
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
}
__device__ __forceinline__ void store8(float* dst, const float* v) {
  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// softmax probabilities of query row q against the five keys
__device__ __forceinline__ void prob_row(const float* q, const float (*k)[DH], float* p) {
  const float scale = 0.35355339059327379f;      // 1 / sqrt(8)
  float mx = -3.0e38f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < DH; ++c) s = fmaf(q[c], k[j][c], s);
    p[j] = s * scale;
    mx = fmaxf(mx, p[j]);
  }
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < T; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < T; ++j) p[j] *= inv;
}

__global__ void __launch_bounds__(128) lma_attention_fwd_kernel(int64_t n_bh, int H, const float* __restrict__ qkv, float* __restrict__ y,
                                                                Drop drop, uint64_t seed) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_bh) return;
  const int64_t b = idx / H;
  const int h = (int)(idx - b * H), D = H * DH;
  const float* base = qkv + b * (int64_t)(T * 3 * D) + h * DH;
  float k[T][DH], v[T][DH], m[T][T];
#pragma unroll
  for (int t = 0; t < T; ++t) { load8(k[t], base + t * 3 * D + D); load8(v[t], base + t * 3 * D + 2 * D); }
  keep_factors(seed, (uint64_t)idx, drop, m);
  float* out = y + b * (int64_t)(T * D) + h * DH;
#pragma unroll
  for (int i = 0; i < T; ++i) {
    float q[DH], p[T], o[DH];
    load8(q, base + i * 3 * D);
    prob_row(q, k, p);
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = 0.0f;
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const float w = p[j] * m[i][j];
#pragma unroll
      for (int c = 0; c < DH; ++c) o[c] = fmaf(w, v[j][c], o[c]);
    }
    store8(out + i * D, o);
  }
}

__global__ void __launch_bounds__(128) lma_attention_bwd_kernel(int64_t n_bh, int H, const float* __restrict__ qkv, const float* __restrict__ dy,
                                                                float* __restrict__ dqkv, Drop drop, uint64_t seed) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_bh) return;
  const int64_t b = idx / H;
  const int h = (int)(idx - b * H), D = H * DH;
  const float scale = 0.35355339059327379f;
  const float* base = qkv + b * (int64_t)(T * 3 * D) + h * DH;
  const float* gy = dy + b * (int64_t)(T * D) + h * DH;
  float* gbase = dqkv + b * (int64_t)(T * 3 * D) + h * DH;
  float k[T][DH], v[T][DH], dk[T][DH], dv[T][DH], m[T][T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    load8(k[t], base + t * 3 * D + D);
    load8(v[t], base + t * 3 * D + 2 * D);
#pragma unroll
    for (int c = 0; c < DH; ++c) { dk[t][c] = 0.0f; dv[t][c] = 0.0f; }
  }
  keep_factors(seed, (uint64_t)idx, drop, m);
#pragma unroll
  for (int i = 0; i < T; ++i) {
    float q[DH], g[DH], p[T], dp[T], dq[DH];
    load8(q, base + i * 3 * D);
    load8(g, gy + i * D);
    prob_row(q, k, p);
    float dot = 0.0f;
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const float w = p[j] * m[i][j];                      // weight actually applied in the forward
      float gv = 0.0f;
#pragma unroll
      for (int c = 0; c < DH; ++c) { dv[j][c] = fmaf(w, g[c], dv[j][c]); gv = fmaf(g[c], v[j][c], gv); }
      dp[j] = gv * m[i][j];                                // d loss / d p_j
      dot = fmaf(dp[j], p[j], dot);
    }
#pragma unroll
    for (int c = 0; c < DH; ++c) dq[c] = 0.0f;
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const float ds = p[j] * (dp[j] - dot) * scale;       // d loss / d score_j, times the 1/sqrt(dh) of the score
#pragma unroll
      for (int c = 0; c < DH; ++c) { dq[c] = fmaf(ds, k[j][c], dq[c]); dk[j][c] = fmaf(ds, q[c], dk[j][c]); }
    }
    store8(gbase + i * 3 * D, dq);
  }
#pragma unroll
  for (int t = 0; t < T; ++t) { store8(gbase + t * 3 * D + D, dk[t]); store8(gbase + t * 3 * D + 2 * D, dv[t]); }
}

__global__ void lma_attention_mask_kernel(int64_t n_bh, float* __restrict__ mask, Drop drop, uint64_t seed) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_bh) return;
  float m[T][T];
  keep_factors(seed, (uint64_t)idx, drop, m);
  for (int i = 0; i < T; ++i)
    for (int j = 0; j < T; ++j) mask[idx * (T * T) + i * T + j] = m[i][j];
}

int check_shape(const char* who, int64_t batch, int seq_len, int heads, int head_dim, float p) {
  if (batch <= 0 || heads <= 0) return f16_internal_fail("f16_lma_attention: batch and heads must be positive");
  if (seq_len != T || head_dim != DH) return f16_internal_fail("f16_lma_attention: only seq_len 5 and head_dim 8 are built (the LMA configuration of train.py:21-32)");
  if (!(p >= 0.0f && p < 1.0f)) return f16_internal_fail("f16_lma_attention: dropout_p must be in [0, 1)");
  (void)who;
  return 0;
}
int finish() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
}  // namespace

extern "C" int f16_lma_attention_forward(int64_t batch, int seq_len, int heads, int head_dim, const float* qkv, float* y, float dropout_p,
                                         uint64_t seed, void* stream) {
  if (int rc = check_shape("forward", batch, seq_len, heads, head_dim, dropout_p)) return rc;
  if (!qkv || !y) return f16_internal_fail("f16_lma_attention_forward: NULL pointer");
  const int64_t n = batch * heads;
  lma_attention_fwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, heads, qkv, y, make_drop(dropout_p), seed);
  return finish();
}
extern "C" int f16_lma_attention_backward(int64_t batch, int seq_len, int heads, int head_dim, const float* qkv, const float* dy, float* dqkv,
                                          float dropout_p, uint64_t seed, void* stream) {
  if (int rc = check_shape("backward", batch, seq_len, heads, head_dim, dropout_p)) return rc;
  if (!qkv || !dy || !dqkv) return f16_internal_fail("f16_lma_attention_backward: NULL pointer");
  const int64_t n = batch * heads;
  lma_attention_bwd_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, heads, qkv, dy, dqkv, make_drop(dropout_p), seed);
  return finish();
}
extern "C" int f16_lma_attention_mask(int64_t batch, int seq_len, int heads, float* mask, float dropout_p, uint64_t seed, void* stream) {
  if (int rc = check_shape("mask", batch, seq_len, heads, DH, dropout_p)) return rc;
  if (!mask) return f16_internal_fail("f16_lma_attention_mask: NULL pointer");
  const int64_t n = batch * heads;
  lma_attention_mask_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, mask, make_drop(dropout_p), seed);
  return finish();
}
