// f16_lma_norm.cu - LayerNorm over the 32 channels of the LMA extractor's latent tokens (include/f16_lma.h).
// Rows are 128 bytes: a warp takes 32 rows (4 KB, contiguous) with eight fully coalesced float4 loads per lane,
// so each row lives in eight neighbouring lanes (four columns each) and its mean / variance are two 3-step
// shuffle reductions. HBM-bound: 128 B in + 128 B out per row forward; x, dy in and dx out backward, with the
// weight / bias gradients accumulated in registers over the warp's tiles and reduced once per CTA.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_lma.h"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int D = 32;              // channels per row (d_new = embed_dim / 2 of train.py:21-32; LayerNorm, jsbsim_gym/LMA_features.py:172-185)
constexpr int BLOCK = 256;         // 8 warps
constexpr int TILE_ROWS = 32;      // rows per warp tile

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// statistics of the row this lane's float4 belongs to (two-pass: mean, then centred second moment)
__device__ __forceinline__ void row_stats(const float4& v, float eps, float& mean, float& rstd) {
  mean = group8_sum((v.x + v.y) + (v.z + v.w)) * (1.0f / D);
  const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
  const float var = group8_sum((a * a + b * b) + (c * c + d * d)) * (1.0f / D);
  rstd = rsqrtf(var + eps);
}

__global__ void __launch_bounds__(BLOCK) lma_layernorm_fwd_kernel(int64_t rows, const float4* __restrict__ x, const float4* __restrict__ w,
                                                                  const float4* __restrict__ b, float eps, float4* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t warps = (int64_t)gridDim.x * (BLOCK / 32);
  const float4 w4 = w[lane & 7];
  const float4 b4 = b ? b[lane & 7] : make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  for (int64_t t = warp; t < tiles; t += warps) {
    const int64_t base = t * (TILE_ROWS * D / 4);          // float4 index of the tile
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t row = t * TILE_ROWS + i * 4 + (lane >> 3);
      v[i] = row < rows ? x[base + i * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float mean, rstd;
      row_stats(v[i], eps, mean, rstd);
      const int64_t row = t * TILE_ROWS + i * 4 + (lane >> 3);
      float4 o;
      o.x = (v[i].x - mean) * rstd * w4.x + b4.x;
      o.y = (v[i].y - mean) * rstd * w4.y + b4.y;
      o.z = (v[i].z - mean) * rstd * w4.z + b4.z;
      o.w = (v[i].w - mean) * rstd * w4.w + b4.w;
      if (row < rows) y[base + i * 32 + lane] = o;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w; dw += dy * xhat; db += dy
__global__ void __launch_bounds__(BLOCK) lma_layernorm_bwd_kernel(int64_t rows, const float4* __restrict__ x, const float4* __restrict__ w,
                                                                  const float4* __restrict__ dy, float eps, float4* __restrict__ dx,
                                                                  float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[2][BLOCK / 32][D];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * (BLOCK / 32) + wib;
  const int64_t warps = (int64_t)gridDim.x * (BLOCK / 32);
  const float4 w4 = w[lane & 7];
  float4 aw = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  for (int64_t t = warp; t < tiles; t += warps) {
    const int64_t base = t * (TILE_ROWS * D / 4);
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int64_t row = t * TILE_ROWS + i * 4 + (lane >> 3);
      const bool live = row < rows;
      const float4 v = live ? x[base + i * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 g = live ? dy[base + i * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
      float mean, rstd;
      row_stats(v, eps, mean, rstd);
      const float hx = (v.x - mean) * rstd, hy = (v.y - mean) * rstd, hz = (v.z - mean) * rstd, hw = (v.w - mean) * rstd;
      const float gx = g.x * w4.x, gy = g.y * w4.y, gz = g.z * w4.z, gw = g.w * w4.w;
      const float m1 = group8_sum((gx + gy) + (gz + gw)) * (1.0f / D);
      const float m2 = group8_sum((gx * hx + gy * hy) + (gz * hz + gw * hw)) * (1.0f / D);
      float4 o;
      o.x = rstd * (gx - m1 - hx * m2);
      o.y = rstd * (gy - m1 - hy * m2);
      o.z = rstd * (gz - m1 - hz * m2);
      o.w = rstd * (gw - m1 - hw * m2);
      if (live) dx[base + i * 32 + lane] = o;
      aw.x += g.x * hx; aw.y += g.y * hy; aw.z += g.z * hz; aw.w += g.w * hw;      // dead rows contribute zeros
      ab.x += g.x; ab.y += g.y; ab.z += g.z; ab.w += g.w;
    }
  }
  // lanes l, l+8, l+16, l+24 hold the same four columns
#pragma unroll
  for (int s = 8; s <= 16; s <<= 1) {
    aw.x += __shfl_xor_sync(0xffffffffu, aw.x, s); aw.y += __shfl_xor_sync(0xffffffffu, aw.y, s);
    aw.z += __shfl_xor_sync(0xffffffffu, aw.z, s); aw.w += __shfl_xor_sync(0xffffffffu, aw.w, s);
    ab.x += __shfl_xor_sync(0xffffffffu, ab.x, s); ab.y += __shfl_xor_sync(0xffffffffu, ab.y, s);
    ab.z += __shfl_xor_sync(0xffffffffu, ab.z, s); ab.w += __shfl_xor_sync(0xffffffffu, ab.w, s);
  }
  if (lane < 8) {
    float* rw = &red[0][wib][lane * 4];
    float* rb = &red[1][wib][lane * 4];
    rw[0] = aw.x; rw[1] = aw.y; rw[2] = aw.z; rw[3] = aw.w;
    rb[0] = ab.x; rb[1] = ab.y; rb[2] = ab.z; rb[3] = ab.w;
  }
  __syncthreads();
  if (threadIdx.x < 2 * D) {
    const int which = threadIdx.x / D, col = threadIdx.x % D;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < BLOCK / 32; ++k) s += red[which][k][col];
    float* dst = which == 0 ? dw : db;
    if (dst) atomicAdd(dst + col, s);
  }
}

int finish() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
unsigned grid_for(int64_t rows) {
  const int64_t tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t ctas = (tiles + BLOCK / 32 - 1) / (BLOCK / 32);
  const int64_t cap = 148 * 8;                      // grid-stride beyond eight CTAs per SM
  return (unsigned)(ctas < cap ? ctas : cap);
}
int check(const char* who, int64_t rows, int dim, const void* a, const void* b, const void* c) {
  if (rows <= 0) return f16_internal_fail("f16_lma_layernorm: rows must be positive");
  if (dim != D) return f16_internal_fail("f16_lma_layernorm: only 32-channel rows are built (the LMA configuration of train.py:21-32)");
  if (!a || !b || !c) return f16_internal_fail("f16_lma_layernorm: NULL pointer");
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) != 0) return f16_internal_fail("f16_lma_layernorm: pointers must be 16-byte aligned");
  (void)who;
  return 0;
}
}  // namespace

extern "C" int f16_lma_layernorm_forward(int64_t rows, int dim, const float* x, const float* weight, const float* bias, float eps, float* y,
                                         void* stream) {
  if (int rc = check("forward", rows, dim, x, weight, y)) return rc;
  lma_layernorm_fwd_kernel<<<grid_for(rows), BLOCK, 0, (cudaStream_t)stream>>>(rows, (const float4*)x, (const float4*)weight, (const float4*)bias,
                                                                               eps, (float4*)y);
  return finish();
}
extern "C" int f16_lma_layernorm_backward(int64_t rows, int dim, const float* x, const float* weight, const float* dy, float eps, float* dx,
                                          float* dweight, float* dbias, void* stream) {
  if (int rc = check("backward", rows, dim, x, weight, dy)) return rc;
  if (!dx || !dweight) return f16_internal_fail("f16_lma_layernorm_backward: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dweight, 0, D * sizeof(float), st) != cudaSuccess) return f16_internal_fail("f16_lma_layernorm_backward: memset failed");
  if (dbias && cudaMemsetAsync(dbias, 0, D * sizeof(float), st) != cudaSuccess) return f16_internal_fail("f16_lma_layernorm_backward: memset failed");
  lma_layernorm_bwd_kernel<<<grid_for(rows), BLOCK, 0, st>>>(rows, (const float4*)x, (const float4*)weight, (const float4*)dy, eps, (float4*)dx,
                                                             dweight, dbias);
  return finish();
}
