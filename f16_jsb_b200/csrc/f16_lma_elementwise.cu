// f16_lma_elementwise.cu - the two fused elementwise steps of the LMA extractor's training forward / backward
// (include/f16_lma.h).
//
//   embed activation   y = dropout(relu(a) + pos[row mod T])      (_InitialTransform, jsbsim_gym/LMA_features.py:221-279:
//                                                                  ReLU, sinusoidal positions, embedding dropout)
//   residual dropout   y = z + dropout(x)                          (LMA block: z + drop(attn(..)), z + drop(mlp(..)), :386-407)
//
// torch runs them as three and two kernels with a stored byte mask (6.25 and 5.25 passes over the tensor, backward 5.25 and
// 2.25); here each is one pass forward (read, read, write) and one backward, and the keep mask is regenerated from
// Philox4x32-10 keyed by (seed, group of eight elements) instead of being stored. On the embedding's 1.3 M x 64 activations
// (335 MB) that is ~0.35 ms per AM-PPO minibatch step. Elementwise, bound by HBM; float4 accesses, grid-stride.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_lma.h"
#include "f16_model.cuh"      // philox4x32_10

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
struct Keep {
  uint32_t thr;      // an element is dropped when its 16-bit draw < thr
  float scale;       // 65536 / (65536 - thr)
};
Keep make_keep(float p) {
  Keep k;
  long t = lroundf(p * 65536.0f);
  k.thr = (uint32_t)(t < 0 ? 0 : (t > 65535 ? 65535 : t));
  k.scale = 65536.0f / (65536.0f - (float)k.thr);
  return k;
}
// keep factors (0 or scale) of elements 8 g .. 8 g + 7
__device__ __forceinline__ void keep8(uint64_t seed, uint64_t g, Keep k, float* m) {
  uint32_t w[4];
  f16::philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), 0x454cu, 0x4d57u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = (((w[i >> 1] >> (16 * (i & 1))) & 0xffffu) < k.thr) ? 0.0f : k.scale;
}

// MODE 0: y = (relu(a) + pos) * keep      MODE 1: da = dy * keep * (a > 0)      (C % 8 == 0: a group never straddles rows)
// With H > 1 heads, y (and dy) are laid out head-stacked: element (b, t, h * C/H + c) lives at b*T*C + h * (T * C/H) + t * C/H + c,
// the permutation LMA_features.py:255-270 applies before re-chunking the T*C values into latent tokens ((C/H) % 8 == 0: a group
// stays contiguous). The keep mask is keyed by the natural index (b, t, channel) either way.
template <int MODE>
__global__ void __launch_bounds__(256) embed_act_kernel(int64_t groups, int C, int T, int H, const float4* __restrict__ a, const float4* __restrict__ pos,
                                                        const float4* __restrict__ dy, float4* __restrict__ out, Keep k, uint64_t seed) {
  const int gpr = C >> 3;                                   // groups per row
  const int gph = gpr / H;                                  // groups per head
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    float m[8];
    keep8(seed, (uint64_t)g, k, m);
    const float4 a0 = __ldg(a + 2 * g), a1 = __ldg(a + 2 * g + 1);
    float4 o0, o1;
    const int64_t row = g / gpr;
    const int c = (int)(g - row * gpr), t = (int)(row % T);
    int64_t gs = g;                                         // group index in the stacked layout
    if (H > 1) gs = (row - t) * gpr + (int64_t)(c / gph) * (T * gph) + (int64_t)t * gph + (c % gph);
    if (MODE == 0) {
      const float4 p0 = __ldg(pos + (size_t)t * (C >> 2) + 2 * c), p1 = __ldg(pos + (size_t)t * (C >> 2) + 2 * c + 1);
      o0 = make_float4((fmaxf(a0.x, 0.f) + p0.x) * m[0], (fmaxf(a0.y, 0.f) + p0.y) * m[1], (fmaxf(a0.z, 0.f) + p0.z) * m[2], (fmaxf(a0.w, 0.f) + p0.w) * m[3]);
      o1 = make_float4((fmaxf(a1.x, 0.f) + p1.x) * m[4], (fmaxf(a1.y, 0.f) + p1.y) * m[5], (fmaxf(a1.z, 0.f) + p1.z) * m[6], (fmaxf(a1.w, 0.f) + p1.w) * m[7]);
    } else {
      const float4 d0 = __ldg(dy + 2 * gs), d1 = __ldg(dy + 2 * gs + 1);
      o0 = make_float4(a0.x > 0.f ? d0.x * m[0] : 0.f, a0.y > 0.f ? d0.y * m[1] : 0.f, a0.z > 0.f ? d0.z * m[2] : 0.f, a0.w > 0.f ? d0.w * m[3] : 0.f);
      o1 = make_float4(a1.x > 0.f ? d1.x * m[4] : 0.f, a1.y > 0.f ? d1.y * m[5] : 0.f, a1.z > 0.f ? d1.z * m[6] : 0.f, a1.w > 0.f ? d1.w * m[7] : 0.f);
    }
    const int64_t go = MODE == 0 ? gs : g;
    out[2 * go] = o0;
    out[2 * go + 1] = o1;
  }
}

// MODE 0: y = z + x * keep      MODE 1: dx = dy * keep      (n % 8 == 0)
template <int MODE>
__global__ void __launch_bounds__(256) dropout_add_kernel(int64_t groups, const float4* __restrict__ x, const float4* __restrict__ z,
                                                          float4* __restrict__ out, Keep k, uint64_t seed) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    float m[8];
    keep8(seed, (uint64_t)g, k, m);
    const float4 x0 = __ldg(x + 2 * g), x1 = __ldg(x + 2 * g + 1);
    float4 o0 = make_float4(x0.x * m[0], x0.y * m[1], x0.z * m[2], x0.w * m[3]);
    float4 o1 = make_float4(x1.x * m[4], x1.y * m[5], x1.z * m[6], x1.w * m[7]);
    if (MODE == 0) {
      const float4 z0 = __ldg(z + 2 * g), z1 = __ldg(z + 2 * g + 1);
      o0.x += z0.x; o0.y += z0.y; o0.z += z0.z; o0.w += z0.w;
      o1.x += z1.x; o1.y += z1.y; o1.z += z1.z; o1.w += z1.w;
    }
    out[2 * g] = o0;
    out[2 * g + 1] = o1;
  }
}

unsigned grid_for(int64_t groups) {
  int64_t b = (groups + 255) / 256;
  const int64_t cap = 148 * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}
bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
int finish(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  (void)what;
  f16_internal_count_launch();
  return 0;
}
}  // namespace

namespace {
const char* bad_stacking(int64_t rows, int channels, int seq_len, int stack_heads) {
  if (stack_heads < 1) return "stack_heads must be >= 1";
  if (stack_heads > 1 && (channels % stack_heads || (channels / stack_heads) % 8 || rows % seq_len))
    return "head stacking needs channels / stack_heads a multiple of 8 and rows a multiple of seq_len";
  return nullptr;
}
}  // namespace

extern "C" int f16_lma_embed_act_forward(int64_t rows, int channels, int seq_len, int stack_heads, const float* a, const float* pos, float* y,
                                         float dropout_p, uint64_t seed, void* stream) {
  if (rows <= 0 || channels <= 0 || channels % 8 || seq_len <= 0) return f16_internal_fail("f16_lma_embed_act_forward: rows > 0, channels a positive multiple of 8, seq_len > 0");
  if (bad_stacking(rows, channels, seq_len, stack_heads)) return f16_internal_fail(bad_stacking(rows, channels, seq_len, stack_heads));
  if (!a || !pos || !y || !aligned16(a) || !aligned16(pos) || !aligned16(y)) return f16_internal_fail("f16_lma_embed_act_forward: pointers must be non-NULL and 16-byte aligned");
  if (!(dropout_p >= 0.0f && dropout_p < 1.0f)) return f16_internal_fail("f16_lma_embed_act_forward: dropout_p must be in [0, 1)");
  const int64_t groups = rows * (channels / 8);
  embed_act_kernel<0><<<grid_for(groups), 256, 0, (cudaStream_t)stream>>>(groups, channels, seq_len, stack_heads, (const float4*)a, (const float4*)pos, nullptr, (float4*)y,
                                                                          make_keep(dropout_p), seed);
  return finish("f16_lma_embed_act_forward");
}

extern "C" int f16_lma_embed_act_backward(int64_t rows, int channels, int seq_len, int stack_heads, const float* a, const float* dy, float* da,
                                          float dropout_p, uint64_t seed, void* stream) {
  if (rows <= 0 || channels <= 0 || channels % 8 || seq_len <= 0) return f16_internal_fail("f16_lma_embed_act_backward: rows > 0, channels a positive multiple of 8, seq_len > 0");
  if (bad_stacking(rows, channels, seq_len, stack_heads)) return f16_internal_fail(bad_stacking(rows, channels, seq_len, stack_heads));
  if (!a || !dy || !da || !aligned16(a) || !aligned16(dy) || !aligned16(da)) return f16_internal_fail("f16_lma_embed_act_backward: pointers must be non-NULL and 16-byte aligned");
  if (!(dropout_p >= 0.0f && dropout_p < 1.0f)) return f16_internal_fail("f16_lma_embed_act_backward: dropout_p must be in [0, 1)");
  const int64_t groups = rows * (channels / 8);
  embed_act_kernel<1><<<grid_for(groups), 256, 0, (cudaStream_t)stream>>>(groups, channels, seq_len, stack_heads, (const float4*)a, nullptr, (const float4*)dy, (float4*)da,
                                                                          make_keep(dropout_p), seed);
  return finish("f16_lma_embed_act_backward");
}

extern "C" int f16_lma_dropout_add_forward(int64_t n, const float* x, const float* z, float* y, float dropout_p, uint64_t seed, void* stream) {
  if (n <= 0 || n % 8) return f16_internal_fail("f16_lma_dropout_add_forward: n must be a positive multiple of 8");
  if (!x || !z || !y || !aligned16(x) || !aligned16(z) || !aligned16(y)) return f16_internal_fail("f16_lma_dropout_add_forward: pointers must be non-NULL and 16-byte aligned");
  if (!(dropout_p >= 0.0f && dropout_p < 1.0f)) return f16_internal_fail("f16_lma_dropout_add_forward: dropout_p must be in [0, 1)");
  dropout_add_kernel<0><<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>(n / 8, (const float4*)x, (const float4*)z, (float4*)y, make_keep(dropout_p), seed);
  return finish("f16_lma_dropout_add_forward");
}

extern "C" int f16_lma_dropout_backward(int64_t n, const float* dy, float* dx, float dropout_p, uint64_t seed, void* stream) {
  if (n <= 0 || n % 8) return f16_internal_fail("f16_lma_dropout_backward: n must be a positive multiple of 8");
  if (!dy || !dx || !aligned16(dy) || !aligned16(dx)) return f16_internal_fail("f16_lma_dropout_backward: pointers must be non-NULL and 16-byte aligned");
  if (!(dropout_p >= 0.0f && dropout_p < 1.0f)) return f16_internal_fail("f16_lma_dropout_backward: dropout_p must be in [0, 1)");
  dropout_add_kernel<1><<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>(n / 8, (const float4*)dy, nullptr, (float4*)dx, make_keep(dropout_p), seed);
  return finish("f16_lma_dropout_backward");
}
