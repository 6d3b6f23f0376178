// f16_features.cu - JSBSimFeatureExtractor.forward (jsbsim_gym/features.py:37-67) as one HBM-bound kernel:
// 60 B in, 68 B out per frame. Each warp moves 32 frames through shared memory so that both the global
// loads (15 x 128 B) and stores (17 x 128 B) are fully coalesced; rows of 15 / 17 floats are
// bank-conflict free (odd strides). Accurate libm sincosf / atan2f / sqrtf and IEEE division, to match the
// PyTorch float32 reference to 1-2 ulp.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_features.h"

extern "C" int f16_internal_fail(const char* msg);
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int FI = 15, FO = 17, WARPS = 8;

__global__ void __launch_bounds__(WARPS * 32) features17_kernel(int64_t n, const float* __restrict__ in, float* __restrict__ out) {
  __shared__ float s_in[WARPS][32 * FI];
  __shared__ float s_out[WARPS][32 * FO];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * WARPS;
  for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile * 32 < n; tile += warps_total) {
    const int64_t f0 = tile * 32;
    const int cnt = (int)((n - f0) < 32 ? (n - f0) : 32);
    const float* src = in + f0 * FI;
#pragma unroll
    for (int k = 0; k < FI; ++k) {
      int i = k * 32 + lane;
      if (i < cnt * FI) s_in[warp][i] = src[i];
    }
    __syncwarp();
    if (lane < cnt) {
      const float* o = &s_in[warp][lane * FI];
      const float dx = o[12] - o[0], dy = o[13] - o[1], dz = o[14] - o[2];
      const float distance = sqrtf(dx * dx + dy * dy);
      const float rel = atan2f(dy, dx) - o[11];
      float ca, sa, cb, sb, cp, sp, ct, st, cr, sr;
      sincosf(o[4], &sa, &ca);
      sincosf(o[5], &sb, &cb);
      sincosf(o[9], &sp, &cp);
      sincosf(o[10], &st, &ct);
      sincosf(rel, &sr, &cr);
      float* y = &s_out[warp][lane * FO];
      y[0] = 1.0f / (1.0f + distance * 1e-3f);
      y[1] = dz / 15000.0f;
      y[2] = o[2] / 15000.0f;
      y[3] = o[3];
      y[4] = o[6]; y[5] = o[7]; y[6] = o[8];
      y[7] = ca; y[8] = cb; y[9] = sa; y[10] = sb;
      y[11] = cp; y[12] = ct; y[13] = sp; y[14] = st;
      y[15] = cr; y[16] = sr;
    }
    __syncwarp();
    float* dst = out + f0 * FO;
#pragma unroll
    for (int k = 0; k < FO; ++k) {
      int i = k * 32 + lane;
      if (i < cnt * FO) dst[i] = s_out[warp][i];
    }
    __syncwarp();
  }
}
}  // namespace

extern "C" int f16_features17(int64_t n_frames, const float* frames, float* out, void* stream) {
  if (n_frames <= 0) return f16_internal_fail("f16_features17: n_frames must be positive");
  if (!frames || !out) return f16_internal_fail("f16_features17: NULL pointer");
  int64_t tiles = (n_frames + 31) / 32;
  int64_t ctas = (tiles + WARPS - 1) / WARPS;
  unsigned grid = (unsigned)(ctas < 148 * 8 ? ctas : 148 * 8);
  features17_kernel<<<grid, WARPS * 32, 0, (cudaStream_t)stream>>>(n_frames, frames, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  f16_internal_count_launch();
  return 0;
}
