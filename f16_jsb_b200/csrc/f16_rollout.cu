// f16_rollout.cu - GPU-resident rollout storage, GAE and minibatch gather (include/f16_rollout.h).
// Pure HBM-bound byte movers plus one sequential-in-time scan per env; everything is coalesced over
// the env index.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/f16_rollout.h"

extern "C" int f16_internal_fail(const char* msg);   // sets f16_last_error (f16_b200.cu)
extern "C" void f16_internal_count_launch(void);

namespace {
constexpr int FR = 15, STACK = 10;

__global__ void rollout_add_kernel(int64_t n, int64_t t, int64_t T, const float* __restrict__ obs, const float* __restrict__ actions,
                                   const float* __restrict__ rewards, const uint8_t* __restrict__ episode_starts,
                                   const float* __restrict__ values, const float* __restrict__ log_probs, const uint8_t* __restrict__ age0,
                                   float* __restrict__ frames, uint8_t* __restrict__ age, float* __restrict__ actions_buf,
                                   float* __restrict__ rewards_buf, float* __restrict__ es_buf, float* __restrict__ values_buf,
                                   float* __restrict__ log_probs_buf) {
  // frames: one thread per float; rows_to_copy rows of 15 floats per env
  const int rows = (t == 0) ? STACK : 1;
  const int first_row = STACK - rows;
  const int64_t total = n * rows * FR;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / (rows * FR);
    const int rem = (int)(i - e * (rows * FR));
    const int r = first_row + rem / FR, c = rem % FR;
    // slot of obs row r at step t: t + r (row 9 -> t + 9)
    frames[((t + r) * n + e) * FR + c] = obs[(e * STACK + r) * FR + c];
  }
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t es = episode_starts[e];
    const uint8_t prev = (t == 0) ? (age0 ? age0[e] : (uint8_t)9) : age[(t - 1) * n + e];
    age[t * n + e] = es ? (uint8_t)0 : (uint8_t)(prev >= 9 ? 9 : prev + 1);
    reinterpret_cast<float4*>(actions_buf)[t * n + e] = reinterpret_cast<const float4*>(actions)[e];
    rewards_buf[t * n + e] = rewards[e];
    es_buf[t * n + e] = es ? 1.0f : 0.0f;
    values_buf[t * n + e] = values[e];
    log_probs_buf[t * n + e] = log_probs[e];
  }
}

// buffers.py:426-438 in float32, operation by operation (NumPy >= 2 keeps float32 with Python-float scalars)
__global__ void rollout_gae_kernel(int64_t n, int64_t T, float gamma, float gamma_lambda, const float* __restrict__ rewards,
                                   const float* __restrict__ values, const float* __restrict__ episode_starts,
                                   const float* __restrict__ last_values, const uint8_t* __restrict__ dones,
                                   float* __restrict__ advantages, float* __restrict__ returns) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float last_gae_lam = 0.0f;
  float next_values = last_values[e];
  float next_non_terminal = __fsub_rn(1.0f, dones[e] ? 1.0f : 0.0f);
  // the recurrence is sequential in time, the loads are not: fetch CH steps ahead of the dependent chain
  constexpr int CH = 8;
  for (int64_t hi = T; hi > 0; hi -= CH) {
    float r[CH], v[CH], es[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int64_t step = hi - 1 - k;
      if (step >= 0) { r[k] = rewards[step * n + e]; v[k] = values[step * n + e]; es[k] = episode_starts[step * n + e]; }
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int64_t step = hi - 1 - k;
      if (step < 0) break;
      const float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(gamma, next_values), next_non_terminal)), v[k]);
      last_gae_lam = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, next_non_terminal), last_gae_lam));
      advantages[step * n + e] = last_gae_lam;
      returns[step * n + e] = __fadd_rn(last_gae_lam, v[k]);
      next_values = v[k];
      next_non_terminal = __fsub_rn(1.0f, es[k]);
    }
  }
}

// one warp per sample: 150 floats of the rebuilt stack, lanes stride over them
__global__ void rollout_gather_kernel(int64_t n, int64_t T, int64_t batch, const int64_t* __restrict__ indices,
                                      const float* __restrict__ frames, const uint8_t* __restrict__ age,
                                      const float* __restrict__ actions_buf, const float* __restrict__ values_buf,
                                      const float* __restrict__ log_probs_buf, const float* __restrict__ advantages,
                                      const float* __restrict__ returns, float* __restrict__ obs_out, float* __restrict__ actions_out,
                                      float* __restrict__ values_out, float* __restrict__ log_probs_out,
                                      float* __restrict__ advantages_out, float* __restrict__ returns_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < batch; b += warps) {
    const int64_t idx = indices[b];
    const int64_t e = idx / T, t = idx - e * T;       // swap_and_flatten order: index = env * T + step
    const int k = age[t * n + e];                      // valid history frames (0 = reset observation)
    for (int j = lane; j < STACK * FR; j += 32) {
      const int r = j / FR, c = j - r * FR;
      // row r of obs_t is the frame of step t-9+r, but never older than the episode start (t-k)
      int back = 9 - r;
      if (back > k) back = k;
      obs_out[b * (STACK * FR) + j] = frames[((t + 9 - back) * n + e) * FR + c];
    }
    if (lane < 4) actions_out[b * 4 + lane] = actions_buf[(t * n + e) * 4 + lane];
    if (lane == 4) values_out[b] = values_buf[t * n + e];
    if (lane == 5) log_probs_out[b] = log_probs_buf[t * n + e];
    if (lane == 6) advantages_out[b] = advantages[t * n + e];
    if (lane == 7) returns_out[b] = returns[t * n + e];
  }
}

// a warp per env: the terminal observation of an env cut by the time limit goes to the next free slot of the side buffer
__global__ void __launch_bounds__(256) park_truncated_kernel(int64_t n, int64_t step, int64_t capacity, int obs_floats,
                                                             const uint8_t* __restrict__ truncated, const float* __restrict__ terminal_obs,
                                                             float* __restrict__ parked_obs, int64_t* __restrict__ parked_flat,
                                                             unsigned long long* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += warps) {
    if (!truncated[e]) continue;
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(count, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if ((int64_t)slot >= capacity) continue;               // the count keeps running: the caller sees the overflow
    for (int i = lane; i < obs_floats; i += 32) parked_obs[slot * obs_floats + i] = terminal_obs[e * obs_floats + i];
    if (lane == 0) parked_flat[slot] = step * n + e;
  }
}

int check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return f16_internal_fail(cudaGetErrorString(e));
  (void)what;
  f16_internal_count_launch();
  return 0;
}
}  // namespace

extern "C" {

int f16_rollout_add(int64_t n, int64_t t, int64_t T, const float* obs, const float* actions, const float* rewards,
                    const uint8_t* episode_starts, const float* values, const float* log_probs, const uint8_t* age0, float* frames,
                    uint8_t* age, float* actions_buf, float* rewards_buf, float* es_buf, float* values_buf, float* log_probs_buf,
                    void* stream) {
  if (n <= 0 || T <= 0 || t < 0 || t >= T) return f16_internal_fail("f16_rollout_add: bad sizes");
  if (!obs || !actions || !rewards || !episode_starts || !values || !log_probs || !frames || !age) return f16_internal_fail("f16_rollout_add: NULL pointer");
  const int64_t work = n * ((t == 0) ? STACK : 1) * FR;
  unsigned grid = (unsigned)((work + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  rollout_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, t, T, obs, actions, rewards, episode_starts, values, log_probs, age0, frames,
                                                              age, actions_buf, rewards_buf, es_buf, values_buf, log_probs_buf);
  return check("rollout_add");
}

int f16_rollout_gae(int64_t n, int64_t T, float gamma, float gae_lambda, const float* rewards, const float* values,
                    const float* episode_starts, const float* last_values, const uint8_t* dones, float* advantages, float* returns,
                    void* stream) {
  if (n <= 0 || T <= 0) return f16_internal_fail("f16_rollout_gae: bad sizes");
  // self.gamma * self.gae_lambda is a Python float product, rounded to float32 when it meets the arrays
  const float gl = (float)((double)gamma * (double)gae_lambda);
  rollout_gae_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, T, gamma, gl, rewards, values, episode_starts,
                                                                                     last_values, dones, advantages, returns);
  return check("rollout_gae");
}

int f16_rollout_gather(int64_t n, int64_t T, int64_t batch, const int64_t* indices, const float* frames, const uint8_t* age,
                       const float* actions_buf, const float* values_buf, const float* log_probs_buf, const float* advantages,
                       const float* returns, float* obs_out, float* actions_out, float* values_out, float* log_probs_out,
                       float* advantages_out, float* returns_out, void* stream) {
  if (n <= 0 || T <= 0 || batch <= 0) return f16_internal_fail("f16_rollout_gather: bad sizes");
  unsigned grid = (unsigned)((batch * 32 + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  rollout_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, T, batch, indices, frames, age, actions_buf, values_buf, log_probs_buf,
                                                                 advantages, returns, obs_out, actions_out, values_out, log_probs_out,
                                                                 advantages_out, returns_out);
  return check("rollout_gather");
}

int f16_rollout_park_truncated(int64_t n, int64_t step, int64_t capacity, const uint8_t* truncated, const float* terminal_obs,
                               float* parked_obs, int64_t* parked_flat, int64_t* count, void* stream) {
  if (n <= 0 || step < 0 || capacity <= 0) return f16_internal_fail("f16_rollout_park_truncated: bad sizes");
  if (!truncated || !terminal_obs || !parked_obs || !parked_flat || !count) return f16_internal_fail("f16_rollout_park_truncated: NULL pointer");
  unsigned grid = (unsigned)((n + 7) / 8);
  if (grid > 148u * 8u) grid = 148u * 8u;
  park_truncated_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, step, capacity, STACK * FR, truncated, terminal_obs, parked_obs, parked_flat,
                                                                 (unsigned long long*)count);
  return check("rollout_park_truncated");
}

}  // extern "C"
