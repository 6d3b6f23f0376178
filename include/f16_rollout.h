/* f16_rollout.h - C ABI of the GPU-resident rollout store (libf16b200.so), SURVEY.md 8(f) row 1.
 *
 * Replaces, for device-resident training loops, the host-NumPy RolloutBuffer of the reference's vendored
 * Stable-Baselines3 (stable_baselines3/common/buffers.py:343-521) and the Python GAE loop
 * (buffers.py:426-438). Instead of the full (T, N, 10, 15) observation tensor (600 B per transition) it
 * stores the newest 15-float frame per transition (60 B) plus the number of valid history frames, and
 * rebuilds the (10,15) stacks when minibatches are gathered.
 *
 * All pointers are device pointers owned by the caller; launches go to the caller's stream.
 * Layouts: [T][N] row-major for per-transition scalars, [T+9][N][15] for frames (slots 0..8 hold rows
 * 0..8 of the first observation), flat sample index = n*T + t as SB3's swap_and_flatten produces.
 */
#ifndef F16_ROLLOUT_H
#define F16_ROLLOUT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* RolloutBuffer.add (buffers.py:440-478) for step t: frames[t+9] <- obs[:, 9, :]; when t == 0 also
 * frames[0..8] <- obs[:, 0..8, :]. age[t][n] = 0 if episode_start[n] else min(age_prev[n] + 1, 9)
 * (age_prev = age[t-1], or age0 for t == 0; age0 may be NULL = 9 unless an episode starts).
 * actions/rewards/values/log_probs/episode_starts are copied into their [T][N] slots. */
int f16_rollout_add(int64_t n_envs, int64_t t, int64_t T, const float* obs, const float* actions, const float* rewards,
                    const uint8_t* episode_starts, const float* values, const float* log_probs, const uint8_t* age0,
                    float* frames, uint8_t* age, float* actions_buf, float* rewards_buf, float* episode_starts_buf,
                    float* values_buf, float* log_probs_buf, void* stream);

/* RolloutBuffer.compute_returns_and_advantage (buffers.py:404-438): float32 GAE(lambda), bit-compatible
 * with the NumPy loop (same operation order, no FMA). last_values: [N], dones: [N] uint8. */
int f16_rollout_gae(int64_t n_envs, int64_t T, float gamma, float gae_lambda, const float* rewards, const float* values,
                    const float* episode_starts, const float* last_values, const uint8_t* dones, float* advantages,
                    float* returns, void* stream);

/* RolloutBuffer._get_samples (buffers.py:508-521) for `batch` flat indices (n*T + t): rebuilds the
 * stacked observations and gathers actions, values, log_probs, advantages, returns. */
int f16_rollout_gather(int64_t n_envs, int64_t T, int64_t batch, const int64_t* indices, const float* frames,
                       const uint8_t* age, const float* actions_buf, const float* values_buf, const float* log_probs_buf,
                       const float* advantages, const float* returns, float* obs_out, float* actions_out,
                       float* values_out, float* log_probs_out, float* advantages_out, float* returns_out, void* stream);

/* Time-limit bootstrap, first half (stable_baselines3/common/on_policy_algorithm.py:236-245: reward += gamma * V(terminal
 * observation) for episodes cut by the time limit). The policy is frozen during a rollout, so the terminal observations can be
 * valued once after the last step instead of inside the loop: for every env with truncated[env] != 0 this call copies
 * terminal_obs[env] ([N][10][15]) to parked_obs[slot] and writes parked_flat[slot] = step * n_envs + env (the flat [T][N] index
 * of the reward to adjust), slot = (*count)++ on the device. *count keeps running past `capacity` (entries beyond it are
 * dropped; the caller checks). One launch, no host synchronisation. */
int f16_rollout_park_truncated(int64_t n_envs, int64_t step, int64_t capacity, const uint8_t* truncated, const float* terminal_obs,
                               float* parked_obs, int64_t* parked_flat, int64_t* count, void* stream);

#ifdef __cplusplus
}
#endif
#endif
