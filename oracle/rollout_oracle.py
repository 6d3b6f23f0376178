"""NumPy restatement of the reference's RolloutBuffer (TEST INFRASTRUCTURE ONLY).

Follows stable_baselines3/common/buffers.py of the reference line by line for the parts the GPU
rollout store replaces: reset :392-402, add :440-478 (full observations are stored, as SB3 does),
compute_returns_and_advantage :404-438, swap_and_flatten :63-75 and _get_samples :508-521. PINNED: the reference's real
class (imported from /root/reference over oracle/refshim by tools/make_golden_rollout.py) produced
tests/golden/sb3_rollout_buffer.npz, which this restatement reproduces bit for bit (tests/test_rollout_oracle.py). It is
the checker for tests/test_gpu_rollout.py, where the reference tree is not available.
"""
import numpy as np


class RolloutBufferOracle:
    def __init__(self, buffer_size, n_envs, obs_shape=(10, 15), action_dim=4, gae_lambda=1.0, gamma=0.99):
        self.buffer_size, self.n_envs, self.obs_shape, self.action_dim = buffer_size, n_envs, obs_shape, action_dim
        self.gae_lambda, self.gamma = gae_lambda, gamma
        self.reset()

    def reset(self):
        T, N = self.buffer_size, self.n_envs
        self.observations = np.zeros((T, N, *self.obs_shape), dtype=np.float32)
        self.actions = np.zeros((T, N, self.action_dim), dtype=np.float32)
        self.rewards = np.zeros((T, N), dtype=np.float32)
        self.returns = np.zeros((T, N), dtype=np.float32)
        self.episode_starts = np.zeros((T, N), dtype=np.float32)
        self.values = np.zeros((T, N), dtype=np.float32)
        self.log_probs = np.zeros((T, N), dtype=np.float32)
        self.advantages = np.zeros((T, N), dtype=np.float32)
        self.pos, self.full = 0, False

    def add(self, obs, action, reward, episode_start, value, log_prob):
        self.observations[self.pos] = np.array(obs)
        self.actions[self.pos] = np.array(action).reshape((self.n_envs, self.action_dim))
        self.rewards[self.pos] = np.array(reward)
        self.episode_starts[self.pos] = np.array(episode_start)
        self.values[self.pos] = np.array(value).flatten()
        self.log_probs[self.pos] = np.array(log_prob)
        self.pos += 1
        if self.pos == self.buffer_size:
            self.full = True

    def compute_returns_and_advantage(self, last_values, dones):
        last_values = np.array(last_values, dtype=np.float32).flatten()
        last_gae_lam = 0
        for step in reversed(range(self.buffer_size)):
            if step == self.buffer_size - 1:
                next_non_terminal = 1.0 - dones.astype(np.float32)
                next_values = last_values
            else:
                next_non_terminal = 1.0 - self.episode_starts[step + 1]
                next_values = self.values[step + 1]
            delta = self.rewards[step] + self.gamma * next_values * next_non_terminal - self.values[step]
            last_gae_lam = delta + self.gamma * self.gae_lambda * next_non_terminal * last_gae_lam
            self.advantages[step] = last_gae_lam
        self.returns = self.advantages + self.values

    @staticmethod
    def swap_and_flatten(arr):
        shape = arr.shape
        if len(shape) < 3:
            shape = (*shape, 1)
        return arr.swapaxes(0, 1).reshape(shape[0] * shape[1], *shape[2:])

    def get_samples(self, batch_inds):
        f = self.swap_and_flatten
        return (f(self.observations)[batch_inds], f(self.actions)[batch_inds], f(self.values)[batch_inds].flatten(),
                f(self.log_probs)[batch_inds].flatten(), f(self.advantages)[batch_inds].flatten(), f(self.returns)[batch_inds].flatten())
