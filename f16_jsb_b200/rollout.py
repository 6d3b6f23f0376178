"""GpuRolloutBuffer - device-resident replacement of SB3's RolloutBuffer for the F-16 env.

Same life cycle and sample layout as stable_baselines3/common/buffers.py:343-521 (`reset`, `add`,
`compute_returns_and_advantage`, `get` yielding observations / actions / old_values / old_log_prob /
advantages / returns with flat index = env * n_steps + step), but every array stays on the GPU, the GAE
scan is one CUDA kernel, and only the newest 15-float frame of each observation is stored: the (10,15)
stacks are rebuilt when a minibatch is gathered (60 B instead of 600 B per transition).
"""
import ctypes as C
from typing import Generator, NamedTuple, Optional

import torch

from . import _lib
from .constants import NUM_FEATURES, NUM_STACKED_FRAMES


class RolloutBufferSamples(NamedTuple):   # stable_baselines3/common/type_aliases.py:31-37
    observations: torch.Tensor
    actions: torch.Tensor
    old_values: torch.Tensor
    old_log_prob: torch.Tensor
    advantages: torch.Tensor
    returns: torch.Tensor


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class GpuRolloutBuffer:
    def __init__(self, buffer_size: int, n_envs: int, device="cuda", gae_lambda: float = 1.0, gamma: float = 0.99):
        self.lib = _lib.load()
        self.buffer_size, self.n_envs = int(buffer_size), int(n_envs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.F16Error("GpuRolloutBuffer only runs on CUDA devices")
        self.gae_lambda, self.gamma = float(gae_lambda), float(gamma)
        T, N = self.buffer_size, self.n_envs
        f32 = dict(dtype=torch.float32, device=self.device)
        self.frames = torch.zeros((T + NUM_STACKED_FRAMES - 1, N, NUM_FEATURES), **f32)
        self.age = torch.zeros((T, N), dtype=torch.uint8, device=self.device)
        self.actions = torch.zeros((T, N, 4), **f32)
        self.rewards = torch.zeros((T, N), **f32)
        self.episode_starts = torch.zeros((T, N), **f32)
        self.values = torch.zeros((T, N), **f32)
        self.log_probs = torch.zeros((T, N), **f32)
        self.advantages = torch.zeros((T, N), **f32)
        self.returns = torch.zeros((T, N), **f32)
        self._last_age = None          # history depth at the end of the previous rollout
        self.pos, self.full = 0, False

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self) -> None:
        if self.full:
            self._last_age = self.age[self.buffer_size - 1].clone()
        self.pos, self.full = 0, False

    def size(self) -> int:
        return self.buffer_size if self.full else self.pos

    def add(self, obs: torch.Tensor, action: torch.Tensor, reward: torch.Tensor, episode_start: torch.Tensor,
            value: torch.Tensor, log_prob: torch.Tensor) -> None:
        """obs (N,10,15) f32, action (N,4) f32, reward (N,) f32, episode_start (N,) bool/uint8, value (N,) or (N,1), log_prob (N,)."""
        assert self.pos < self.buffer_size, "buffer is full"
        N = self.n_envs
        obs = obs.to(self.device, torch.float32).contiguous()
        action = action.to(self.device, torch.float32).reshape(N, 4).contiguous()
        reward = reward.to(self.device, torch.float32).reshape(N).contiguous()
        es = episode_start.to(self.device).reshape(N).to(torch.uint8).contiguous()
        value = value.to(self.device, torch.float32).reshape(N).contiguous()
        log_prob = log_prob.to(self.device, torch.float32).reshape(N).contiguous()
        _lib.check(self.lib.f16_rollout_add(N, self.pos, self.buffer_size, _p(obs), _p(action), _p(reward), _p(es), _p(value),
                                            _p(log_prob), _p(self._last_age), _p(self.frames), _p(self.age), _p(self.actions),
                                            _p(self.rewards), _p(self.episode_starts), _p(self.values), _p(self.log_probs),
                                            self._stream()), "f16_rollout_add")
        self.pos += 1
        if self.pos == self.buffer_size:
            self.full = True

    def compute_returns_and_advantage(self, last_values: torch.Tensor, dones: torch.Tensor) -> None:
        N = self.n_envs
        lv = last_values.to(self.device, torch.float32).reshape(N).contiguous()
        d = dones.to(self.device).reshape(N).to(torch.uint8).contiguous()
        _lib.check(self.lib.f16_rollout_gae(N, self.buffer_size, self.gamma, self.gae_lambda, _p(self.rewards), _p(self.values),
                                            _p(self.episode_starts), _p(lv), _p(d), _p(self.advantages), _p(self.returns),
                                            self._stream()), "f16_rollout_gae")

    def gather(self, indices: torch.Tensor) -> RolloutBufferSamples:
        idx = indices.to(self.device, torch.int64).contiguous()
        B = idx.numel()
        f32 = dict(dtype=torch.float32, device=self.device)
        obs = torch.empty((B, NUM_STACKED_FRAMES, NUM_FEATURES), **f32)
        act = torch.empty((B, 4), **f32)
        val, lp, adv, ret = (torch.empty(B, **f32) for _ in range(4))
        _lib.check(self.lib.f16_rollout_gather(self.n_envs, self.buffer_size, B, _p(idx), _p(self.frames), _p(self.age), _p(self.actions),
                                               _p(self.values), _p(self.log_probs), _p(self.advantages), _p(self.returns), _p(obs), _p(act),
                                               _p(val), _p(lp), _p(adv), _p(ret), self._stream()), "f16_rollout_gather")
        return RolloutBufferSamples(obs, act, val, lp, adv, ret)

    def get(self, batch_size: Optional[int] = None, generator: Optional[torch.Generator] = None) -> Generator[RolloutBufferSamples, None, None]:
        assert self.full, "rollout buffer must be full before sampling"
        total = self.buffer_size * self.n_envs
        indices = torch.randperm(total, device=self.device, generator=generator)
        if batch_size is None:
            batch_size = total
        for start in range(0, total, batch_size):
            yield self.gather(indices[start:start + batch_size])
