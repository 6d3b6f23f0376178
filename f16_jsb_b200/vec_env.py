"""F16VecEnv - Stable-Baselines3 VecEnv over the batched CUDA env.

Drop-in for what `PPO('MlpPolicy', gym.make("JSBSim-v0"))` builds internally in the reference
(Monitor + DummyVecEnv around JSBSimEnv/PositionReward/TimeLimit; train.py:34,113,
stable_baselines3/common/base_class.py:204-224): pass an instance to PPO/SAC and SB3 uses it as is
(`isinstance(env, VecEnv)` short-circuits the wrapping). Conventions reproduced from
stable_baselines3/common/vec_env/dummy_vec_env.py:56-73 and common/monitor.py:85-111:
done = terminated or truncated; info["TimeLimit.truncated"] = truncated and not terminated;
on done info["terminal_observation"] holds the last stacked observation, the env is reset in the
same step and the returned observation is the reset one; info["episode"] = {"r", "l", "t"}.
"""
import time
from collections.abc import Sequence
from typing import Any, Optional

import numpy as np
import torch

from ._compat import VecEnvBase, spaces
from .batched_env import F16BatchedEnv
from .host_window import HostWindow
from .constants import (ACTION_HIGH, ACTION_LOW, NUM_FEATURES, NUM_STACKED_FRAMES, SINGLE_OBS_HIGH, SINGLE_OBS_LOW,
                        sample_goal_numpy)

_EMPTY_INFO_KEYS = ("TimeLimit.truncated",)


class _LazyInfos(Sequence):
    """list[dict]-like view: envs that did not finish share one read-only dict, finished envs get
    their own dict. Keeps SB3's per-step O(N) Python loops cheap for thousands of envs."""

    def __init__(self, n: int, done_infos):
        self._n = n
        self._done = done_infos
        self._shared = {"TimeLimit.truncated": False}

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._done.get(i, self._shared)


class _RecordInfos:
    """dict-like {env index: info dict} over the done records of one window-mode step; the dicts are
    built on first access (at a million envs a step finishes thousands of episodes)."""

    def __init__(self, records: np.ndarray, terminal_obs: np.ndarray, t: float):
        self._rec, self._term, self._t = records, terminal_obs, t
        self._where = None
        self._built: dict = {}

    def get(self, i, default=None):
        if self._rec.size == 0:
            return default
        if self._where is None:
            self._where = {int(e): j for j, e in enumerate(self._rec["env"].tolist())}
        j = self._where.get(int(i))
        if j is None:
            return default
        d = self._built.get(j)
        if d is None:
            r = self._rec[j]
            # an owned copy: the row lives in a library buffer that is recycled two steps later, and SB3 users keep
            # info dicts (DummyVecEnv hands out owned arrays too, dummy_vec_env.py:66-67)
            d = {"TimeLimit.truncated": bool(r["flags"] & 1), "terminal_observation": self._term[j].copy(),
                 "episode": {"r": float(r["ep_return"]), "l": int(r["ep_len"]), "t": self._t}}
            self._built[j] = d
        return d


def make_spaces():
    obs_space = spaces.Box(low=np.tile(SINGLE_OBS_LOW, (NUM_STACKED_FRAMES, 1)),
                           high=np.tile(SINGLE_OBS_HIGH, (NUM_STACKED_FRAMES, 1)),
                           shape=(NUM_STACKED_FRAMES, NUM_FEATURES), dtype=np.float32)
    act_space = spaces.Box(low=ACTION_LOW.copy(), high=ACTION_HIGH.copy(), shape=(4,), dtype=np.float32)
    return obs_space, act_space


class F16VecEnv(VecEnvBase):
    """num_envs F-16 goal-reaching envs on one GPU behind the SB3 VecEnv interface.

    host_obs: "window" (default) keeps the ten-frame windows in pinned host memory and moves only the
    newest frame of every env across PCIe each step (60 B instead of 600 B per env-step); the returned
    observation is a strided (N,10,15) float32 view that stays intact until the step after next
    (host_rings=2; host_rings=1 halves the PCIe traffic again and keeps it intact until the next step).
    "copy" keeps the stacks on the device and copies all of them out every step into `host_ring`
    rotating pinned buffers. copy_obs=True returns fresh arrays (DummyVecEnv's behaviour) in either mode.
    `action_buffer()` hands out pinned staging arrays: actions written there go to the device by plain DMA.
    reset_mode: "snapshot" (default) restarts a finished env from the state of a fresh reference env object;
    "carryover" applies the reference's run_ic() + set-running to the env as the episode left it, which is what
    DummyVecEnv's env objects go through from their second episode on (F16BatchedEnv, include/f16_b200.h).
    """

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}   # jsbsim_gym.py:120

    def __init__(self, num_envs: int, device=None, mode: str = "fp32", seed: int = 0, host_ring: int = 2,
                 copy_obs: bool = False, lazy_infos: Optional[bool] = None, env_id_base: int = 0,
                 obs_layout: str = "stacked", host_obs: str = "window", host_rings: int = 2,
                 host_dma_both: bool = False, reset_mode: str = "snapshot"):
        obs_space, act_space = make_spaces()
        if host_obs not in ("window", "copy"):
            raise ValueError("host_obs must be 'window' or 'copy'")
        self.host_obs = host_obs
        self.env = F16BatchedEnv(num_envs, device=device, mode=mode, seed=seed, env_id_base=env_id_base,
                                 obs_layout="frame" if host_obs == "window" else obs_layout, reset_mode=reset_mode)
        # what a finished env restarts from: F16_AUTO_RESET_SNAPSHOT (fresh-env state) or F16_AUTO_RESET_CARRYOVER
        # (the reference's run_ic() on the used env object; include/f16_b200.h)
        self._auto_reset = 2 if reset_mode == "carryover" else 1
        self.render_mode = None
        try:
            super().__init__(num_envs, obs_space, act_space)
        except Exception:  # SB3's ctor calls get_attr("render_mode"); keep going with plain attributes
            self.num_envs, self.observation_space, self.action_space = num_envs, obs_space, act_space
            self.reset_infos = [{} for _ in range(num_envs)]
            self._seeds = [None for _ in range(num_envs)]
            self._options = [{} for _ in range(num_envs)]
        self.copy_obs = copy_obs
        self.lazy_infos = (num_envs > 64) if lazy_infos is None else lazy_infos
        ring = max(2, int(host_ring))
        n = num_envs
        self._win = None
        if host_obs == "window":
            with torch.cuda.device(self.env.device):
                self._win = HostWindow(n, n_rings=int(host_rings), pin=True, dma_both=host_dma_both)
            self._act_bufs = self._win.action_buffers
        else:
            self._h_obs = [torch.empty((n, NUM_STACKED_FRAMES, NUM_FEATURES), dtype=torch.float32, pin_memory=True) for _ in range(ring)]
            self._h_rew = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(ring)]
            self._h_done = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(ring)]
            self._h_trunc = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(ring)]
            self._h_act = [torch.empty((n, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)]
            self._act_bufs = [t.numpy() for t in self._h_act]
        self._act_next = 0
        self._slot = 0
        self._actions = None
        self._t_start = time.time()
        # attributes of the single env that get_attr serves (jsbsim_gym.py:131,157-163)
        self.num_stacked_frames = NUM_STACKED_FRAMES
        self.down_sample = 4
        self.max_episode_steps = 1200
        self.dg = 100.0

    # ------------------------------------------------------------------ VecEnv API
    def reset(self) -> np.ndarray:
        goals = None
        if any(s is not None for s in self._seeds):
            # reference semantics: goal drawn from np.random.default_rng(seed) per env (jsbsim_gym.py:312-323)
            g = np.stack([sample_goal_numpy(s) for s in self._seeds])
            goals = torch.from_numpy(g).to(self.env.device)
        if self._auto_reset == 2 and self._actions is not None:
            self.env._last_actions = torch.from_numpy(np.ascontiguousarray(self._actions)).to(self.env.device)
        obs = self.env.reset(goals=goals)
        if self._win is not None:
            out = self._win.reset(self.env, self.env._stream()).obs
        else:
            host = self._h_obs[self._slot]
            host.copy_(obs, non_blocking=True)
            torch.cuda.current_stream(self.env.device).synchronize()
            out = host.numpy()
        self._reset_seeds()
        self._reset_options()
        self.reset_infos = [{} for _ in range(self.num_envs)]
        return out.copy() if self.copy_obs else out

    def action_buffer(self) -> np.ndarray:
        """A pinned (N,4) float32 staging array (two rotate): fill it and pass it to step()/step_async()
        and the host->device copy is a plain DMA with no intermediate copy."""
        self._act_next ^= 1
        return self._act_bufs[self._act_next]

    def step_async(self, actions: np.ndarray) -> None:
        for b in self._act_bufs:                      # the caller filled one of our pinned staging buffers
            if actions is b:
                self._actions = b
                return
        a = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, 4)
        for b in self._act_bufs:
            if a.ctypes.data == b.ctypes.data:
                self._actions = b
                return
        b = self.action_buffer()
        b[...] = a
        self._actions = b

    def _step_wait_window(self):
        res = self._win.step(self.env, self._actions, self.env._stream(), auto_reset=self._auto_reset)
        dones = res.done.view(np.bool_)
        done_infos = _RecordInfos(res.records, res.terminal_obs, round(time.time() - self._t_start, 6))
        if self.lazy_infos:
            infos: Any = _LazyInfos(self.num_envs, done_infos)
        else:
            shared = {"TimeLimit.truncated": False}
            infos = [done_infos.get(i, shared) for i in range(self.num_envs)]
        if self.copy_obs:
            return self._win.gather(res.ring, res.first_slot), res.reward.copy(), dones.copy(), infos
        return res.obs, res.reward, dones, infos

    def step_wait(self):
        if self._win is not None:
            return self._step_wait_window()
        self._slot = (self._slot + 1) % len(self._h_obs)
        k = self._slot
        obs, rew, done, trunc = self._h_obs[k].numpy(), self._h_rew[k].numpy(), self._h_done[k].numpy(), self._h_trunc[k].numpy()
        self.env.step_host(self._actions, obs, rew, done, trunc, auto_reset=self._auto_reset)
        dones = done.astype(bool)
        done_infos: dict = {}
        idx = np.flatnonzero(dones)
        if idx.size:
            sel = torch.from_numpy(idx).to(self.env.device)
            term_obs = self.env.terminal_obs.index_select(0, sel).cpu().numpy()
            ep_r = self.env.ep_return.index_select(0, sel).cpu().numpy()
            ep_l = self.env.ep_len.index_select(0, sel).cpu().numpy()
            t = round(time.time() - self._t_start, 6)
            for j, i in enumerate(idx.tolist()):
                done_infos[i] = {"TimeLimit.truncated": bool(trunc[i]), "terminal_observation": term_obs[j],
                                 "episode": {"r": float(ep_r[j]), "l": int(ep_l[j]), "t": t}}
        if self.lazy_infos:
            infos: Any = _LazyInfos(self.num_envs, done_infos)
        else:
            infos = [done_infos.get(i, {"TimeLimit.truncated": False}) for i in range(self.num_envs)]
        if self.copy_obs:
            return obs.copy(), rew.copy(), dones, infos
        return obs, rew.copy(), dones, infos

    def close(self) -> None:
        if self._win is not None:
            self._win.close(self.env)
            self._win = None
        self.env.close()

    def has_attr(self, attr_name: str) -> bool:
        return hasattr(self, attr_name)

    def get_attr(self, attr_name: str, indices=None):
        n = len(list(self._get_indices(indices)))
        return [getattr(self, attr_name, None) for _ in range(n)]

    def set_attr(self, attr_name: str, value, indices=None) -> None:
        setattr(self, attr_name, value)

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs):
        """base_vec_env.py:190-201. There are no per-env Python objects; a method of this VecEnv that makes sense per
        env (e.g. `render`, `close`-free queries) is called once and its result replicated, anything else raises the
        AttributeError DummyVecEnv's getattr would raise."""
        n = len(list(self._get_indices(indices)))
        fn = getattr(self, method_name, None)
        if fn is None or not callable(fn) or method_name in ("step", "step_async", "step_wait", "reset", "close", "seed"):
            raise AttributeError("the F-16 env has no per-env method %r" % method_name)
        out = fn(*method_args, **method_kwargs)
        return [out for _ in range(n)]

    def render(self, mode: str = "human"):
        return None   # the reference's viewer is not on the step path (jsbsim_gym.py:333-468)

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = len(list(self._get_indices(indices)))
        return [False for _ in range(n)]

    # ------------------------------------------------------------------ device-tensor fast path
    def _need_device_stacks(self):
        if self._win is not None:
            raise RuntimeError("the device-tensor path needs the stacked observations on the device: construct "
                               "F16VecEnv(..., host_obs='copy') or use F16BatchedEnv directly")

    def reset_torch(self, goals: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._need_device_stacks()
        return self.env.reset(goals=goals)

    def step_torch(self, actions: Optional[torch.Tensor]):
        """Zero-copy path: cuda actions in, (obs, reward, done, truncated) device tensors out."""
        self._need_device_stacks()
        return self.env.step(actions, auto_reset=True)
