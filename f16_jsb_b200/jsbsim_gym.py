"""Single-env Gymnasium surface of the reference (jsbsim_gym/jsbsim_gym.py:95-545) on the CUDA path.

Same names, argument meaning and return values: `JSBSimEnv` (reset/step/close/render, spaces,
attributes), `PositionReward`, `wrap_jsbsim`, and - when gymnasium is importable - registration of
"JSBSim-v0" with max_episode_steps=1200. One env on a GPU is launch-latency bound; this module exists
so code written against the reference's class keeps working. Throughput comes from F16VecEnv.
"""
import numpy as np
import torch

from . import _compat
from .batched_env import F16BatchedEnv
from .constants import NUM_STACKED_FRAMES, REWARD_GAIN, sample_goal_numpy
from .vec_env import make_spaces

_Base = _compat.gym.Env if _compat.HAVE_GYMNASIUM else object
_WrapperBase = _compat.gym.Wrapper if _compat.HAVE_GYMNASIUM else object


class JSBSimEnv(_Base):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}

    def __init__(self, root: str = ".", device=None, mode: str = "fp64", reset_mode: str = "carryover"):
        """reset_mode "carryover" (default here): a second reset() of this object does what the reference object
        does - run_ic() + set-running on top of what the last episode left in the control laws, air data and
        accelerations (jsbsim_gym.py:305-306). "snapshot": every episode starts like the first one of a new env."""
        if _compat.HAVE_GYMNASIUM:
            super().__init__()
        self.num_stacked_frames = NUM_STACKED_FRAMES
        self.observation_space, self.action_space = make_spaces()
        self._env = F16BatchedEnv(1, device=device, mode=mode, with_terminal_obs=False, reset_mode=reset_mode, obs_layout="stacked")
        self.down_sample = 4
        self.current_step = 0
        self.max_episode_steps = 1200
        self.goal = np.zeros(3, dtype=np.float32)
        self.dg = 100.0
        self.viewer = None
        # pinned host staging: one call = upload, step, download, ONE stream synchronisation (f16_step_host)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        self._act, self._obs = pin((1, 4), torch.float32), pin((1, NUM_STACKED_FRAMES, 15), torch.float32)
        self._rew, self._done, self._trunc = pin((1,), torch.float32), pin((1,), torch.uint8), pin((1,), torch.uint8)

    def step(self, action):
        """jsbsim_gym.py:199-287: returns (obs (10,15) f32, base reward, terminated, truncated, {})."""
        self.current_step += 1
        self._act[...] = np.asarray(action, dtype=np.float32).reshape(1, 4)
        self._env.step_host(self._act, self._obs, self._rew, self._done, self._trunc, auto_reset=False)
        obs = self._obs[0].copy()
        truncated = bool(self._trunc[0])
        terminated = bool(self._done[0]) and not truncated
        # The kernel's reward has PositionReward's shaping fused in; this class returns the BASE reward, which the
        # reference computes from the very float32 frame just returned (jsbsim_gym.py:237-256): crash below 10 m -> -10
        # (it takes precedence), otherwise a terminated episode reached the goal cylinder -> +10, else 0.
        reward = 0.0
        if terminated:
            reward = -10.0 if obs[-1][2] < 10.0 else 10.0
        return obs, reward, terminated, truncated, {}

    def reset(self, seed: int = None, options: dict = None):
        """jsbsim_gym.py:289-331: goal from np.random.default_rng(seed), stack filled with the reset frame."""
        self.current_step = 0
        self.goal = sample_goal_numpy(seed)
        obs = self._env.reset(goals=torch.from_numpy(self.goal.reshape(1, 3)).to(self._env.device))
        return obs[0].cpu().numpy(), {}

    def render(self, mode: str = "human"):
        return None   # the reference's viewer is not on the step path (and is broken upstream)

    def close(self):
        self._env.close()


class PositionReward(_WrapperBase):
    """jsbsim_gym.py:470-519: reward += gain * (last_distance - distance), float32 3-D distance."""

    def __init__(self, env: JSBSimEnv, gain: float):
        if _compat.HAVE_GYMNASIUM:
            super().__init__(env)
        else:
            self.env = env
            self.observation_space, self.action_space = env.observation_space, env.action_space
        self.gain = gain
        self.last_distance = 0.0

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        current_distance = np.linalg.norm(obs[-1][-3:] - obs[-1][:3])
        reward += self.gain * (self.last_distance - current_distance)
        self.last_distance = current_distance
        return obs, reward, terminated, truncated, info

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        self.last_distance = np.linalg.norm(obs[-1][-3:] - obs[-1][:3])
        return obs, info

    def close(self):
        return self.env.close()


def wrap_jsbsim(**kwargs) -> PositionReward:
    """jsbsim_gym.py:521-533."""
    return PositionReward(JSBSimEnv(**kwargs), gain=REWARD_GAIN)


if _compat.HAVE_GYMNASIUM:  # pragma: no cover - gymnasium is absent in the build container
    try:
        _compat.gym.register(id="JSBSim-v0", entry_point="f16_jsb_b200.jsbsim_gym:wrap_jsbsim", max_episode_steps=1200)
    except Exception:
        pass
