"""gymnasium / stable_baselines3 interop. Uses the real packages when they import, else minimal
stand-ins with the same attribute surface (neither is installable in the build container)."""
from abc import ABC, abstractmethod

import numpy as np

try:  # pragma: no cover - depends on the user's environment
    import gymnasium as gym
    from gymnasium import spaces
    HAVE_GYMNASIUM = True
except Exception:  # pragma: no cover
    gym = None
    HAVE_GYMNASIUM = False

    class _Spaces:
        class Space:
            pass

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
                self.dtype = np.dtype(dtype)
                self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
                self._rng = np.random.default_rng(seed)

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1.0)
                hi = np.where(np.isfinite(self.high), self.high, 1.0)
                return self._rng.uniform(lo, hi).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return "Box(%s, %s)" % (self.shape, self.dtype)

    spaces = _Spaces()

try:  # pragma: no cover
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _SB3VecEnv
    HAVE_SB3 = True
except Exception:  # pragma: no cover
    _SB3VecEnv = None
    HAVE_SB3 = False


class _FallbackVecEnv(ABC):
    """Attribute-for-attribute stand-in for stable_baselines3 VecEnv (base_vec_env.py:50-357)."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]
        self.render_mode = None

    def _reset_seeds(self):
        self._seeds = [None for _ in range(self.num_envs)]

    def _reset_options(self):
        self._options = [{} for _ in range(self.num_envs)]

    @abstractmethod
    def reset(self): ...

    @abstractmethod
    def step_async(self, actions): ...

    @abstractmethod
    def step_wait(self): ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def seed(self, seed=None):
        if seed is None:
            seed = int(np.random.randint(0, np.iinfo(np.uint32).max, dtype=np.uint32))
        self._seeds = [seed + idx for idx in range(self.num_envs)]
        return self._seeds

    def set_options(self, options=None):
        if options is None:
            options = {}
        self._options = [dict(options) for _ in range(self.num_envs)] if isinstance(options, dict) else list(options)

    @property
    def unwrapped(self):
        return self

    def getattr_depth_check(self, name, already_found):
        return None

    def _get_indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    def render(self, mode=None):
        return None

    def get_images(self):
        return [None for _ in range(self.num_envs)]


VecEnvBase = _SB3VecEnv if HAVE_SB3 else _FallbackVecEnv
