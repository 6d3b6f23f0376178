"""Minimal gymnasium look-alike, enough for jsbsim_gym/jsbsim_gym.py. TEST INFRASTRUCTURE ONLY."""
import importlib

import numpy as np

from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401

__version__ = "0.29.1"      # what stable_baselines3/common/utils.py:get_system_info prints

_REGISTRY = {}


class Env:
    def __class_getitem__(cls, item):       # gym.Env[ObsType, ActType] in type annotations / base-class lists
        return cls

    metadata = {}
    render_mode = None
    spec = None
    np_random = None

    def reset(self, *, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    @property
    def observation_space(self):
        return self.__dict__.get("_observation_space") or self.env.observation_space

    @observation_space.setter
    def observation_space(self, v):
        self.__dict__["_observation_space"] = v

    @property
    def action_space(self):
        return self.__dict__.get("_action_space") or self.env.action_space

    @action_space.setter
    def action_space(self, v):
        self.__dict__["_action_space"] = v

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def close(self):
        return self.env.close()


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


class _Spec:
    def __init__(self, id, entry_point, max_episode_steps, kwargs):
        self.id, self.entry_point, self.max_episode_steps, self.kwargs = id, entry_point, max_episode_steps, kwargs


def register(id, entry_point, max_episode_steps=None, kwargs=None, **_):
    _REGISTRY[id] = _Spec(id, entry_point, max_episode_steps, kwargs or {})


def make(id, **kwargs):
    spec = _REGISTRY[id]
    ep = spec.entry_point
    if isinstance(ep, str):
        mod, fn = ep.split(":")
        ep = getattr(importlib.import_module(mod), fn)
    env = ep(**{**spec.kwargs, **kwargs})
    if spec.max_episode_steps:
        env = TimeLimit(env, spec.max_episode_steps)
    env.spec = spec
    return env


class ObservationWrapper(Wrapper):
    def observation(self, observation):
        return observation


class RewardWrapper(Wrapper):
    def reward(self, reward):
        return reward


class ActionWrapper(Wrapper):
    def action(self, action):
        return action

