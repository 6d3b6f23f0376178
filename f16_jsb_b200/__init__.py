"""f16_jsb_b200 - B200-native batched F-16 flight-dynamics environment.

Drop-in for the env-step path of Soham4001A/F16_JSB (jsbsim_gym/jsbsim_gym.py): hand-written sm_100a
CUDA kernels behind a C ABI (include/f16_b200.h), called through ctypes; PyTorch owns device memory.
"""
from .constants import (MAX_EPISODE_STEPS, NUM_FEATURES, NUM_STACKED_FRAMES, RADIUS, STATE_FORMAT,  # noqa: F401
                        normalize_angle_0_2pi, normalize_angle_mpi_pi, sample_goal_numpy)

__all__ = ["F16BatchedEnv", "F16VecEnv", "JSBSimEnv", "PositionReward", "wrap_jsbsim"]


def __getattr__(name):  # lazy: importing the package must not require torch/CUDA
    if name == "F16BatchedEnv":
        from .batched_env import F16BatchedEnv
        return F16BatchedEnv
    if name == "F16VecEnv":
        from .vec_env import F16VecEnv
        return F16VecEnv
    if name in ("JSBSimEnv", "PositionReward", "wrap_jsbsim"):
        from . import jsbsim_gym
        return getattr(jsbsim_gym, name)
    raise AttributeError(name)
