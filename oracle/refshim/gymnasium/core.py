"""gymnasium.core look-alike (stable_baselines3/common/monitor.py:12 imports the type variables)."""
from typing import TypeVar

from . import ActionWrapper, Env, ObservationWrapper, RewardWrapper, Wrapper  # noqa: F401

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")
