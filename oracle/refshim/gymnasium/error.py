"""gymnasium.error look-alike (imported by stable_baselines3/common/vec_env/vec_video_recorder.py:6)."""


class Error(Exception):
    pass


class DependencyNotInstalled(Error):
    pass
