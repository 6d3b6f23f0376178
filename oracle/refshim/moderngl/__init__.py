"""Empty stub: jsbsim_gym/visualization/rendering.py:3 imports moderngl at module top."""


class Context:  # used only in annotations
    pass
