"""Empty stub: jsbsim_gym/visualization/rendering.py:1 imports pygame at module top."""
