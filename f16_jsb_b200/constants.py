"""Constants of the reference env layer, same names and values as jsbsim_gym/jsbsim_gym.py:12-58."""
import numpy as np

# jsbsim_gym/jsbsim_gym.py:12-25
STATE_FORMAT = [
    "position/lat-gc-rad", "position/long-gc-rad", "position/h-sl-meters", "velocities/mach",
    "aero/alpha-rad", "aero/beta-rad", "velocities/p-rad_sec", "velocities/q-rad_sec",
    "velocities/r-rad_sec", "attitude/phi-rad", "attitude/theta-rad", "attitude/psi-rad",
]
EPSILON = 1e-5  # jsbsim_gym.py:28
_PI = np.pi
# jsbsim_gym.py:32-53
SINGLE_OBS_LOW = np.array([-np.inf, -np.inf, -np.inf, 0, -_PI - EPSILON, -_PI - EPSILON, -np.inf, -np.inf, -np.inf,
                           -_PI - EPSILON, -_PI / 2 - EPSILON, -_PI - EPSILON, -np.inf, -np.inf, 0], dtype=np.float32)
SINGLE_OBS_HIGH = np.array([np.inf, np.inf, np.inf, np.inf, _PI + EPSILON, _PI + EPSILON, np.inf, np.inf, np.inf,
                            _PI + EPSILON, _PI / 2 + EPSILON, _PI + EPSILON, np.inf, np.inf, np.inf], dtype=np.float32)
RADIUS = 6.3781e6           # jsbsim_gym.py:56
NUM_STACKED_FRAMES = 10     # jsbsim_gym.py:58
NUM_FEATURES = 15
ACTION_LOW = np.array([-1, -1, -1, 0], dtype=np.float32)    # jsbsim_gym.py:143-148
ACTION_HIGH = np.array([1, 1, 1, 1], dtype=np.float32)
DOWN_SAMPLE = 4             # jsbsim_gym.py:157
MAX_EPISODE_STEPS = 1200    # jsbsim_gym.py:159, 541
GOAL_RADIUS_M = 100.0       # jsbsim_gym.py:163
REWARD_GAIN = 1e-2          # jsbsim_gym.py:532


def normalize_angle_mpi_pi(angle_rad: float) -> float:
    """jsbsim_gym.py:60-78 (host-side helper; the kernel does the same in float32)."""
    if np.isnan(angle_rad) or np.isinf(angle_rad):
        return 0.0
    angle_rad = angle_rad % (2 * np.pi)
    if angle_rad >= np.pi:
        angle_rad -= 2 * np.pi
    return angle_rad


def normalize_angle_0_2pi(angle_rad: float) -> float:
    """jsbsim_gym.py:80-93."""
    return angle_rad % (2 * np.pi)


def sample_goal_numpy(seed) -> np.ndarray:
    """Goal draw of JSBSimEnv.reset (jsbsim_gym.py:312-323): PCG64 default_rng(seed), three uniforms."""
    rng = np.random.default_rng(seed)
    distance_m = rng.uniform(1000.0, 10000.0)
    bearing_rad = rng.uniform(0, 2 * np.pi)
    altitude_m = rng.uniform(1000.0, 4000.0)
    g = np.zeros(3, dtype=np.float32)
    g[0] = distance_m * np.cos(bearing_rad)
    g[1] = distance_m * np.sin(bearing_rad)
    g[2] = altitude_m
    return g
