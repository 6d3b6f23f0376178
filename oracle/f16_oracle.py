"""ctypes binding of the CPU oracle (oracle/libf16oracle.so). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under f16_jsb_b200/ does. Two handle types are exposed:

* ``OracleFDM``  - FGFDMExec-shaped (set/get_property_value, run_ic, run): what the reference env
  calls on ``jsbsim.FGFDMExec`` (jsbsim_gym/jsbsim_gym.py:151-155,168-170,182,219-232,305-306).
* ``OracleEnv``  - JSBSimEnv + PositionReward restated in C (jsbsim_gym/jsbsim_gym.py:172-331,470-519).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libf16oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("f16_oracle.cpp", "f16_oracle_gen.inc")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.f16o_fdm_create.restype = C.c_void_p
        L.f16o_fdm_destroy.argtypes = [C.c_void_p]
        L.f16o_fdm_set_property.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.f16o_fdm_get_property.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
        L.f16o_fdm_run_ic.argtypes = [C.c_void_p]
        L.f16o_fdm_run.argtypes = [C.c_void_p]
        L.f16o_fdm_pack_state.argtypes = [C.c_void_p, C.c_void_p]
        L.f16o_fdm_unpack_state.argtypes = [C.c_void_p, C.c_void_p]
        L.f16o_env_create.restype = C.c_void_p
        L.f16o_env_destroy.argtypes = [C.c_void_p]
        L.f16o_env_fdm.restype = C.c_void_p
        L.f16o_env_fdm.argtypes = [C.c_void_p]
        L.f16o_env_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.f16o_env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.f16o_rollout.restype = C.c_longlong
        L.f16o_rollout.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
        L.f16o_batch_trajectory.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_int]
        _lib = L
    return _lib


class OracleFDM:
    """FGFDMExec look-alike backed by the oracle."""

    def __init__(self, handle=None, owner=None):
        self._owned = handle is None
        self._h = lib().f16o_fdm_create() if handle is None else handle
        self._owner = owner  # keeps the env alive when this is a borrowed handle

    def __del__(self):
        if getattr(self, "_owned", False) and self._h:
            lib().f16o_fdm_destroy(self._h)
            self._h = None

    def set_property_value(self, name: str, value: float) -> None:
        if lib().f16o_fdm_set_property(self._h, name.encode(), float(value)) != 0:
            raise KeyError("oracle FDM has no settable property %r" % name)

    def get_property_value(self, name: str) -> float:
        out = C.c_double()
        if lib().f16o_fdm_get_property(self._h, name.encode(), C.byref(out)) != 0:
            raise KeyError("oracle FDM has no property %r" % name)
        return out.value

    def __getitem__(self, name):
        return self.get_property_value(name)

    def __setitem__(self, name, value):
        self.set_property_value(name, value)

    def run_ic(self) -> bool:
        return lib().f16o_fdm_run_ic(self._h) == 0

    def run(self) -> bool:
        return lib().f16o_fdm_run(self._h) == 0

    def pack_state(self) -> np.ndarray:
        n = lib().f16o_num_state_fields()
        out = np.zeros(n, dtype=np.float64)
        lib().f16o_fdm_pack_state(self._h, out.ctypes.data)
        return out


    def unpack_state(self, packed) -> None:
        """Overwrite the packed fields (include/f16_state_fields.h); the FDM must be in flight configuration."""
        p = np.ascontiguousarray(packed, dtype=np.float64)
        assert p.shape == (lib().f16o_num_state_fields(),)
        lib().f16o_fdm_unpack_state(self._h, p.ctypes.data)


class OracleEnv:
    """JSBSimEnv wrapped in PositionReward(gain=1e-2), one env, goals supplied by the caller."""

    def __init__(self):
        self._h = lib().f16o_env_create()
        self.fdm = OracleFDM(lib().f16o_env_fdm(self._h), owner=self)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().f16o_env_destroy(self._h)
            self._h = None

    def reset(self, goal) -> np.ndarray:
        g = np.ascontiguousarray(goal, dtype=np.float32)
        obs = np.zeros((10, 15), dtype=np.float32)
        lib().f16o_env_reset(self._h, g.ctypes.data, obs.ctypes.data)
        return obs

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float32)
        obs = np.zeros((10, 15), dtype=np.float32)
        r = C.c_float()
        fl = lib().f16o_env_step(self._h, a.ctypes.data, obs.ctypes.data, C.byref(r))
        return obs, np.float32(r.value), bool(fl & 1), bool(fl & 2)


def sample_goal(seed) -> np.ndarray:
    """Goal draw of JSBSimEnv.reset (jsbsim_gym/jsbsim_gym.py:312-323), bit-identical on the host."""
    rng = np.random.default_rng(seed)
    distance_m = rng.uniform(1000.0, 10000.0)
    bearing_rad = rng.uniform(0, 2 * np.pi)
    altitude_m = rng.uniform(1000.0, 4000.0)
    g = np.zeros(3, dtype=np.float32)
    g[0] = distance_m * np.cos(bearing_rad)
    g[1] = distance_m * np.sin(bearing_rad)
    g[2] = altitude_m
    return g


def rollout(n_envs: int, n_steps: int, seed: int = 0, n_threads: int = 1):
    """Bounded random-action rollout on host threads; returns (env_steps, reward checksum)."""
    cs = C.c_double()
    n = lib().f16o_rollout(int(n_envs), int(n_steps), int(seed), int(n_threads), C.byref(cs))
    return int(n), cs.value


def batch_trajectory(goals: np.ndarray, actions: np.ndarray, n_threads: int = 0, want_states: bool = False):
    """goals (N,3) f32, actions (T,N,4) f32 -> frames (T,N,15) f32, rewards (T,N) f32, flags (T,N) u8
    [, final packed states (N,NF) f64]; no auto-reset, stepping of an env stops at its first done."""
    goals = np.ascontiguousarray(goals, dtype=np.float32)
    actions = np.ascontiguousarray(actions, dtype=np.float32)
    T, N = actions.shape[0], actions.shape[1]
    assert goals.shape == (N, 3) and actions.shape == (T, N, 4)
    frames = np.zeros((T, N, 15), dtype=np.float32)
    rewards = np.zeros((T, N), dtype=np.float32)
    flags = np.zeros((T, N), dtype=np.uint8)
    nf = lib().f16o_num_state_fields()
    states = np.zeros((N, nf), dtype=np.float64) if want_states else None
    if n_threads <= 0:
        n_threads = min(os.cpu_count() or 1, max(1, N))
    lib().f16o_batch_trajectory(N, T, goals.ctypes.data, actions.ctypes.data, frames.ctypes.data, rewards.ctypes.data,
                                flags.ctypes.data, states.ctypes.data if want_states else None, int(n_threads))
    return (frames, rewards, flags, states) if want_states else (frames, rewards, flags)
