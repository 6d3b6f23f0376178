import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_env_traces.npz"))
    traces = {}
    for key in z.files:
        if "/" not in key:
            continue
        name, field = key.split("/")
        traces.setdefault(name, {})[field] = z[key]
    return traces


@pytest.fixture(scope="session")
def state_fields():
    txt = open(os.path.join(ROOT, "include", "f16_state_fields.h")).read()
    body = txt.split("enum f16_state_field")[1].split("F16_NUM_STATE_FIELDS")[0]
    names = []
    for m in re.finditer(r"F16S_([A-Z0-9_]+)", body):
        if m.group(1) not in names:
            names.append(m.group(1))
    return names


@pytest.fixture(scope="session")
def oracle():
    from oracle import f16_oracle
    f16_oracle.lib()
    return f16_oracle


class HostSim:
    """g++ build of the kernel's per-env source (tests/hostsim) - debugging harness for CPU tests."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "hostsim")
        so = os.path.join(d, "libf16hostsim.so")
        srcs = [os.path.join(d, "f16_hostsim.cpp")] + [os.path.join(ROOT, "f16_jsb_b200", "csrc", f) for f in
                                                      ("f16_model.cuh", "f16_ground.cuh", "f16_env.cuh", "f16_host_setup.h", "f16_model_data.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-D_GNU_SOURCE",
                                   "-o", so, srcs[0]])
        L = C.CDLL(so)
        L.hs_env_create.restype = C.c_void_p
        L.hs_env_create.argtypes = [C.c_int]
        L.hs_env_destroy.argtypes = [C.c_void_p]
        L.hs_env_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_env_reset_carryover.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
        L.hs_env_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.hs_env_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.hs_snapshot.argtypes = [C.c_void_p]
        L.hs_mass_set.argtypes = [C.c_int, C.c_void_p]
        self.L = L

    def snapshot(self, nf):
        out = np.zeros(nf + 12)
        self.L.hs_snapshot(out.ctypes.data)
        return out[:nf], out[nf:]

    def env(self, mode):
        return HostSimEnv(self.L, mode)


class HostSimEnv:
    def __init__(self, L, mode):
        self.L, self.h = L, L.hs_env_create(mode)

    def __del__(self):
        self.L.hs_env_destroy(self.h)

    def reset(self, goal):
        g = np.ascontiguousarray(goal, dtype=np.float32)
        obs = np.zeros((10, 15), np.float32)
        self.L.hs_env_reset(self.h, g.ctypes.data, obs.ctypes.data)
        return obs

    def reset_carryover(self, goal, last_action):
        g = np.ascontiguousarray(goal, dtype=np.float32)
        a = np.ascontiguousarray(last_action, dtype=np.float32)
        obs = np.zeros((10, 15), np.float32)
        self.L.hs_env_reset_carryover(self.h, g.ctypes.data, a.ctypes.data, obs.ctypes.data)
        return obs

    def step(self, action, auto_reset=0, seed=0, env_id=0):
        a = np.ascontiguousarray(action, dtype=np.float32)
        obs = np.zeros((10, 15), np.float32)
        tobs = np.zeros((10, 15), np.float32)
        r = C.c_float()
        fl = self.L.hs_env_step(self.h, a.ctypes.data, auto_reset, seed, env_id, obs.ctypes.data, C.byref(r), tobs.ctypes.data)
        return obs, np.float32(r.value), fl, tobs

    def get_state(self, nf):
        s = np.zeros(nf)
        self.L.hs_env_get_state(self.h, s.ctypes.data)
        return s

    def set_state(self, packed, current_step):
        p = np.ascontiguousarray(packed, dtype=np.float64)
        self.L.hs_env_set_state(self.h, p.ctypes.data, int(current_step))


@pytest.fixture(scope="session")
def hostsim():
    return HostSim()


# per-field floors for relative state errors: |x - ref| / max(|ref|, floor). Floors are the natural
# scale below which a field is "zero": angles/rates/actuators 1e-3 (rad, rad/s, norm), linear
# accelerations 1 ft/s2 (0.03 g), pilot load factors 1e-2 g, positions/velocities 1 ft, 1 ft/s.
def state_floors(names):
    f = np.full(len(names), 1e-3)
    for i, n in enumerate(names):
        if n.startswith(("AI0", "AI1", "ABODY")):
            f[i] = 1.0
        elif n in ("NPY", "NPZ"):
            f[i] = 1e-2
        elif n.startswith(("RI", "VI")):
            f[i] = 1.0
        elif n.startswith("WDOT"):
            f[i] = 1e-2
        elif n in ("VC_KTS", "VG", "N2"):
            f[i] = 1.0
    return f
