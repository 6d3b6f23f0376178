"""HostWindow - host-resident (N,10,15) observation windows fed with 60 B per env-step (include/f16_hostwin.h).

The reference returns the whole ten-frame stack from every step (jsbsim_gym/jsbsim_gym.py:150,235,263); nine of
those rows were already on the host. The step kernel's frame layout emits only the newest frame, the C library
DMA-s it into the next slot of a slot-major pinned ring whose pages are mapped twice back to back, and the
stacked observation is a zero-copy strided NumPy view of that ring. Everything numeric happens in
libf16b200.so; this class only wraps pointers.
"""
import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .constants import NUM_FEATURES, NUM_STACKED_FRAMES

PIN, NO_ALIAS, DMA_BOTH = 1, 2, 4
RECORD_DTYPE = np.dtype([("env", "<i4"), ("flags", "<i4"), ("ep_return", "<f4"), ("ep_len", "<i4"),
                         ("terminal_frame", "<f4", 16), ("reset_frame", "<f4", 16)])
assert RECORD_DTYPE.itemsize == C.sizeof(_lib.DoneRecord) == 144
FLAG_TRUNCATED, FLAG_CRASH, FLAG_GOAL = 1, 2, 4


def _as_array(ptr: int, shape, dtype) -> np.ndarray:
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if n == 0:
        return np.empty(shape, dtype=dtype)
    buf = (C.c_char * n).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class StepResult:
    """What one step (or reset) leaves in host memory. Arrays are views of library-owned buffers:
    `obs`, `reward`, `done`, `truncated`, `terminal_obs` stay valid until the step after next (n_rings=2;
    until the next step with n_rings=1), `records` is copied."""
    __slots__ = ("obs", "reward", "done", "truncated", "records", "terminal_obs", "ring", "first_slot")


class HostWindow:
    def __init__(self, num_envs: int, n_rings: int = 2, pin: bool = True, alias: bool = True, dma_both: bool = False):
        self.lib = _lib.load()
        self.num_envs = int(num_envs)
        self.n_rings = int(n_rings)
        h = C.c_void_p()
        flags = (PIN if pin else 0) | (0 if alias else NO_ALIAS) | (DMA_BOTH if dma_both else 0)
        _lib.check(self.lib.f16_hostwin_create(C.byref(h), self.num_envs, self.n_rings, flags), "f16_hostwin_create")
        self._h = h
        self._rings = []
        self.aliased = []
        for r in range(self.n_rings):
            base, pitch, slots, aliased = C.c_void_p(), C.c_int64(), C.c_int32(), C.c_int32()
            _lib.check(self.lib.f16_hostwin_layout(h, r, C.byref(base), C.byref(pitch), C.byref(slots), C.byref(aliased)), "f16_hostwin_layout")
            flat = _as_array(base.value, (slots.value * pitch.value // 4,), np.float32)
            self._rings.append((flat, pitch.value))
            self.aliased.append(bool(aliased.value))
        self.action_buffers = [_as_array(self.lib.f16_hostwin_action_buffer(h, k), (self.num_envs, 4), np.float32) for k in (0, 1)]
        self._action_addr = [(b, b.ctypes.data) for b in self.action_buffers]
        self._res = _lib.HostwinResult()
        # the library hands out the same few buffers over and over (22 window positions, two sets of scalars): build each
        # NumPy view once - at a few thousand envs constructing them was a third of a step's wall time
        self._views = {}

    def _view(self, ptr: int, shape, dtype) -> np.ndarray:
        key = (ptr, shape)
        v = self._views.get(key)
        if v is None:
            v = self._views[key] = _as_array(ptr, shape, dtype)
        return v

    def close(self, env=None):
        """env: the F16BatchedEnv this window served, if it is still alive - its done list points into this window's
        mapped memory and is un-pointed first."""
        if getattr(self, "_h", None) is not None and self._h.value:
            self._rings, self.action_buffers, self._views, self._action_addr = [], [], {}, []
            if env is not None and getattr(env, "_h", None) is not None and env._h.value:
                self.lib.f16_hostwin_detach(self._h, env._h)
            self.lib.f16_hostwin_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ views
    def window(self, ring: int, first_slot: int) -> np.ndarray:
        """(N,10,15) float32 strided view: row k of env n lives in slot first_slot + k."""
        key = ("window", ring, first_slot)
        v = self._views.get(key)
        if v is None:
            flat, pitch = self._rings[ring]
            v = self._views[key] = np.lib.stride_tricks.as_strided(flat[first_slot * (pitch // 4):],
                                                                   shape=(self.num_envs, NUM_STACKED_FRAMES, NUM_FEATURES),
                                                                   strides=(NUM_FEATURES * 4, pitch, 4))
        return v

    def gather(self, ring: int, first_slot: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """An owned, contiguous (N,10,15) copy of a window, assembled by the library's worker threads."""
        if out is None:
            out = np.empty((self.num_envs, NUM_STACKED_FRAMES, NUM_FEATURES), np.float32)
        assert out.flags.c_contiguous and out.dtype == np.float32 and out.shape == (self.num_envs, NUM_STACKED_FRAMES, NUM_FEATURES)
        _lib.check(self.lib.f16_hostwin_gather(self._h, int(ring), int(first_slot), C.c_void_p(out.ctypes.data)), "f16_hostwin_gather")
        return out

    def _result(self) -> StepResult:
        r, n = self._res, self.num_envs
        out = StepResult()
        out.ring, out.first_slot = int(r.ring), int(r.first_slot)
        out.obs = self.window(out.ring, out.first_slot)
        out.reward = self._view(r.reward, (n,), np.float32)
        out.done = self._view(r.done, (n,), np.uint8)
        out.truncated = self._view(r.truncated, (n,), np.uint8)
        k = int(r.n_done)
        out.records = _as_array(r.records, (k,), RECORD_DTYPE).copy() if (k and r.records) else np.empty(0, dtype=RECORD_DTYPE)
        out.terminal_obs = (_as_array(r.terminal_obs, (k, NUM_STACKED_FRAMES, NUM_FEATURES), np.float32) if (k and r.terminal_obs)
                            else np.empty((0, NUM_STACKED_FRAMES, NUM_FEATURES), np.float32))
        return out

    # ------------------------------------------------------------------ CUDA path
    def reset(self, env, stream_ptr) -> StepResult:
        _lib.check(self.lib.f16_hostwin_reset(self._h, env._h, stream_ptr, C.byref(self._res)), "f16_hostwin_reset")
        return self._result()

    def step(self, env, actions: np.ndarray, stream_ptr, auto_reset: bool = True) -> StepResult:
        assert actions.dtype == np.float32 and actions.flags.c_contiguous and actions.shape == (self.num_envs, 4)
        addr = None
        for b, a in self._action_addr:                     # the pinned staging buffers come back every other step
            if actions is b:
                addr = a
        if addr is None:
            addr = actions.ctypes.data
        _lib.check(self.lib.f16_hostwin_step(self._h, env._h, C.c_void_p(addr), int(auto_reset), stream_ptr,
                                             C.byref(self._res)), "f16_hostwin_step")
        return self._result()

    PHASES = ("enqueue", "wait_upload_kernel", "fix_older_slots", "wait_d2h", "wait_carry_over", "fix_newest_slot", "hand_off", "carry_over_duration")

    @property
    def numa_node(self) -> int:
        """NUMA node the ring lives on and the worker threads run on (-1: no placement); include/f16_hostwin.h."""
        return int(self.lib.f16_hostwin_numa_node(self._h))

    def timing(self, reset: bool = False) -> dict:
        """Average milliseconds per step spent in each phase of f16_hostwin_step since the last reset."""
        out = (C.c_double * 8)()
        _lib.check(self.lib.f16_hostwin_timing(self._h, out, int(reset)), "f16_hostwin_timing")
        return {k: 1e3 * v for k, v in zip(self.PHASES, out)}

    # ------------------------------------------------------------------ host-only path (tests, replay)
    def fill(self, frames: np.ndarray) -> StepResult:
        f = np.ascontiguousarray(frames, dtype=np.float32)
        assert f.shape == (self.num_envs, NUM_FEATURES)
        _lib.check(self.lib.f16_hostwin_fill(self._h, C.c_void_p(f.ctypes.data), C.byref(self._res)), "f16_hostwin_fill")
        return self._result()

    def push(self, frames: np.ndarray, reward: Optional[np.ndarray], done: Optional[np.ndarray], truncated: Optional[np.ndarray],
             records: np.ndarray) -> StepResult:
        f = np.ascontiguousarray(frames, dtype=np.float32)
        assert f.shape == (self.num_envs, NUM_FEATURES)
        rec = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        keep = [f, rec]

        def p(a, dt):
            if a is None:
                return C.c_void_p(0)
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return C.c_void_p(a.ctypes.data)

        _lib.check(self.lib.f16_hostwin_push(self._h, C.c_void_p(f.ctypes.data), p(reward, np.float32), p(done, np.uint8), p(truncated, np.uint8),
                                             C.c_void_p(rec.ctypes.data) if rec.size else C.c_void_p(0), int(rec.size), C.byref(self._res)),
                   "f16_hostwin_push")
        out = self._result()
        if rec.size:
            out.records = rec.copy()
        return out
