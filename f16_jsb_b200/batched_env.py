"""F16BatchedEnv - N F-16 environments resident on one B200, stepped by libf16b200.so.

PyTorch only owns the device memory and the stream; every numeric operation is in the CUDA library
(f16_jsb_b200/csrc). The torch-tensor API (`reset`, `step`) never leaves the device; `step_host` is
the NumPy-in / NumPy-out call with the host<->device copies inside (what SB3's stock loops use).
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from .constants import NUM_FEATURES, NUM_STACKED_FRAMES

MODE_FP64, MODE_FP32 = 0, 1
NUM_STATS = 8
STAT_NAMES = ("episodes", "return_sum", "length_sum", "crashes", "goals", "truncations", "env_steps", "ground_redos")


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class F16BatchedEnv:
    """Batched replacement of JSBSimEnv+PositionReward (jsbsim_gym/jsbsim_gym.py:95-533).

    mode: "fp64" (parity: all model math in double) or "fp32" (throughput: float math, double
    kinematic state). Observations are (N, 10, 15) float32, row 0 oldest, row 9 newest.

    ground_reactions: None keeps the mode's default - on in "fp64", off in "fp32". JSBSim's ground contacts
    (aircraft/f16/f16.xml:85-215) can only act inside the last env-step of an episode that ends in a crash;
    with them on, such a step is redone by a cold copy of the step that includes the contact and friction
    forces (one crash in fifteen under random actions; see DESIGN.md for what it costs).

    obs_layout: where the ten-frame window of each env lives.
    "ring" (default) keeps a slot-major ring buffer (20, N, 15) that the step kernel writes: the newest frames go into
    their slot and into the mirror slot ten further on as two contiguous (N, 15) planes, so `obs` is always a zero-copy
    strided (N, 10, 15) view of the current window (row 0 oldest, row 9 newest; strides (15, N*15, 1)) - the reference's
    values with 120 B written per env-step instead of 540 B read + 600 B written. obs[:, k, :] is contiguous over the
    envs (what a GPU consumer wants); obs.reshape(N, 150) copies. Like the reference's array the view is only valid
    until the next step. "stacked" keeps the reference's contiguous (N, 10, 15)
    tensor and shifts it in place every step. "frame" keeps no history on the device: `obs` is the (N, 15) tensor of
    newest frames, 60 B written per env-step; the windows then live in host memory (host_window.py, F16VecEnv's
    default) or in the frame-only rollout store (rollout.py).

    reset_mode: "snapshot" (default) starts every episode from the canonical state of a FRESH reference env
    object. "carryover" does what JSBSimEnv.reset does to an env object that already exists
    (jsbsim_gym.py:305-306: run_ic() + set-running on top of the actuator positions, control-law memories, stale
    air data and accelerations the last episode left behind; include/f16_b200.h, f16_reset_carryover) - the
    second and later episodes of one reference env. It turns ground reactions on (both belong to the
    reference-detail build of the step kernel).
    """

    def __init__(self, num_envs: int, device=None, mode: str = "fp32", seed: int = 0,
                 with_terminal_obs: bool = True, env_id_base: int = 0, obs_layout: str = "ring",
                 done_list: bool = False, ground_reactions=None, reset_mode: str = "snapshot"):
        if not torch.cuda.is_available():
            raise _lib.F16Error("F16BatchedEnv needs a CUDA device: the F-16 env has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.F16Error("F16BatchedEnv only runs on CUDA devices (got %s)" % self.device)
        self.num_envs = int(num_envs)
        self.mode = {"fp64": MODE_FP64, "fp32": MODE_FP32}[mode]
        self.mode_name = mode
        self.seed = int(seed)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        _lib.check(self.lib.f16_create(C.byref(h), self.num_envs, dev_index, self.mode), "f16_create")
        self._h = h
        if reset_mode not in ("snapshot", "carryover"):
            raise ValueError("reset_mode must be 'snapshot' or 'carryover'")
        self.reset_mode = reset_mode
        self._last_actions = None      # carry-over reset: the action of the last step (fcs/*-cmd-norm stay set)
        if reset_mode == "carryover":
            if ground_reactions is not None and not ground_reactions:
                raise ValueError("reset_mode='carryover' needs ground_reactions on")
            ground_reactions = True
        # ground reactions (include/f16_b200.h): None keeps the mode's default (on in fp64, off in fp32)
        if ground_reactions is not None:
            _lib.check(self.lib.f16_set_ground_reactions(h, 1 if ground_reactions else 0), "f16_set_ground_reactions")
        self.ground_reactions = bool(self.lib.f16_get_ground_reactions(h))
        n = self.num_envs
        with torch.cuda.device(self.device):
            self.state = torch.zeros(self.lib.f16_state_bytes(h), dtype=torch.uint8, device=self.device)
            if obs_layout not in ("stacked", "ring", "frame"):
                raise ValueError("obs_layout must be 'stacked', 'ring' or 'frame'")
            self.obs_layout = obs_layout
            if obs_layout == "frame":
                self._obs_buf = torch.zeros((n, NUM_FEATURES), dtype=torch.float32, device=self.device)
                with_terminal_obs = False
            elif obs_layout == "ring":
                self._obs_buf = torch.zeros((2 * NUM_STACKED_FRAMES, n, NUM_FEATURES), dtype=torch.float32, device=self.device)   # [slot][env][15]
            else:
                self._obs_buf = torch.zeros((n, NUM_STACKED_FRAMES, NUM_FEATURES), dtype=torch.float32, device=self.device)
            self.reward = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.done = torch.zeros(n, dtype=torch.uint8, device=self.device)
            self.truncated = torch.zeros(n, dtype=torch.uint8, device=self.device)
            self.terminal_obs = (torch.zeros((n, NUM_STACKED_FRAMES, NUM_FEATURES), dtype=torch.float32, device=self.device)
                                 if with_terminal_obs else None)
            self.ep_return = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.ep_len = torch.zeros(n, dtype=torch.int32, device=self.device)
        self.done_records = self.done_count = None
        if obs_layout == "frame":
            if done_list:
                # f16_done_record[N] (36 words each: env, flags, ep_return, ep_len, terminal frame[16], reset frame[16])
                with torch.cuda.device(self.device):
                    self.done_records = torch.zeros((n, 36), dtype=torch.int32, device=self.device)
                    self.done_count = torch.zeros(1, dtype=torch.int32, device=self.device)
            _lib.check(self.lib.f16_bind_frames(h, _ptr(self.state), _ptr(self._obs_buf), _ptr(self.reward), _ptr(self.done),
                                                _ptr(self.truncated), _ptr(self.done_records), _ptr(self.done_count)), "f16_bind_frames")
        else:
            bind = self.lib.f16_bind_ring if obs_layout == "ring" else self.lib.f16_bind
            _lib.check(bind(h, _ptr(self.state), _ptr(self._obs_buf), _ptr(self.reward), _ptr(self.done),
                            _ptr(self.truncated), _ptr(self.terminal_obs), _ptr(self.ep_return),
                            _ptr(self.ep_len)), "f16_bind")
        if env_id_base:
            _lib.check(self.lib.f16_set_env_id_base(h, int(env_id_base)), "f16_set_env_id_base")
        sp = C.c_void_p()
        _lib.check(self.lib.f16_stats_device_ptr(h, C.byref(sp)), "f16_stats_device_ptr")
        self._stats_ptr = sp.value

    @property
    def obs(self) -> torch.Tensor:
        """Current stacked observations (N, 10, 15): the bound tensor itself (stacked layout) or the
        window view of the ring (ring layout; a new view after every step). Frame layout: the newest
        frames (N, 15)."""
        if self.obs_layout in ("stacked", "frame"):
            return self._obs_buf
        first = C.c_int()
        _lib.check(self.lib.f16_obs_window(self._h, C.byref(first)), "f16_obs_window")
        return self._obs_buf[first.value:first.value + NUM_STACKED_FRAMES].permute(1, 0, 2)

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.f16_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ device-resident API
    def reset(self, mask: Optional[torch.Tensor] = None, goals: Optional[torch.Tensor] = None,
              seed: Optional[int] = None) -> torch.Tensor:
        """Reset masked envs (all if mask is None). goals: (N,3) float32 cuda tensor or None (Philox)."""
        if seed is not None:
            self.seed = int(seed)
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        if goals is not None:
            goals = goals.to(device=self.device, dtype=torch.float32).contiguous()
            assert goals.shape == (self.num_envs, 3)
        if self.reset_mode == "carryover":
            _lib.check(self.lib.f16_reset_carryover(self._h, _ptr(mask), _ptr(goals), self.seed, _ptr(self._last_actions),
                                                    self._stream()), "f16_reset_carryover")
        else:
            _lib.check(self.lib.f16_reset(self._h, _ptr(mask), _ptr(goals), self.seed, self._stream()), "f16_reset")
        return self.obs

    def step(self, actions: Optional[torch.Tensor], auto_reset: bool = True):
        """One env-step on the current CUDA stream; returns views of the bound device tensors
        (obs, reward, done, truncated). actions: (N,4) float32 cuda tensor, or None for in-kernel
        uniform random actions."""
        if actions is not None:
            if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
                actions = actions.to(device=self.device, dtype=torch.float32).contiguous()
            assert actions.shape == (self.num_envs, 4)
        mode = int(auto_reset)
        if self.reset_mode == "carryover":
            mode *= 2                      # F16_AUTO_RESET_CARRYOVER
            if actions is not None:
                if self._last_actions is None:
                    self._last_actions = torch.zeros((self.num_envs, 4), dtype=torch.float32, device=self.device)
                self._last_actions.copy_(actions)
        _lib.check(self.lib.f16_step(self._h, _ptr(actions), mode, self._stream()), "f16_step")
        return self.obs, self.reward, self.done, self.truncated

    # ------------------------------------------------------------------ host-buffer API
    def step_host(self, actions: np.ndarray, obs_out: Optional[np.ndarray], reward_out: Optional[np.ndarray],
                  done_out: Optional[np.ndarray], truncated_out: Optional[np.ndarray], auto_reset: bool = True):
        """NumPy in / NumPy out; host->device and device->host copies and a stream sync inside."""
        assert actions.dtype == np.float32 and actions.flags.c_contiguous and actions.shape == (self.num_envs, 4)

        def p(a):
            return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)

        mode = 1 if auto_reset else 0
        if self.reset_mode == "carryover":
            mode *= 2                      # F16_AUTO_RESET_CARRYOVER
            self._last_actions = torch.from_numpy(actions).to(self.device, non_blocking=True)   # fcs/*-cmd-norm stay set across reset()
        _lib.check(self.lib.f16_step_host(self._h, p(actions), mode, p(obs_out), p(reward_out), p(done_out),
                                          p(truncated_out), self._stream()), "f16_step_host")

    # ------------------------------------------------------------------ parity / debugging
    @property
    def num_state_fields(self) -> int:
        return int(self.lib.f16_num_state_fields())

    def get_state(self, env: int) -> np.ndarray:
        out = np.zeros(self.num_state_fields, dtype=np.float64)
        _lib.check(self.lib.f16_get_state(self._h, int(env), C.c_void_p(out.ctypes.data), out.size), "f16_get_state")
        return out

    def set_state(self, env: int, packed: np.ndarray, current_step: Optional[int] = None) -> None:
        a = np.ascontiguousarray(packed, dtype=np.float64)
        _lib.check(self.lib.f16_set_state(self._h, int(env), C.c_void_p(a.ctypes.data), a.size), "f16_set_state")
        if current_step is not None:
            _lib.check(self.lib.f16_set_env_step(self._h, int(env), int(current_step)), "f16_set_env_step")

    def pack_states(self) -> torch.Tensor:
        out = torch.empty((self.num_envs, self.num_state_fields), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.f16_pack_states(self._h, _ptr(out), self._stream()), "f16_pack_states")
        return out

    def unpack_states(self, packed: torch.Tensor) -> None:
        packed = packed.to(device=self.device, dtype=torch.float64).contiguous()
        assert packed.shape == (self.num_envs, self.num_state_fields)
        _lib.check(self.lib.f16_unpack_states(self._h, _ptr(packed), self._stream()), "f16_unpack_states")
        torch.cuda.current_stream(self.device).synchronize()

    def snapshot(self):
        st = np.zeros(self.num_state_fields, dtype=np.float64)
        props = np.zeros(12, dtype=np.float64)
        _lib.check(self.lib.f16_get_snapshot(self._h, C.c_void_p(st.ctypes.data), C.c_void_p(props.ctypes.data)), "f16_get_snapshot")
        return st, props

    # ------------------------------------------------------------------ statistics
    def stats(self, reset: bool = False) -> dict:
        out = np.zeros(NUM_STATS, dtype=np.float64)
        _lib.check(self.lib.f16_get_stats(self._h, C.c_void_p(out.ctypes.data), int(reset), self._stream()), "f16_get_stats")
        return dict(zip(STAT_NAMES, out.tolist()))

    def launch_count(self) -> int:
        return int(self.lib.f16_launch_count())
