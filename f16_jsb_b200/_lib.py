"""ctypes binding of libf16b200.so (include/f16_b200.h). No CPU fallback: errors are loud."""
import ctypes as C
import os

from . import build as _build

_lib = None


class F16Error(RuntimeError):
    pass


class DoneRecord(C.Structure):
    """f16_done_record (include/f16_b200.h): one finished env of a frame-layout step."""
    _fields_ = [("env", C.c_int32), ("flags", C.c_int32), ("ep_return", C.c_float), ("ep_len", C.c_int32),
                ("terminal_frame", C.c_float * 16), ("reset_frame", C.c_float * 16)]


class HostwinResult(C.Structure):
    """f16_hostwin_result (include/f16_hostwin.h)."""
    _fields_ = [("ring", C.c_int32), ("first_slot", C.c_int32), ("n_done", C.c_int64), ("reward", C.c_void_p),
                ("done", C.c_void_p), ("truncated", C.c_void_p), ("records", C.c_void_p), ("terminal_obs", C.c_void_p)]


def load():
    """Load the CUDA library, building it in-tree first if the sources are newer."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("F16_B200_LIB") or _build.LIB_PATH   # override: tuning variants built side by side
    if path == _build.LIB_PATH and _build.needs_build():
        try:
            _build.build_library()
        except Exception as e:  # pragma: no cover - depends on toolchain
            if not os.path.exists(path):
                raise F16Error("libf16b200.so is missing and could not be built (%s); "
                               "there is no CPU fallback for the F-16 env" % e) from e
            import warnings
            warnings.warn("libf16b200.so is OLDER than its sources and rebuilding it failed (%s): running the stale "
                          "library" % e, RuntimeWarning)
    L = C.CDLL(path)
    vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
    L.f16_create.argtypes = [C.POINTER(vp), i64, i32, i32]
    L.f16_destroy.argtypes = [vp]
    L.f16_state_bytes.argtypes = [vp]
    L.f16_state_bytes.restype = C.c_size_t
    L.f16_set_ground_reactions.argtypes = [vp, i32]
    L.f16_get_ground_reactions.argtypes = [vp]
    L.f16_bind.argtypes = [vp] + [vp] * 8
    L.f16_bind_ring.argtypes = [vp] + [vp] * 8
    L.f16_bind_frames.argtypes = [vp] + [vp] * 7
    L.f16_set_done_list.argtypes = [vp, vp, vp]
    L.f16_obs_window.argtypes = [vp, C.POINTER(i32)]
    L.f16_hostwin_create.argtypes = [C.POINTER(vp), i64, i32, i32]
    L.f16_hostwin_destroy.argtypes = [vp]
    L.f16_hostwin_detach.argtypes = [vp, vp]
    L.f16_hostwin_gather.argtypes = [vp, i32, i32, vp]
    L.f16_hostwin_layout.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.f16_hostwin_action_buffer.argtypes = [vp, i32]
    L.f16_hostwin_action_buffer.restype = vp
    L.f16_hostwin_reset.argtypes = [vp, vp, vp, C.POINTER(HostwinResult)]
    L.f16_hostwin_step.argtypes = [vp, vp, vp, i32, vp, C.POINTER(HostwinResult)]
    L.f16_hostwin_timing.argtypes = [vp, vp, i32]
    L.f16_hostwin_numa_node.argtypes = [vp]
    L.f16_hostwin_fill.argtypes = [vp, vp, C.POINTER(HostwinResult)]
    L.f16_hostwin_push.argtypes = [vp, vp, vp, vp, vp, vp, i64, C.POINTER(HostwinResult)]
    L.f16_reset.argtypes = [vp, vp, vp, u64, vp]
    L.f16_reset_carryover.argtypes = [vp, vp, vp, u64, vp, vp]
    L.f16_step.argtypes = [vp, vp, i32, vp]
    L.f16_step_begin.argtypes = [vp, vp]
    L.f16_step_range.argtypes = [vp, vp, i32, i64, i64, vp]
    L.f16_step_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    L.f16_set_env_id_base.argtypes = [vp, i64]
    L.f16_get_state.argtypes = [vp, i64, vp, i32]
    L.f16_set_state.argtypes = [vp, i64, vp, i32]
    L.f16_pack_states.argtypes = [vp, vp, vp]
    L.f16_unpack_states.argtypes = [vp, vp, vp]
    L.f16_set_env_step.argtypes = [vp, i64, C.c_int32]
    L.f16_get_snapshot.argtypes = [vp, vp, vp]
    L.f16_get_stats.argtypes = [vp, vp, i32, vp]
    L.f16_stats_device_ptr.argtypes = [vp, C.POINTER(vp)]
    L.f16_rollout_add.argtypes = [i64, i64, i64] + [vp] * 15
    L.f16_rollout_gae.argtypes = [i64, i64, C.c_float, C.c_float] + [vp] * 8
    L.f16_rollout_gather.argtypes = [i64, i64, i64] + [vp] * 15
    L.f16_rollout_park_truncated.argtypes = [i64, i64, i64] + [vp] * 6
    L.f16_rollout_park_truncated.restype = i32
    L.f16_features17.argtypes = [i64, vp, vp, vp]
    L.f16_lma_attention_forward.argtypes = [i64, i32, i32, i32, vp, vp, C.c_float, u64, vp]
    L.f16_lma_attention_backward.argtypes = [i64, i32, i32, i32, vp, vp, vp, C.c_float, u64, vp]
    L.f16_lma_attention_mask.argtypes = [i64, i32, i32, vp, C.c_float, u64, vp]
    L.f16_lma_layernorm_forward.argtypes = [i64, i32, vp, vp, vp, C.c_float, vp, vp]
    L.f16_lma_layernorm_backward.argtypes = [i64, i32, vp, vp, vp, C.c_float, vp, vp, vp, vp]
    L.f16_lma_linear_wgrad.argtypes = [i64, i32, i32, vp, vp, vp, vp, vp]
    L.f16_lma_linear_forward.argtypes = [i64, i32, i32, vp, vp, vp, vp, vp]
    L.f16_lma_linear_supported.argtypes = [i32, i32]
    L.f16_lma_linear_wgrad_tc.argtypes = [i64, i32, i32, vp, vp, vp, vp, vp]
    L.f16_lma_linear_wgrad_tc_supported.argtypes = [i32, i32]
    L.f16_lma_embed_act_forward.argtypes = [i64, i32, i32, i32, vp, vp, vp, C.c_float, u64, vp]
    L.f16_lma_embed_act_backward.argtypes = [i64, i32, i32, i32, vp, vp, vp, C.c_float, u64, vp]
    L.f16_lma_dropout_add_forward.argtypes = [i64, vp, vp, vp, C.c_float, u64, vp]
    L.f16_lma_dropout_backward.argtypes = [i64, vp, vp, C.c_float, u64, vp]
    L.f16_lma_policy_packed_size.argtypes = []
    L.f16_lma_policy_packed_size.restype = i64
    L.f16_lma_policy_entries.argtypes = []
    L.f16_lma_policy_entries.restype = i32
    L.f16_lma_policy_entry.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
    L.f16_lma_policy_entry.restype = i32
    L.f16_lma_policy_forward.argtypes = [i64, vp, vp, i64] + [vp] * 10
    L.f16_lma_policy_forward.restype = i32
    for name in ("f16_lma_attention_forward", "f16_lma_attention_backward", "f16_lma_attention_mask", "f16_lma_layernorm_forward",
                 "f16_lma_layernorm_backward", "f16_lma_linear_wgrad", "f16_lma_linear_forward", "f16_lma_linear_supported",
                 "f16_lma_linear_wgrad_tc", "f16_lma_linear_wgrad_tc_supported", "f16_lma_embed_act_forward",
                 "f16_lma_embed_act_backward", "f16_lma_dropout_add_forward", "f16_lma_dropout_backward"):
        getattr(L, name).restype = i32
    L.f16_features17.restype = i32
    for name in ("f16_rollout_add", "f16_rollout_gae", "f16_rollout_gather"):
        getattr(L, name).restype = i32
    L.f16_launch_count.restype = i64
    L.f16_num_state_fields.restype = i32
    L.f16_last_error.restype = C.c_char_p
    L.f16_version.restype = C.c_char_p
    for name in HOSTWIN_SYMBOLS:
        if name != "f16_hostwin_action_buffer":
            getattr(L, name).restype = i32
    for name in ("f16_create", "f16_destroy", "f16_bind", "f16_bind_ring", "f16_bind_frames", "f16_set_done_list", "f16_obs_window", "f16_reset", "f16_reset_carryover", "f16_step", "f16_step_begin", "f16_step_range", "f16_step_host",
                 "f16_set_env_id_base", "f16_get_state", "f16_set_state", "f16_pack_states",
                 "f16_unpack_states", "f16_set_env_step", "f16_get_snapshot", "f16_get_stats",
                 "f16_stats_device_ptr"):
        getattr(L, name).restype = i32
    _lib = L
    return L


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """`with torch.cuda.device(device)` only when `device` is not already current: with one process per GPU it always is, and the
    context manager costs several microseconds on each of the ~60 kernel calls of an update step."""
    import torch
    idx = device.index
    if idx is None or torch.cuda.current_device() == idx:
        return _NO_GUARD
    return torch.cuda.device(idx)


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().f16_last_error().decode(errors="replace")
        raise F16Error("%s failed: %s" % (what or "f16 call", msg))


EXPORTED_SYMBOLS = (
    "f16_create", "f16_destroy", "f16_state_bytes", "f16_set_ground_reactions", "f16_get_ground_reactions", "f16_bind", "f16_bind_ring", "f16_bind_frames", "f16_set_done_list", "f16_obs_window", "f16_reset", "f16_reset_carryover", "f16_step", "f16_step_begin", "f16_step_range", "f16_step_host",
    "f16_set_env_id_base", "f16_get_state", "f16_set_state", "f16_pack_states", "f16_unpack_states",
    "f16_set_env_step", "f16_get_snapshot", "f16_get_stats", "f16_stats_device_ptr", "f16_launch_count", "f16_num_state_fields",
    "f16_last_error", "f16_version")
HOSTWIN_SYMBOLS = ("f16_hostwin_create", "f16_hostwin_destroy", "f16_hostwin_detach", "f16_hostwin_gather", "f16_hostwin_layout", "f16_hostwin_action_buffer", "f16_hostwin_reset",
                   "f16_hostwin_step", "f16_hostwin_timing", "f16_hostwin_numa_node", "f16_hostwin_fill", "f16_hostwin_push")
ROLLOUT_SYMBOLS = ("f16_rollout_add", "f16_rollout_gae", "f16_rollout_gather", "f16_rollout_park_truncated", "f16_features17")
