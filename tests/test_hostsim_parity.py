"""CPU-side parity of the KERNEL SOURCE (f16_model.cuh / f16_env.cuh compiled with g++ by
tests/hostsim) against the oracle. The same checks run against the real CUDA build in
test_gpu_parity.py; these catch arithmetic regressions in the container that has no GPU.

Tolerances (stated):
  FP64 mode, teacher-forced one env-step from an oracle state: <= 1e-9 relative per field
      (contract from north_star: <= 1e-6; the two codings differ only by summation order / FMA).
  FP32 mode, teacher-forced: <= 1e-3 relative per field with the per-field floors of conftest
      (measured worst case over both golden traces: 4e-4, on the linear accelerations).
"""
import numpy as np
import pytest

from conftest import state_floors

NF = 53


def rel_err(a, ref, floors):
    return np.abs(a - ref) / np.maximum(np.abs(ref), floors)


def test_snapshot_matches_oracle_reset_state(oracle, hostsim, state_fields):
    assert len(state_fields) == NF
    env = oracle.OracleEnv()
    env.reset(oracle.sample_goal(0))
    want = env.fdm.pack_state()
    got, props = hostsim.snapshot(NF)
    e = rel_err(got, want, state_floors(state_fields))
    assert e.max() < 1e-12, state_fields[int(e.argmax())]
    from f16_jsb_b200.constants import STATE_FORMAT
    for i, p in enumerate(STATE_FORMAT):
        assert props[i] == pytest.approx(env.fdm[p], abs=1e-13)


@pytest.mark.parametrize("mode,tol", [(0, 1e-9), (1, 1e-3)])
@pytest.mark.parametrize("name", ["random0", "gentle0"])
def test_teacher_forced_step_parity(hostsim, golden, state_fields, mode, tol, name):
    t = golden[name]
    floors = state_floors(state_fields)
    env = hostsim.env(mode)
    env.reset(t["goal"])
    worst = 0.0
    for k in range(0, len(t["actions"]), 7):
        env.set_state(t["states"][k], current_step=k)
        obs, r, fl, _ = env.step(t["actions"][k])
        e = rel_err(env.get_state(NF), t["states"][k + 1], floors)
        worst = max(worst, e.max())
        assert e.max() < tol, (k, state_fields[int(e.argmax())], e.max())
        if mode == 0:
            assert np.allclose(obs[-1][:12], t["frames"][k][:12], rtol=0, atol=1e-6 * np.maximum(1.0, np.abs(t["frames"][k][:12])))
    assert worst > 0.0


def check_free_run(step_fn, t, early_steps=300, early_tol=1e-5, late_tol=1e-1):
    """Free-running comparison against a golden trace. The airframe is open-loop unstable once the
    actuators saturate, so round-off level differences (summation order, FMA, libm) are amplified
    late in aggressive episodes: frames must agree to `early_tol` (relative, floor 1e-2) for the
    first `early_steps` steps and to `late_tol` afterwards; the episode must end within one step of
    the golden one with the same flags. Returns (max early error, max late error)."""
    n = len(t["actions"])
    e_early = e_late = 0.0
    for k in range(n):
        frame, reward, done, trunc = step_fn(t["actions"][k])
        f = t["frames"][k]
        e = float((np.abs(frame[:12] - f[:12]) / np.maximum(np.abs(f[:12]), 1e-2)).max())
        if k < early_steps:
            e_early = max(e_early, e)
            assert e <= early_tol, "frame %d differs by %.3g" % (k, e)
            assert abs(float(reward) - float(t["rewards"][k])) < 2e-5
        else:
            e_late = max(e_late, e)
            assert e <= late_tol, "frame %d differs by %.3g" % (k, e)
        want_done = bool(t["terminated"][k] or t["truncated"][k])
        if done or want_done:
            assert k >= n - 2, "episode ended at step %d, golden at %d" % (k, n - 1)
            if done and want_done:
                assert trunc == bool(t["truncated"][k])
            break
    return e_early, e_late


@pytest.mark.parametrize("name", ["random0", "random1", "random2", "random3", "gentle0", "gentle1"])
def test_fp64_free_run_matches_golden(hostsim, golden, name):
    """Free-running FP64 kernel source vs the reference-Python-on-oracle trace, whole episode."""
    t = golden[name]
    env = hostsim.env(0)
    obs = env.reset(t["goal"])
    assert np.allclose(obs, t["reset_obs"], rtol=0, atol=1e-9)

    def step(a):
        obs, r, fl, _ = env.step(a)
        return obs[-1], r, bool(fl & 8), bool(fl & 16)

    check_free_run(step, t)


def test_fp32_free_run_divergence_is_bounded(hostsim, golden):
    """FP32 mode is a different trajectory of a sensitive system: report divergence, bound it loosely."""
    t = golden["gentle0"]
    env = hostsim.env(1)
    env.reset(t["goal"])
    errs = []
    for k, a in enumerate(t["actions"][:1000]):
        obs, r, fl, _ = env.step(a)
        errs.append(np.abs(obs[-1][:3] - t["frames"][k][:3]).max())
        if fl & 8:
            break
    errs = np.array(errs)
    assert errs[:100].max() < 0.05       # metres over the first 100 steps (3.3 s)
    assert errs.max() < 50.0             # metres over 1000 steps of flight covering ~9 km


def test_auto_reset_and_terminal_observation(hostsim, golden):
    t = golden["gentle1"]     # crashes after 663 steps; numerically benign (no late chaotic amplification)
    env = hostsim.env(0)
    env.reset(t["goal"])
    for k, a in enumerate(t["actions"]):
        prev = obs if k else None
        obs, r, fl, tobs = env.step(a, auto_reset=1, seed=123, env_id=5)
    assert fl & 8 and fl & 2 and fl & 4
    # terminal observation = old stack shifted + terminal frame; returned obs = ten copies of the reset frame
    assert np.allclose(tobs[-1], t["frames"][-1], rtol=1e-6, atol=1e-5)
    assert np.array_equal(tobs[:-1], prev[1:])
    assert np.all(obs == obs[0])
    assert obs[0, 2] == np.float32(1524.0)
    g = obs[0, 12:]
    d = np.hypot(g[0], g[1])
    assert 1000.0 <= d < 10000.0 and 1000.0 <= g[2] < 4000.0
    # and the env keeps flying from the snapshot
    obs2, r2, fl2, _ = env.step(np.zeros(4, np.float32), auto_reset=1, seed=123, env_id=5)
    assert not (fl2 & 8) and np.array_equal(obs2[:-1], obs[1:])
