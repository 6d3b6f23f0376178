"""Carry-over reset (f16_reset_carryover / auto_reset = 2): the second and later episodes of ONE reference env object.

JSBSimEnv.reset is run_ic() + propulsion/set-running (jsbsim_gym.py:305-306) and run_ic() re-initialises only the
kinematic state, so actuator positions, PID memories, stale air data and accelerations leak from one episode into
the next. The golden file was produced by the reference's own jsbsim_gym.py driving one env object through four
episodes (tools/make_golden.py --multi; FDM underneath = the oracle, see DESIGN.md "Oracle").

CPU tests run the kernel's per-env source built with g++ (tests/hostsim); the GPU tests call libf16b200.so.
Tolerances as everywhere else: FP64 <= 1e-9 relative per state field, FP32 <= 1e-3 (floors in conftest).
"""
import os

import numpy as np
import pytest

from conftest import ROOT, state_floors
from test_hostsim_parity import NF, rel_err


@pytest.fixture(scope="module")
def one_object():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_env_one_object_4_episodes.npz"))
    eps = []
    for ep in range(4):
        eps.append({k.split("/")[1]: z[k] for k in z.files if k.startswith("ep%d/" % ep)})
    return eps


def test_oracle_reproduces_the_reference_env_over_four_episodes_of_one_object(oracle, one_object):
    env = oracle.OracleEnv()
    for t in one_object:
        obs = env.reset(t["goal"])
        assert np.array_equal(obs, t["reset_obs"])
        assert np.array_equal(env.fdm.pack_state(), t["states"][0])
        for k, a in enumerate(t["actions"]):
            obs, r, term, trunc = env.step(a)
            assert np.array_equal(obs[-1], t["frames"][k]) and r == t["rewards"][k]
            assert term == t["terminated"][k] and trunc == t["truncated"][k]
        assert np.array_equal(env.fdm.pack_state(), t["states"][-1])


def test_the_previous_episode_really_leaks_into_the_next(one_object, hostsim, state_fields):
    """What the default (snapshot) reset deliberately leaves out: after a crash the roll-rate derivative, the
    aileron position and the control-law memories of the next episode differ from a fresh env's by orders of
    magnitude relative to their scale."""
    snap, _ = hostsim.snapshot(NF)
    floors = state_floors(state_fields)
    assert np.array_equal(one_object[0]["states"][0], snap) or rel_err(one_object[0]["states"][0], snap, floors).max() < 1e-12
    for t in one_object[1:]:
        d = rel_err(t["states"][0], snap, floors)
        kin = [i for i, n in enumerate(state_fields) if n.startswith(("Q", "RI", "VI_", "EPA", "WI"))]
        assert d[kin].max() < 1e-12          # run_ic() does re-initialise FGPropagate
        assert d.max() > 1.0, d.max()        # ... and nothing else


@pytest.mark.parametrize("mode,tol", [(0, 1e-9), (1, 1e-3)])
def test_kernel_source_carryover_reset_matches_one_env_object(hostsim, one_object, state_fields, mode, tol):
    """Teacher-forced: the end state of each recorded episode is lifted into the env, the env is reset in carry-over
    mode with the recorded last action, and the state after reset plus every 5th env-step of the next episode are
    compared with the record."""
    floors = state_floors(state_fields)
    env = hostsim.env(mode)
    t0 = one_object[0]
    env.reset(t0["goal"])
    worst_reset = 0.0
    for ep in range(1, 4):
        prev, t = one_object[ep - 1], one_object[ep]
        n_prev = len(prev["actions"])
        # bring the env to the end of the previous episode: last recorded step, teacher-forced
        env.set_state(prev["states"][n_prev - 1], current_step=n_prev - 1)
        env.step(prev["actions"][n_prev - 1])
        obs = env.reset_carryover(t["goal"], prev["actions"][n_prev - 1])
        assert np.array_equal(obs, t["reset_obs"])
        e = rel_err(env.get_state(NF), t["states"][0], floors)
        worst_reset = max(worst_reset, e.max())
        assert e.max() < tol, (ep, state_fields[int(e.argmax())], e.max())
        for k in list(range(0, 8)) + list(range(8, len(t["actions"]), 5)):
            env.set_state(t["states"][k], current_step=k)
            o, r, fl, _ = env.step(t["actions"][k])
            e = rel_err(env.get_state(NF), t["states"][k + 1], floors)
            assert e.max() < tol, (ep, k, state_fields[int(e.argmax())], e.max())
            if mode == 0:
                assert np.allclose(o[-1][:12], t["frames"][k][:12], rtol=0, atol=1e-6 * np.maximum(1.0, np.abs(t["frames"][k][:12])))
    assert worst_reset > 0.0 or mode == 0


def test_kernel_source_auto_reset_carryover_free_run(hostsim, oracle, state_fields):
    """Free-running with auto_reset = 2 against one oracle env object that is reset whenever it finishes (the goal
    of each new episode is read back from the reset observation, as a consumer would). Dives, so episodes end
    every ~250 steps; FP64 frames stay bit-identical or within 1e-6 and the state right after each reset within 1e-9."""
    floors = state_floors(state_fields)
    env = hostsim.env(0)
    ref = oracle.OracleEnv()
    g = oracle.sample_goal(9)
    env.reset(g)
    ref.reset(g)
    rng = np.random.default_rng(77)
    resets = 0
    for k in range(900):
        a = np.clip([0.3 * rng.standard_normal(), 0.9 + 0.05 * rng.standard_normal(), 0.1 * rng.standard_normal(), 0.8],
                    [-1, -1, -1, 0], [1, 1, 1, 1]).astype(np.float32)
        obs, r, fl, tobs = env.step(a, auto_reset=2, seed=3, env_id=11)
        o2, r2, term, trunc = ref.step(a)
        assert bool(fl & 8) == (term or trunc)
        assert abs(r - r2) < 1e-5
        if term or trunc:
            resets += 1
            assert np.allclose(tobs[-1][:12], o2[-1][:12], rtol=1e-6, atol=1e-6)
            o2 = ref.reset(obs[-1][12:15])
            assert np.array_equal(obs, o2)
            e = rel_err(env.get_state(NF), ref.fdm.pack_state(), floors)
            assert e.max() < 1e-9, (k, state_fields[int(e.argmax())], e.max())
        else:
            assert np.allclose(obs[-1][:12], o2[-1][:12], rtol=1e-6, atol=1e-6), k
    assert resets >= 2


def test_kernel_source_second_reset_without_a_step(hostsim, oracle, state_fields):
    """reset(); reset() on a fresh env object: gear still down, 1500 lb in the tanks, the FCS ticks two more frames."""
    floors = state_floors(state_fields)
    env = hostsim.env(0)
    ref = oracle.OracleEnv()
    g = oracle.sample_goal(1)
    env.reset(g)
    ref.reset(g)
    for _ in range(2):
        env.reset_carryover(g, np.zeros(4, np.float32))
        ref.reset(g)
        e = rel_err(env.get_state(NF), ref.fdm.pack_state(), floors)
        assert e.max() < 1e-9, (state_fields[int(e.argmax())], e.max())
    a = np.array([0.1, -0.2, 0.05, 0.7], np.float32)
    for k in range(5):                       # the first flight frame still has the 1500-lb tanks' CG
        env.step(a)
        ref.step(a)
        e = rel_err(env.get_state(NF), ref.fdm.pack_state(), floors)
        assert e.max() < 1e-9, (k, state_fields[int(e.argmax())], e.max())


# ------------------------------------------------------------------------------------------------ B200
@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [("fp64", 1e-9), ("fp32", 1e-3)])
def test_gpu_carryover_reset_matches_one_env_object(one_object, state_fields, mode, tol):
    """The same teacher-forced check through the C ABI: one env per recorded boundary between two episodes."""
    import torch
    from f16_jsb_b200 import F16BatchedEnv, _lib
    floors = state_floors(state_fields)
    n = 3
    env = F16BatchedEnv(n, mode=mode, reset_mode="carryover")
    assert env.ground_reactions
    goals0 = np.stack([one_object[0]["goal"]] * n)
    env.reset(goals=torch.from_numpy(goals0).cuda())            # never reset before: canonical bring-up
    st, _ = env.snapshot()
    assert rel_err(env.get_state(0), st, floors).max() < (1e-12 if mode == "fp64" else 1e-6)
    last = np.zeros((n, 4), np.float32)
    for i in range(n):
        prev = one_object[i]
        m = len(prev["actions"])
        env.set_state(i, prev["states"][m - 1], current_step=m - 1)
        last[i] = prev["actions"][m - 1]
    obs, rew, done, trunc = env.step(torch.from_numpy(last).cuda(), auto_reset=False)
    assert done.cpu().numpy().all()
    goals = np.stack([one_object[i + 1]["goal"] for i in range(n)])
    obs = env.reset(goals=torch.from_numpy(goals).cuda())
    for i in range(n):
        t = one_object[i + 1]
        assert np.allclose(obs[i].cpu().numpy(), t["reset_obs"], rtol=0, atol=1e-9)   # device libm / FMA: alpha -2e-17 for 0
        e = rel_err(env.get_state(i), t["states"][0], floors)
        assert e.max() < tol, (i, state_fields[int(e.argmax())], e.max())
    # first env-steps of the new episodes, free-running from the reset state
    for k in range(6):
        a = np.stack([one_object[i + 1]["actions"][k] for i in range(n)])
        env.step(torch.from_numpy(a).cuda(), auto_reset=False)
        for i in range(n):
            e = rel_err(env.get_state(i), one_object[i + 1]["states"][k + 1], floors)
            assert e.max() < tol * (1 if mode == "fp64" else 3), (i, k, state_fields[int(e.argmax())], e.max())
    # auto_reset = 2 is refused without ground reactions
    env2 = F16BatchedEnv(32, mode="fp32")
    env2.reset()
    with pytest.raises(_lib.F16Error):
        _lib.check(env2.lib.f16_step(env2._h, None, 2, env2._stream()), "f16_step")


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["stacked", "frame"])
def test_gpu_auto_reset_carryover_free_run_vs_one_oracle_object_per_env(oracle, state_fields, layout):
    """64 envs dive with auto_reset = 2 for 700 steps; each has its own oracle env object that is reset (with the
    goal read back from the device) whenever it finishes. FP64: frames within 1e-6; the state after every reset
    within 1e-6 as well (free-running: what leaks is the end state of a crash, where round-off differences of the
    device build - FMA contraction, libm - have been amplified to ~1e-8; the teacher-forced test above holds 1e-9)."""
    import torch
    from f16_jsb_b200 import F16BatchedEnv
    floors = state_floors(state_fields)
    n, steps = 64, 700
    env = F16BatchedEnv(n, mode="fp64", reset_mode="carryover", obs_layout=layout, seed=5)
    goals = np.stack([oracle.sample_goal(100 + i) for i in range(n)])
    env.reset(goals=torch.from_numpy(goals).cuda())
    refs = [oracle.OracleEnv() for _ in range(n)]
    for i in range(n):
        refs[i].reset(goals[i])
    rng = np.random.default_rng(8)
    resets = 0
    for k in range(steps):
        a = np.stack([0.3 * rng.standard_normal(n), 0.9 + 0.05 * rng.standard_normal(n), 0.1 * rng.standard_normal(n),
                      np.full(n, 0.8)], axis=1)
        a = np.clip(a, [-1, -1, -1, 0], [1, 1, 1, 1]).astype(np.float32)
        obs, rew, done, trunc = env.step(torch.from_numpy(a).cuda(), auto_reset=True)
        newest = (obs if layout == "frame" else obs[:, -1, :]).cpu().numpy()
        d = done.cpu().numpy().astype(bool)
        r = rew.cpu().numpy()
        for i in range(n):
            o2, r2, term, trunc2 = refs[i].step(a[i])
            assert d[i] == (term or trunc2), (k, i)
            assert abs(r[i] - r2) < 2e-5, (k, i)
            if d[i]:
                resets += 1
                o2 = refs[i].reset(newest[i, 12:15])
                assert np.allclose(newest[i], o2[-1], rtol=0, atol=1e-9), (k, i)
                e = rel_err(env.get_state(i), refs[i].fdm.pack_state(), floors)
                assert e.max() < 1e-6, (k, i, state_fields[int(e.argmax())], e.max())
            else:
                assert np.allclose(newest[i, :12], o2[-1][:12], rtol=1e-6, atol=1e-6), (k, i)
    assert resets >= n


@pytest.mark.gpu
@pytest.mark.parametrize("host_obs", ["window", "copy"])
def test_gpu_vecenv_carryover_through_host_windows(oracle, state_fields, host_obs):
    """F16VecEnv(reset_mode="carryover") - NumPy in / NumPy out through the host-resident windows - against one
    oracle env object per env: terminal observations, reset observations and everything after them."""
    from f16_jsb_b200 import F16VecEnv
    n, steps = 96, 420
    venv = F16VecEnv(n, mode="fp64", reset_mode="carryover", seed=2, host_obs=host_obs)
    venv.seed(500)
    obs = venv.reset()
    refs = [oracle.OracleEnv() for _ in range(n)]
    for i in range(n):
        o = refs[i].reset(oracle.sample_goal(500 + i))
        assert np.allclose(o, obs[i], rtol=0, atol=1e-9)
    rng = np.random.default_rng(9)
    finished = 0
    for k in range(steps):
        a = np.stack([0.3 * rng.standard_normal(n), 0.9 + 0.05 * rng.standard_normal(n), 0.1 * rng.standard_normal(n),
                      np.full(n, 0.8)], axis=1)
        a = np.clip(a, [-1, -1, -1, 0], [1, 1, 1, 1]).astype(np.float32)
        obs, rew, dones, infos = venv.step(a)
        for i in range(n):
            o2, r2, term, trunc = refs[i].step(a[i])
            assert dones[i] == (term or trunc), (k, i)
            assert abs(rew[i] - r2) < 2e-5
            if dones[i]:
                finished += 1
                assert np.allclose(infos[i]["terminal_observation"], o2, rtol=1e-6, atol=1e-6), (k, i)
                o2 = refs[i].reset(obs[i, -1, 12:15])
                assert np.allclose(obs[i], o2, rtol=0, atol=1e-9), (k, i)
            else:
                assert np.allclose(obs[i], o2, rtol=1e-6, atol=1e-6), (k, i)
    assert finished >= n
    venv.close()
