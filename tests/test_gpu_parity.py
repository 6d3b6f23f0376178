"""Parity tests proper: the CUDA library (through its C ABI, via the ctypes host layer) against the
golden vectors and against the oracle on the same seeded inputs. Run on the B200 box with -m gpu.

Stated tolerances (DESIGN.md "Parity"):
  FP64 mode: teacher-forced one env-step <= 1e-9 relative per state field (north_star contract 1e-6);
             free-run frames <= 1e-5 for the first 300 steps, <= 1e-1 afterwards (open-loop unstable airframe:
             measured 2e-2 on the last frame before a crash, median 0 = bit-identical frames).
  FP32 mode: teacher-forced one env-step <= 1e-3 relative per state field (floors in conftest);
             free-run position divergence < 0.05 m after 100 steps, < 50 m after 1000 steps.
"""
import numpy as np
import pytest
import torch

from conftest import state_floors
from test_hostsim_parity import check_free_run, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F16BatchedEnv():
    from f16_jsb_b200 import F16BatchedEnv
    return F16BatchedEnv


def test_native_library_is_loaded_and_kernels_launch(F16BatchedEnv):
    env = F16BatchedEnv(64, mode="fp32")
    n0 = env.launch_count()
    env.reset()
    env.step(None)
    torch.cuda.synchronize()
    assert env.launch_count() >= n0 + 2
    maps = open("/proc/self/maps").read()
    assert "libf16b200.so" in maps


def test_snapshot_matches_oracle(F16BatchedEnv, oracle, state_fields):
    env = F16BatchedEnv(4, mode="fp64")
    st, props = env.snapshot()
    o = oracle.OracleEnv()
    o.reset(oracle.sample_goal(0))
    want = o.fdm.pack_state()
    e = rel_err(st, want, state_floors(state_fields))
    # the parity kernel's double math is within a few ulp of IEEE (no slow paths, exp(y log x) for pow): measured 3e-11
    assert e.max() < 1e-9, (state_fields[int(e.argmax())], float(e.max()))
    from f16_jsb_b200.constants import STATE_FORMAT
    for i, p in enumerate(STATE_FORMAT):
        assert props[i] == pytest.approx(o.fdm[p], rel=1e-10, abs=1e-10)


@pytest.mark.parametrize("mode,tol", [("fp64", 1e-9), ("fp32", 1e-3)])
def test_teacher_forced_step_parity_batched(F16BatchedEnv, golden, state_fields, mode, tol):
    """Every golden state of both recorded episodes is lifted into its own env of one batch, one
    env-step is taken with the recorded action and the result is compared with the next state."""
    floors = state_floors(state_fields)
    for name in ("random0", "gentle0"):
        t = golden[name]
        n = len(t["actions"])
        env = F16BatchedEnv(n, mode=mode)
        goals = torch.from_numpy(np.tile(t["goal"], (n, 1))).cuda()
        env.reset(goals=goals)
        env.unpack_states(torch.from_numpy(t["states"][:n]).cuda())
        # current_step selects the first-flight-frame mass set (step 0) and drives truncation
        steps = np.arange(n, dtype=np.int32)
        _set_all_steps(env, steps)
        obs, rew, done, trunc = env.step(torch.from_numpy(t["actions"][:n]).cuda(), auto_reset=False)
        got = env.pack_states().cpu().numpy()
        e = rel_err(got, t["states"][1:n + 1], floors[None, :])
        worst = np.unravel_index(int(e.argmax()), e.shape)
        assert e.max() < tol, (name, int(worst[0]), state_fields[int(worst[1])], float(e.max()))
        if mode == "fp64":
            fr = obs[:, -1, :12].cpu().numpy()
            assert np.allclose(fr, t["frames"][:n, :12], rtol=1e-6, atol=1e-6)


def _set_all_steps(env, steps):
    """Write current_step for every env (E field EF_STEP) through the public single-env call."""
    from f16_jsb_b200 import _lib
    for k, s in enumerate(steps.tolist()):
        _lib.check(env.lib.f16_set_env_step(env._h, k, int(s)), "f16_set_env_step")


@pytest.mark.parametrize("name", ["random0", "random1", "random2", "random3", "gentle0", "gentle1"])
def test_fp64_free_run_matches_golden(F16BatchedEnv, golden, name):
    t = golden[name]
    env = F16BatchedEnv(1, mode="fp64")
    obs = env.reset(goals=torch.from_numpy(t["goal"].reshape(1, 3)).cuda())
    assert np.allclose(obs[0].cpu().numpy(), t["reset_obs"], rtol=0, atol=1e-9)
    act = torch.zeros((1, 4), dtype=torch.float32, device="cuda")
    state = {"prev": obs[0].cpu().numpy().copy(), "k": 0}

    def step(a):
        act.copy_(torch.from_numpy(a.reshape(1, 4)))
        o, r, d, tr = env.step(act, auto_reset=False)
        o = o[0].cpu().numpy()
        assert np.array_equal(o[:-1], state["prev"][1:]), "stack did not shift by one row"
        state["prev"] = o.copy()
        return o[-1], float(r[0].item()), bool(d[0].item()), bool(tr[0].item())

    check_free_run(step, t)


def test_fp64_batch_of_4096_vs_oracle(F16BatchedEnv, oracle):
    """BASELINE config 2: 4096 envs, FP64 parity mode, seeds 0..4095 for the goals, host-generated
    random actions; 120 steps against the oracle's batch trajectories (frames, rewards, flags)."""
    n, steps = 4096, 120
    goals = np.stack([oracle.sample_goal(s) for s in range(n)])
    rng = np.random.default_rng(42)
    actions = rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1], size=(steps, n, 4)).astype(np.float32)
    frames, rewards, flags = oracle.batch_trajectory(goals, actions)
    env = F16BatchedEnv(n, mode="fp64")
    env.reset(goals=torch.from_numpy(goals).cuda())
    a_dev = torch.from_numpy(actions).cuda()
    alive = np.ones(n, bool)
    worst = 0.0
    for k in range(steps):
        obs, rew, done, trunc = env.step(a_dev[k], auto_reset=False)
        fr = obs[:, -1, :].cpu().numpy()
        d = done.cpu().numpy().astype(bool)
        e = np.abs(fr[alive, :12] - frames[k][alive, :12]) / np.maximum(np.abs(frames[k][alive, :12]), 1e-2)
        worst = max(worst, float(e.max()))
        assert e.max() < 1e-5, (k, float(e.max()))
        assert np.array_equal(fr[alive, 12:], goals[alive])
        assert np.abs(rew.cpu().numpy()[alive] - rewards[k][alive]).max() < 2e-5
        assert np.array_equal(d[alive], (flags[k][alive] & 3) != 0)
        alive &= ~d
    assert alive.sum() > n // 2


def test_fp32_teacher_forced_population_and_free_run(F16BatchedEnv, golden):
    t = golden["gentle0"]
    env = F16BatchedEnv(1, mode="fp32")
    env.reset(goals=torch.from_numpy(t["goal"].reshape(1, 3)).cuda())
    act = torch.zeros((1, 4), dtype=torch.float32, device="cuda")
    errs = []
    for k in range(1000):
        act.copy_(torch.from_numpy(t["actions"][k].reshape(1, 4)))
        o, r, d, tr = env.step(act, auto_reset=False)
        errs.append(float(np.abs(o[0, -1, :3].cpu().numpy() - t["frames"][k][:3]).max()))
        if bool(d[0].item()):
            break
    errs = np.array(errs)
    assert errs[:100].max() < 0.05 and errs.max() < 50.0


def test_batch_invariance_and_env_independence(F16BatchedEnv, golden):
    """Same goal + same actions in every env of a ragged batch (not a multiple of 32 or of the block
    size) -> bit-identical trajectories in every env, identical to a batch of 1."""
    t = golden["random1"]
    n = 1000 + 37
    env = F16BatchedEnv(n, mode="fp32")
    one = F16BatchedEnv(1, mode="fp32")
    g = torch.from_numpy(t["goal"].reshape(1, 3)).cuda()
    env.reset(goals=g.repeat(n, 1))
    one.reset(goals=g)
    for k in range(40):
        a = torch.from_numpy(t["actions"][k].reshape(1, 4)).cuda()
        o, r, d, tr = env.step(a.repeat(n, 1), auto_reset=False)
        o1, r1, d1, tr1 = one.step(a, auto_reset=False)
        assert torch.equal(o, o1.expand_as(o)) and torch.equal(r, r1.expand_as(r))
    assert torch.equal(env.pack_states(), one.pack_states().expand(n, -1))


def test_auto_reset_terminal_obs_and_stats(F16BatchedEnv, golden):
    t = golden["gentle1"]            # crashes after 663 steps; numerically benign episode
    n = 96
    env = F16BatchedEnv(n, mode="fp64", seed=7)
    env.reset(goals=torch.from_numpy(np.tile(t["goal"], (n, 1))).cuda())
    a_dev = torch.from_numpy(t["actions"]).cuda()
    prev = None
    for k in range(len(t["actions"])):
        prev = env.obs.clone()
        obs, rew, done, trunc = env.step(a_dev[k].reshape(1, 4).repeat(n, 1), auto_reset=True)
    assert bool(done.all()) and not bool(trunc.any())
    tob = env.terminal_obs.cpu().numpy()
    assert np.allclose(tob[:, -1, :12], t["frames"][-1][None, :12], rtol=1e-5, atol=1e-4)
    assert np.array_equal(tob[:, :-1], prev[:, 1:].cpu().numpy())
    o = obs.cpu().numpy()
    assert np.all(o == o[:, :1, :]) and np.all(o[:, 0, 2] == np.float32(1524.0))
    g = o[:, 0, 12:]
    d = np.hypot(g[:, 0], g[:, 1])
    assert np.all((d >= 1000) & (d < 10000) & (g[:, 2] >= 1000) & (g[:, 2] < 4000))
    assert len({tuple(x) for x in g.tolist()}) == n          # every env drew its own goal
    assert np.allclose(env.ep_len.cpu().numpy(), len(t["actions"]))
    assert np.allclose(env.ep_return.cpu().numpy(), float(t["rewards"].sum()), atol=1e-2)
    st = env.stats()
    assert st["episodes"] == n and st["crashes"] == n and st["goals"] == 0 and st["truncations"] == 0
    assert st["length_sum"] == n * len(t["actions"]) and st["env_steps"] == n * len(t["actions"])
    # next step continues from the reset state with an intact stack
    obs2, _, done2, _ = env.step(torch.zeros((n, 4), device="cuda"), auto_reset=True)
    assert not bool(done2.any()) and torch.equal(obs2[:, :-1], torch.from_numpy(o).cuda()[:, 1:])


def test_masked_reset_leaves_other_envs_untouched(F16BatchedEnv):
    env = F16BatchedEnv(70, mode="fp32", seed=3)
    env.reset()
    for _ in range(5):
        env.step(None, auto_reset=False)
    before = env.pack_states().clone()
    obs_before = env.obs.clone()
    mask = torch.zeros(70, dtype=torch.uint8, device="cuda")
    mask[::7] = 1
    env.reset(mask=mask)
    after = env.pack_states()
    snap = torch.from_numpy(env.snapshot()[0]).cuda()
    m = mask.bool()
    assert torch.equal(after[~m], before[~m]) and torch.equal(env.obs[~m], obs_before[~m])
    assert torch.allclose(after[m], snap.expand(int(m.sum()), -1).to(after.dtype), rtol=1e-6, atol=1e-6)


def test_random_action_rollout_full_size_properties(F16BatchedEnv):
    """BASELINE config 3 size (65 536 envs, FP32, in-kernel Philox actions, auto-reset): size-independent
    properties - quaternions stay normalised, observations stay finite and inside the declared bounds,
    every done env restarts from the snapshot, the statistics add up."""
    n = 65536
    env = F16BatchedEnv(n, mode="fp32", seed=11)
    env.reset()
    dones = 0
    for k in range(300):
        obs, rew, done, trunc = env.step(None, auto_reset=True)
        dones += int(done.sum().item())
    st = env.pack_states()
    qn = st[:, 0:4].norm(dim=1)
    assert float((qn - 1).abs().max()) < 1e-9
    assert bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
    ang = obs[:, :, 9:12]
    assert float(ang.abs().max()) <= np.pi + 1e-5 and float(obs[:, :, 10].abs().max()) <= np.pi / 2 + 1e-5
    assert float(obs[:, :, 3].min()) >= 0.0
    s = env.stats()
    assert s["episodes"] == dones and s["env_steps"] == 300 * n
    assert s["crashes"] + s["goals"] + s["truncations"] == s["episodes"]


def test_results_do_not_depend_on_how_the_envs_are_sharded(F16BatchedEnv):
    """env ids are partitioned contiguously over the ranks and every random draw (goals on reset and
    auto-reset, in-kernel actions) is keyed by the GLOBAL env id, so one batch of N envs and two shards of it (env_id_base
    = 0 and N/2, as distributed.shard_range hands them out) produce bit-identical observations, rewards and flags."""
    from f16_jsb_b200 import _lib
    n, half, steps = 4096 + 64, 2048 + 32, 40
    whole = F16BatchedEnv(n, mode="fp32", seed=21)
    parts = [F16BatchedEnv(half, mode="fp32", seed=21, env_id_base=0), F16BatchedEnv(half, mode="fp32", seed=21, env_id_base=half)]
    for e in [whole] + parts:
        e.reset()
    # some envs are about to hit the time limit: their auto-reset goals must come out the same in either arrangement
    for i in range(0, n, 97):
        _lib.check(whole.lib.f16_set_env_step(whole._h, i, 1190 + i % 9), "f16_set_env_step")
        p, j = divmod(i, half)
        _lib.check(parts[p].lib.f16_set_env_step(parts[p]._h, j, 1190 + i % 9), "f16_set_env_step")
    finished = 0
    for k in range(steps):
        o, r, d, t = whole.step(None, auto_reset=True)
        outs = [e.step(None, auto_reset=True) for e in parts]
        assert torch.equal(o, torch.cat([x[0] for x in outs])), k
        assert torch.equal(r, torch.cat([x[1] for x in outs])) and torch.equal(d, torch.cat([x[2] for x in outs]))
        assert torch.equal(t, torch.cat([x[3] for x in outs]))
        finished += int(d.sum().item())
    assert finished >= n // 97
    assert torch.equal(whole.pack_states(), torch.cat([e.pack_states() for e in parts]))
    sw, sp = whole.stats(), [e.stats() for e in parts]
    for key in ("episodes", "truncations", "crashes", "length_sum", "env_steps"):
        assert sw[key] == sp[0][key] + sp[1][key], key


@pytest.mark.parametrize("mode", ["fp32", "fp64"])
def test_ring_layout_is_value_identical_to_stacked(F16BatchedEnv, mode):
    """obs_layout='ring' (frame written twice, zero-copy window view) must return the stacks, rewards,
    flags, terminal observations and host copies of the stacked (in-place shift) layout, across
    auto-resets and masked resets, for a ragged batch. The two layouts are separate instantiations of the
    step kernel and the compiler contracts a few multiply-adds differently, so values agree to float
    rounding (1e-5 relative here, over 70 free-running steps), not bit for bit; structure is exact."""
    def same(x, y, what=""):
        assert x.shape == y.shape
        assert torch.allclose(x, y, rtol=2e-5, atol=2e-4), what

    n = 1000 + 37
    a = F16BatchedEnv(n, mode=mode, seed=9, obs_layout="stacked")
    b = F16BatchedEnv(n, mode=mode, seed=9, obs_layout="ring")
    g = torch.Generator(device="cuda").manual_seed(1)
    # goals 150..900 m straight ahead at the start altitude: the envs reach them (+10, terminated) between
    # steps ~10 and ~100, so auto-resets and terminal observations occur all along the run
    goals0 = torch.zeros((n, 3), device="cuda")
    goals0[:, 0] = torch.rand(n, device="cuda", generator=g) * 750 + 150
    goals0[:, 2] = 1524.0
    assert torch.equal(a.reset(goals=goals0), b.reset(goals=goals0).contiguous())     # resets are copies of the snapshot: exact
    host = np.zeros((n, 10, 15), np.float32)
    for k in range(70):
        act = (torch.rand((n, 4), device="cuda", generator=g) * 2 - 1) * 0.1
        act[:, 3] = 0.8
        if k == 30:
            oa, ra, da, ta = a.step(act, auto_reset=True)
            b.step_host(act.cpu().numpy(), host, None, None, None, auto_reset=True)     # strided device->host copy
            same(torch.from_numpy(host).cuda(), oa, "host copy of the ring window")
            ob = b.obs
        else:
            oa, ra, da, ta = a.step(act, auto_reset=True)
            ob, rb, db, tb = b.step(act, auto_reset=True)
            same(ra, rb, "rewards")
            assert torch.equal(da, db) and torch.equal(ta, tb)
        assert ob.shape == (n, 10, 15) and ob.stride() == (15, n * 15, 1)     # slot-major ring: (N,15) planes, zero-copy window
        same(oa, ob.contiguous(), "window differs at step %d" % k)
        same(oa.reshape(n, 150), ob.reshape(n, 150))
        if bool(da.any()):
            m = da.bool()
            same(a.terminal_obs[m], b.terminal_obs[m], "terminal observations")
        if k == 20:
            mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
            mask[5::13] = 1
            gl = torch.rand((n, 3), device="cuda", generator=g) * 3000 + 1000
            a.reset(mask=mask, goals=gl)
            b.reset(mask=mask, goals=gl)
            same(a.obs, b.obs.contiguous())
    sa, sb = a.stats(), b.stats()
    assert sa["goals"] > 100 and sa["episodes"] == sb["episodes"] and sa["goals"] == sb["goals"] and sa["length_sum"] == sb["length_sum"]


@pytest.mark.parametrize("mode,tol", [("fp64", 1e-8), ("fp32", 1e-3)])
def test_ground_contact_step_parity(F16BatchedEnv, oracle, state_fields, mode, tol):
    """ground reactions (f16.xml:85-215). One batch holds every synthetic contact state of
    tests/test_ground_contact.py (each STRUCTURE contact, several at once, static and dynamic friction)
    followed by the crash steps of random-action episodes and by ordinary in-flight states, so lanes of
    one warp take the cold ground path and the hot path side by side; one teacher-forced env-step against
    the oracle. The layouts other than the default take the same path (checked in frame layout too)."""
    from test_ground_contact import CASES, crash_steps, flying_oracle_env, synthetic_state
    floors = state_floors(state_fields)
    act0 = np.array([0.2, -0.1, 0.1, 0.7], np.float32)
    s0, s1, acts, goals, steps, frames, touched = [], [], [], [], [], [], []
    for name, h, phi, th, psi, uvw, pqr, want in CASES:
        env, goal = flying_oracle_env(oracle)
        st = synthetic_state(env.fdm.pack_state(), state_fields, h, phi, th, psi, uvw, pqr)
        env.fdm.unpack_state(st)
        obs, r, term, trunc = env.step(act0)
        s0.append(st); s1.append(env.fdm.pack_state()); acts.append(act0); goals.append(goal); steps.append(5)
        frames.append(obs[-1]); touched.append(True)
        # an ordinary state in flight right beside it
        env2, goal2 = flying_oracle_env(oracle)
        st2 = env2.fdm.pack_state()
        obs2, _, _, _ = env2.step(act0)
        s0.append(st2); s1.append(env2.fdm.pack_state()); acts.append(act0); goals.append(goal2); steps.append(3)
        frames.append(obs2[-1]); touched.append(False)
    for goal, st, t, a, st1, frame, reward, wow in crash_steps(oracle, n_episodes=24):
        s0.append(st); s1.append(st1); acts.append(a); goals.append(goal); steps.append(t); frames.append(frame); touched.append(any(wow))
    n = len(s0)
    for layout in ("stacked", "frame"):
        env = F16BatchedEnv(n, mode=mode, obs_layout=layout, ground_reactions=True)   # the fp32 default is off
        assert env.ground_reactions
        env.reset(goals=torch.from_numpy(np.stack(goals)).cuda())
        env.unpack_states(torch.from_numpy(np.stack(s0)).cuda())
        _set_all_steps(env, np.array(steps, dtype=np.int32))
        out = env.step(torch.from_numpy(np.stack(acts)).cuda(), auto_reset=False)
        got = env.pack_states().cpu().numpy()
        e = rel_err(got, np.stack(s1), floors[None, :])
        if mode == "fp32":
            # stated tolerance of the float mode for a STANDING aircraft: calibrated airspeed within 0.5 kt absolute
            # (float32 cannot resolve pow(1 + 1e-7, 1/3.5) - 1; the value only feeds FCS thresholds at 5 kt and above)
            i_case, i_vc = 2 * [c[0] for c in CASES].index("at_rest_static_friction"), state_fields.index("VC_KTS")
            assert abs(got[i_case, i_vc] - np.stack(s1)[i_case, i_vc]) <= 0.5
            e[i_case, i_vc] = min(e[i_case, i_vc], tol / 2)
        worst = np.unravel_index(int(e.argmax()), e.shape)
        assert e.max() < tol, (layout, int(worst[0]), state_fields[int(worst[1])], float(e.max()))
        if mode == "fp64" and layout == "stacked":
            fr = out[0][:, -1, :12].cpu().numpy()
            assert np.allclose(fr, np.stack(frames)[:, :12], rtol=1e-6, atol=1e-6)
    assert sum(touched) >= len(CASES) + 1


def test_ground_reactions_default_and_switch(F16BatchedEnv):
    """On by default in FP64 (parity) mode, off in FP32 (throughput) mode; with them off the step is the
    ground-less instantiation: a state at the surface then keeps its aerodynamic accelerations."""
    assert F16BatchedEnv(32, mode="fp64").ground_reactions and not F16BatchedEnv(32, mode="fp32").ground_reactions
    from test_ground_contact import synthetic_state
    import re, os
    txt = open(os.path.join(os.path.dirname(__file__), "..", "include", "f16_state_fields.h")).read()
    body = txt.split("enum f16_state_field")[1].split("F16_NUM_STATE_FIELDS")[0]
    fields = []
    for m in re.finditer(r"F16S_([A-Z0-9_]+)", body):
        if m.group(1) not in fields:
            fields.append(m.group(1))
    out = {}
    for on in (False, True):
        env = F16BatchedEnv(32, mode="fp64", ground_reactions=on)
        env.reset()
        st = synthetic_state(env.snapshot()[0], fields, 2.0, 0.0, 0.0, 0.3, (150.0, 0.0, 8.0))
        env.unpack_states(torch.from_numpy(np.tile(st, (32, 1))).cuda())
        env.step(torch.zeros((32, 4), device="cuda"), auto_reset=False)
        out[on] = env.pack_states().cpu().numpy()[0]
    i = fields.index("WDOT_X")
    assert np.abs(out[True][i:i + 3] - out[False][i:i + 3]).max() > 1e-2


@pytest.mark.parametrize("mode,tol", [("fp64", 1e-9), ("fp32", 1e-3)])
def test_envelope_step_parity(F16BatchedEnv, oracle, state_fields, mode, tol):
    """tests/test_envelope_parity.py on the CUDA library: 800 random states over the whole flight envelope
    (Mach 0.14-2.04, up to 60 000 ft, alpha -30..95 deg, lit and unlit afterburner, saturated actuators) in one
    batch, one teacher-forced env-step against the oracle."""
    from test_envelope_parity import envelope_states, oracle_step_all
    from test_ground_contact import flying_oracle_env
    floors = state_floors(state_fields)
    env0, _ = flying_oracle_env(oracle)
    base = env0.fdm.pack_state()
    sa, aa = envelope_states(base, state_fields, 400, seed=11, region="flown")
    sb, ab = envelope_states(base, state_fields, 400, seed=12, region="wide")
    states, acts = np.concatenate([sa, sb]), np.concatenate([aa, ab])
    goal, want, frames = oracle_step_all(oracle, states, acts)
    n = len(states)
    env = F16BatchedEnv(n, mode=mode)
    env.reset(goals=torch.from_numpy(np.tile(goal, (n, 1))).cuda())
    env.unpack_states(torch.from_numpy(states).cuda())
    _set_all_steps(env, np.full(n, 7, dtype=np.int32))
    obs, rew, done, trunc = env.step(torch.from_numpy(acts).cuda(), auto_reset=False)
    e = rel_err(env.pack_states().cpu().numpy(), want, floors[None, :])
    # (the supersonic calibrated airspeed is inside the common tolerance in both modes: csrc/f16_model.cuh, FGAuxiliary block)
    worst = np.unravel_index(int(e.argmax()), e.shape)
    assert e.max() < tol, (int(worst[0]), state_fields[int(worst[1])], float(e.max()))
    if mode == "fp64":
        assert np.allclose(obs[:, -1, :12].cpu().numpy(), frames[:, :12], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("mode", ["fp32", "fp64"])
def test_near_ground_tiles_first_steps_every_env_exactly_once(F16BatchedEnv, mode):
    """The ground-reaction builds start a whole-batch launch with early CTAs for the tiles whose envs may touch the ground
    (csrc/f16_b200.cu, StepArgs::hot_*) and let the regular CTA of such a tile skip it. Whatever the list holds, every env
    must be stepped exactly once per step: a batch stepped whole (scheduling on) against the same batch stepped in three
    ranges (f16_step_range: plain tile order), frame layout, ragged size, most envs diving into the ground with
    auto-reset for 330 steps - frames, rewards and flags bit-identical at every step, and the list really was in use."""
    import ctypes as C
    from f16_jsb_b200 import _lib
    n = 3000 + 21
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.zeros((n, 4), device="cuda")
    a[:, 1] = 0.9 * (torch.rand(n, device="cuda", generator=g) > 0.3).float()
    a[:, 0] = torch.rand(n, device="cuda", generator=g) - 0.5
    a[:, 3] = 1.0
    whole = F16BatchedEnv(n, mode=mode, seed=8, obs_layout="frame", ground_reactions=True)
    parts = F16BatchedEnv(n, mode=mode, seed=8, obs_layout="frame", ground_reactions=True)
    assert torch.equal(whole.reset(), parts.reset())
    cuts = [0, 1024, 2048, n]
    redone = crashes = 0
    for k in range(330):
        whole.step(a, auto_reset=True)
        _lib.check(parts.lib.f16_step_begin(parts._h, parts._stream()), "f16_step_begin")
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            _lib.check(parts.lib.f16_step_range(parts._h, C.c_void_p(a.data_ptr()), 1, lo, hi - lo, parts._stream()), "f16_step_range")
        assert torch.equal(whole.obs, parts.obs), k
        assert torch.equal(whole.reward, parts.reward) and torch.equal(whole.done, parts.done) and torch.equal(whole.truncated, parts.truncated), k
    sw, sp = whole.stats(), parts.stats()
    assert sw["crashes"] == sp["crashes"] > 100 and sw["ground_redos"] == sp["ground_redos"] > 5, (sw, sp)
    assert torch.equal(whole.pack_states(), parts.pack_states())
