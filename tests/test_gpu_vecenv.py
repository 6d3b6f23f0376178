"""SB3 VecEnv / Gymnasium surface on the GPU: return types, shapes, auto-reset infos, seeding."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


HOST_OBS = [("window", 2), ("window", 1), ("copy", 2)]


def _reset_with_near_goals(venv, n, seed):
    """Goals 150..900 m straight ahead at the start altitude: reached (+10, terminated) between steps ~10 and
    ~100, so terminal observations and auto-resets occur all along a short run."""
    g = np.zeros((n, 3), np.float32)
    g[:, 0] = np.random.default_rng(seed).uniform(150, 900, size=n)
    g[:, 2] = 1524.0
    venv.env.reset(goals=torch.from_numpy(g).cuda())
    if venv._win is not None:
        return venv._win.reset(venv.env, venv.env._stream()).obs
    return venv.env.obs.cpu().numpy()


@pytest.mark.parametrize("host_obs,rings", HOST_OBS)
def test_vecenv_api_contract(golden, host_obs, rings):
    from f16_jsb_b200 import F16VecEnv
    n = 8
    env = F16VecEnv(n, mode="fp64", lazy_infos=False, host_obs=host_obs, host_rings=rings)
    assert env.num_envs == n and env.observation_space.shape == (10, 15) and env.action_space.shape == (4,)
    assert env.get_attr("render_mode") == [None] * n and env.env_is_wrapped(object) == [False] * n
    env.seed(0)
    obs = env.reset()
    assert obs.shape == (n, 10, 15) and obs.dtype == np.float32
    # seed(s) -> env i is reset with seed s+i, goal = default_rng(s+i) draw (jsbsim_gym.py:312-323)
    assert np.array_equal(obs[0, 0, 12:], golden["random0"]["goal"]) and np.array_equal(obs[1, 0, 12:], golden["random1"]["goal"])
    t = golden["random0"]
    a = np.tile(t["actions"][0], (n, 1))
    obs2, rew, dones, infos = env.step(a)
    assert obs2.shape == (n, 10, 15) and rew.shape == (n,) and rew.dtype == np.float32 and dones.dtype == bool
    assert len(infos) == n and infos[0]["TimeLimit.truncated"] is False
    assert np.allclose(obs2[0, -1], t["frames"][0], rtol=1e-6, atol=1e-6) and abs(rew[0] - t["rewards"][0]) < 2e-5
    if rings == 2:
        assert obs is not obs2 and np.array_equal(obs[0, 1:], obs2[0, :-1])      # `_last_obs` stays valid across the next step
        assert np.all(obs[0] == obs[0, 0])
    env.close()


@pytest.mark.parametrize("host_obs,rings", HOST_OBS)
def test_vecenv_done_infos_match_dummy_vec_env_conventions(golden, host_obs, rings):
    from f16_jsb_b200 import F16VecEnv
    t = golden["gentle1"]
    n = 70
    env = F16VecEnv(n, mode="fp64", seed=5, host_obs=host_obs, host_rings=rings)
    env.seed(11)         # env 0 gets the golden goal of seed 11 (trace gentle1)
    env.reset()
    for k in range(len(t["actions"])):
        obs, rew, dones, infos = env.step(np.tile(t["actions"][k], (n, 1)))
    assert dones[0] and infos[0]["TimeLimit.truncated"] is False
    assert infos[0]["terminal_observation"].shape == (10, 15)
    assert np.allclose(infos[0]["terminal_observation"][-1][:12], t["frames"][-1][:12], rtol=1e-5, atol=1e-4)
    assert infos[0]["episode"]["l"] == len(t["actions"]) and abs(infos[0]["episode"]["r"] - float(t["rewards"].sum())) < 1e-2
    assert np.all(obs[0] == obs[0, 0])                       # returned obs is the reset observation
    assert sum(1 for i in infos if i.get("episode")) == int(dones.sum())
    env.close()


@pytest.mark.parametrize("rings,dma_both,zero_copy", [(1, False, 1), (2, False, 1), (2, True, 1), (1, False, 0), (2, False, 0)])
def test_window_mode_returns_what_copy_mode_returns(rings, dma_both, zero_copy, monkeypatch):
    """Host-resident windows (60 B per env-step over PCIe) against the device-side stacks copied out whole
    (600 B): identical observations, rewards, flags, terminal observations and episode statistics across
    crashes and auto-resets. FP64 mode: there the two layouts' kernel instantiations agree bit for bit (the
    float instantiations contract a few multiply-adds differently, see test_ring_layout_is_value_identical_to_stacked)."""
    from f16_jsb_b200 import F16VecEnv
    n, steps = 2048, 120
    # zero_copy 1: the kernel stores its frames straight into the mapped ring (default up to 32 768 envs); 0: copy engine
    monkeypatch.setenv("F16_HOSTWIN_ZEROCOPY", str(zero_copy))
    a_env = F16VecEnv(n, mode="fp64", seed=3, host_obs="window", host_rings=rings, host_dma_both=dma_both)
    b_env = F16VecEnv(n, mode="fp64", seed=3, host_obs="copy")
    oa, ob = _reset_with_near_goals(a_env, n, 5), _reset_with_near_goals(b_env, n, 5)
    assert np.array_equal(oa, ob)
    rng = np.random.default_rng(0)
    finished = 0
    prev = None
    for k in range(steps):
        act = (rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1], size=(n, 4)) * 0.1).astype(np.float32)
        act[:, 3] = 0.8
        buf = a_env.action_buffer() if k % 2 else act          # pinned staging and plain arrays both work
        if k % 2:
            buf[...] = act
        oa, ra, da, ia = a_env.step(buf)
        ob, rb, db, ib = b_env.step(act)
        assert np.array_equal(da, db) and np.array_equal(ra, rb), k
        assert np.array_equal(oa, ob), k
        if rings == 2 and prev is not None:
            assert np.array_equal(prev[0], prev[1]), k          # last step's array is still intact
        prev = (oa, ob.copy())
        for i in np.flatnonzero(da):
            assert np.array_equal(ia[i]["terminal_observation"], ib[i]["terminal_observation"]), (k, i)
            assert ia[i]["TimeLimit.truncated"] == ib[i]["TimeLimit.truncated"]
            assert ia[i]["episode"]["l"] == ib[i]["episode"]["l"] and ia[i]["episode"]["r"] == ib[i]["episode"]["r"]
            finished += 1
        assert ia[int(np.flatnonzero(~da)[0])] == {"TimeLimit.truncated": False}
    assert finished > 100
    a_env.close()
    b_env.close()


@pytest.mark.parametrize("chunks", [1, 3, 8])
def test_window_mode_fp32_is_the_deque_of_the_frame_layout(chunks, monkeypatch):
    """FP32 mode, 150 steps over 4 096 envs that reach their goals all along the run: F16VecEnv's host windows against a
    deque model (jsbsim_gym.py:150,235,325-329; dummy_vec_env.py:63-72) fed with the device outputs of a second
    env in the same frame layout - the same kernel instantiation, so everything is bit-exact."""
    from collections import deque

    from f16_jsb_b200 import F16BatchedEnv, F16VecEnv
    monkeypatch.setenv("F16_HOSTWIN_CHUNKS", str(chunks))     # pieces one step is pipelined in (default: by batch size)
    n, steps = 4096 - 37 * (chunks == 3), 150                  # a ragged batch for the odd split
    venv = F16VecEnv(n, mode="fp32", seed=7, host_obs="window", host_rings=2)
    ref = F16BatchedEnv(n, mode="fp32", seed=7, obs_layout="frame", done_list=True)
    g = np.zeros((n, 3), np.float32)
    g[:, 0] = np.random.default_rng(5).uniform(150, 900, size=n)
    g[:, 2] = 1524.0
    obs = _reset_with_near_goals(venv, n, 5)
    f0 = ref.reset(goals=torch.from_numpy(g).cuda()).cpu().numpy()
    stacks = [deque([f0[i]] * 10, maxlen=10) for i in range(n)]
    assert np.array_equal(obs, np.array([np.array(d) for d in stacks]))
    rng = np.random.default_rng(1)
    finished = 0
    for k in range(steps):
        act = (rng.uniform([-1, -1, -1, 0], [1, 1, 1, 1], size=(n, 4)) * 0.1).astype(np.float32)
        act[:, 3] = 0.8
        obs, rew, dones, infos = venv.step(act)
        fr, rw, dn, tr = ref.step(torch.from_numpy(act).cuda(), auto_reset=True)
        fr, rw, dn, tr = fr.cpu().numpy(), rw.cpu().numpy(), dn.cpu().numpy().astype(bool), tr.cpu().numpy().astype(bool)
        cnt = int(ref.done_count.item())
        recs = ref.done_records[:cnt].cpu().numpy()
        assert cnt == int(dn.sum()) and np.array_equal(dones, dn) and np.array_equal(rew, rw)
        by_env = {int(r[0]): r for r in recs}
        for i in range(n):
            if dn[i]:
                r = by_env[i]
                tframe, rframe = r[4:19].view(np.float32), r[20:35].view(np.float32)
                assert np.array_equal(rframe, fr[i])                    # the device emits the reset frame for a finished env
                term = np.array(list(stacks[i])[1:] + [tframe])
                assert np.array_equal(infos[i]["terminal_observation"], term), (k, i)
                assert infos[i]["TimeLimit.truncated"] == bool(tr[i]) == bool(r[1] & 1)
                assert infos[i]["episode"]["l"] == int(r[3]) and infos[i]["episode"]["r"] == float(r[2:3].view(np.float32)[0])
                stacks[i] = deque([fr[i]] * 10, maxlen=10)
                finished += 1
            else:
                stacks[i].append(fr[i])
        if k % 10 == 0 or k == steps - 1:
            assert np.array_equal(obs, np.array([np.array(d) for d in stacks])), k
    assert finished > 300
    venv.close()
    ref.close()


def test_single_env_gymnasium_adapter_matches_golden(golden):
    from f16_jsb_b200 import wrap_jsbsim
    t = golden["random3"]
    env = wrap_jsbsim()
    obs, info = env.reset(seed=3)
    assert info == {} and np.allclose(obs, t["reset_obs"], atol=1e-9)
    assert np.array_equal(env.env.goal, t["goal"])
    for k in range(60):
        obs, r, term, trunc, info = env.step(t["actions"][k])
        assert np.allclose(obs[-1], t["frames"][k], rtol=1e-6, atol=1e-6)
        assert abs(float(r) - float(t["rewards"][k])) < 2e-5 and not term and not trunc
    env.close()


def test_torch_fast_path_is_device_resident():
    from f16_jsb_b200 import F16VecEnv
    env = F16VecEnv(256, mode="fp32", host_obs="copy")
    obs = env.reset_torch()
    a = torch.rand((256, 4), device="cuda") * 2 - 1
    o, r, d, tr = env.step_torch(a)
    assert o.is_cuda and r.is_cuda and d.is_cuda and o.data_ptr() == obs.data_ptr()
    env.close()


def test_single_env_adapter_follows_the_reference_episode(golden):
    """f16_jsb_b200.jsbsim_gym (JSBSimEnv + PositionReward, the reference's single-env surface) against an episode of the
    reference's own jsbsim_gym.py (tests/golden/ref_env_traces.npz, FDM = the oracle): same reset(seed) goal, frames,
    shaped rewards, terminated / truncated flags, one host synchronisation per step."""
    from f16_jsb_b200.jsbsim_gym import wrap_jsbsim
    t = golden["random1"]
    env = wrap_jsbsim(mode="fp64")
    obs, info = env.reset(seed=1)
    assert info == {} and obs.shape == (10, 15) and obs.dtype == np.float32
    assert np.array_equal(obs[-1][12:15], t["goal"].astype(np.float32))
    n = len(t["actions"])
    ended = False
    for k in range(n):
        obs, reward, terminated, truncated, info = env.step(t["actions"][k])
        f = t["frames"][k]
        e = float((np.abs(obs[-1][:12] - f[:12]) / np.maximum(np.abs(f[:12]), 1e-2)).max())
        assert e <= (1e-5 if k < 300 else 1e-1), (k, e)
        if k < 300:
            assert abs(float(reward) - float(t["rewards"][k])) < 2e-5
        if terminated or truncated or t["terminated"][k] or t["truncated"][k]:
            assert k >= n - 2
            assert bool(terminated) == bool(t["terminated"][k]) or k < n - 1
            if terminated and t["terminated"][k]:
                assert abs(float(reward) - float(t["rewards"][k])) < 1e-2       # -10 / +10 base reward + shaping
            ended = True
            break
    assert ended
    env.close()
