"""CPU checks of the rollout-store oracle (SB3 RolloutBuffer restated) and of the ABI exports."""
import numpy as np


def test_gae_matches_closed_form_single_env():
    from oracle.rollout_oracle import RolloutBufferOracle
    T, gamma, lam = 5, 0.9, 0.8
    buf = RolloutBufferOracle(T, 1, gae_lambda=lam, gamma=gamma)
    r = np.array([1, 0, 2, 0, 1], np.float32)
    v = np.array([0.5, 0.4, 0.3, 0.2, 0.1], np.float32)
    for t in range(T):
        buf.add(np.zeros((1, 10, 15), np.float32), np.zeros((1, 4), np.float32), r[t:t + 1], np.array([t == 3]), v[t:t + 1], np.zeros(1, np.float32))
    buf.compute_returns_and_advantage(np.array([0.7], np.float32), np.array([False]))
    # hand recursion in float64
    adv = np.zeros(T)
    last = 0.0
    for t in reversed(range(T)):
        nnt = 1.0 if t == T - 1 else (0.0 if t + 1 == 3 else 1.0)
        nv = 0.7 if t == T - 1 else v[t + 1]
        delta = r[t] + gamma * nv * nnt - v[t]
        last = delta + gamma * lam * nnt * last
        adv[t] = last
    assert np.allclose(buf.advantages[:, 0], adv, rtol=1e-6)
    assert np.allclose(buf.returns[:, 0], adv + v, rtol=1e-6)


def test_swap_and_flatten_index_order():
    from oracle.rollout_oracle import RolloutBufferOracle
    a = np.arange(3 * 2).reshape(3, 2).astype(np.float32)      # [T=3][N=2]
    flat = RolloutBufferOracle.swap_and_flatten(a).flatten()
    # flat index = env * T + step
    assert flat[1 * 3 + 2] == a[2, 1] and flat[0 * 3 + 1] == a[1, 0]


def test_rollout_symbols_exported():
    from f16_jsb_b200 import _lib
    L = _lib.load()
    for s in _lib.ROLLOUT_SYMBOLS:
        assert hasattr(L, s)
    assert L.f16_rollout_gae(0, 0, 0.99, 0.95, None, None, None, None, None, None, None, None) != 0


# ---- pinned to the REFERENCE's real RolloutBuffer: tests/golden/sb3_rollout_buffer.npz (tools/make_golden_rollout.py)
def _golden():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sb3_rollout_buffer.npz"))
    return {k: z[k] for k in z.files}


def _fill_oracle(g):
    from oracle.rollout_oracle import RolloutBufferOracle
    T, N = g["in_rewards"].shape
    buf = RolloutBufferOracle(T, N, gae_lambda=float(g["gae_lambda"]), gamma=float(g["gamma"]))
    for t in range(T):
        buf.add(g["in_obs"][t], g["in_actions"][t], g["in_rewards"][t], g["in_episode_starts"][t], g["in_values"][t], g["in_log_probs"][t])
    buf.compute_returns_and_advantage(g["in_last_values"], g["in_last_dones"])
    return buf


def test_rollout_oracle_matches_the_reference_class_bit_for_bit():
    """oracle/rollout_oracle.py against vectors produced by stable_baselines3/common/buffers.py:327-521 itself."""
    g = _golden()
    buf = _fill_oracle(g)
    assert np.array_equal(buf.advantages, g["advantages"]) and np.array_equal(buf.returns, g["returns"])
    assert np.array_equal(buf.episode_starts, g["stored_episode_starts"])
    B = int(g["batch"])
    for k in range(int(g["n_batches"])):
        idx = g["perm"][k * B:(k + 1) * B]
        got = buf.get_samples(idx)
        for a, name in zip(got, ("observations", "actions", "old_values", "old_log_prob", "advantages", "returns")):
            assert np.array_equal(a, g["b%d_%s" % (k, name)]), (k, name)


def test_reference_rollout_buffer_still_produces_the_golden():
    """Re-runs the real class where the reference tree exists (the build container): the committed fixture is what it makes."""
    import os
    import sys
    import pytest
    if not os.path.isdir("/root/reference/stable_baselines3"):
        pytest.skip("needs /root/reference")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    saved = list(sys.path)
    try:
        import make_golden_rollout as mk
        out = mk.run_reference(mk.make_inputs())
    finally:
        sys.path[:] = saved
    g = _golden()
    for k, v in out.items():
        assert np.array_equal(np.asarray(v), g[k]), k
