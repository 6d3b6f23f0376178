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
