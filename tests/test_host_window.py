"""Host-resident observation windows (include/f16_hostwin.h) against a deque model of the reference's stack
(jsbsim_gym/jsbsim_gym.py:150,235,263,325-329) and DummyVecEnv's terminal_observation / auto-reset convention
(stable_baselines3/common/vec_env/dummy_vec_env.py:63-72). Host-only entry points: no GPU needed."""
from collections import deque

import numpy as np
import pytest

from f16_jsb_b200.host_window import RECORD_DTYPE, HostWindow


class DequeModel:
    def __init__(self, frames0):
        self.stacks = [deque([f.copy() for _ in range(10)], maxlen=10) for f in frames0]

    def step(self, frames, done, reset_frames):
        """frames: what the env produced this step (terminal frame for finished envs)."""
        term = {}
        for i, dq in enumerate(self.stacks):
            dq.append(frames[i].copy())
            if done[i]:
                term[i] = np.array(dq)
                self.stacks[i] = deque([reset_frames[i].copy() for _ in range(10)], maxlen=10)
        return np.array([np.array(dq) for dq in self.stacks]), term


@pytest.mark.parametrize("alias", [True, False])
@pytest.mark.parametrize("n_rings,dma_both", [(1, False), (2, False), (2, True)])
@pytest.mark.parametrize("n", [1, 37, 1500])
def test_windows_match_the_deque_model(n, n_rings, dma_both, alias):
    rng = np.random.default_rng(n * 7 + n_rings)
    w = HostWindow(n, n_rings=n_rings, pin=False, alias=alias, dma_both=dma_both)
    assert all(a == alias for a in w.aliased) or alias     # the aliased mapping may be unavailable; the mirror never is
    f0 = rng.normal(size=(n, 15)).astype(np.float32)
    model = DequeModel(f0)
    res = w.fill(f0)
    assert res.obs.shape == (n, 10, 15) and res.obs.dtype == np.float32
    np.testing.assert_array_equal(res.obs, np.repeat(f0[:, None, :], 10, axis=1))
    prev_view, prev_expected = res.obs, np.array(res.obs)
    for t in range(40):
        frames = rng.normal(size=(n, 15)).astype(np.float32)
        done = rng.random(n) < (0.5 if t in (3, 4) else 0.08)          # steps 3/4: many envs finish back to back
        reset_frames = rng.normal(size=(n, 15)).astype(np.float32)
        idx = np.flatnonzero(done)
        rng.shuffle(idx)                                                # the kernel appends in no particular order
        rec = np.zeros(idx.size, dtype=RECORD_DTYPE)
        rec["env"] = idx
        rec["ep_len"] = t + 1
        rec["terminal_frame"][:, :15] = frames[idx]
        rec["reset_frame"][:, :15] = reset_frames[idx]
        sent = frames.copy()
        sent[idx] = reset_frames[idx]                                   # the device emits the reset frame for finished envs
        reward = rng.normal(size=n).astype(np.float32)
        res = w.push(sent, reward, done.astype(np.uint8), np.zeros(n, np.uint8), rec)
        expected, term = model.step(frames, done, reset_frames)
        np.testing.assert_array_equal(res.obs, expected)
        np.testing.assert_array_equal(res.reward, reward)
        np.testing.assert_array_equal(res.done.astype(bool), done)
        assert res.records.size == idx.size and res.terminal_obs.shape == (idx.size, 10, 15)
        for j, i in enumerate(res.records["env"]):
            np.testing.assert_array_equal(res.terminal_obs[j], term[int(i)])
        if n_rings == 2:
            # SB3 stores `_last_obs` after the next env.step (on_policy_algorithm.py:247): still intact
            np.testing.assert_array_equal(prev_view, prev_expected)
        prev_view, prev_expected = res.obs, expected
    w.close()


@pytest.mark.parametrize("alias", [True, False])
def test_carry_over_between_rings_on_host_threads(alias):
    """Slots above 1 MB are carried over to the second ring by the library's worker threads between two
    steps; a vectorised stack model checks every returned window and that the previous one stays intact."""
    n = 20000
    rng = np.random.default_rng(3)
    w = HostWindow(n, n_rings=2, pin=False, alias=alias)
    f0 = rng.normal(size=(n, 15)).astype(np.float32)
    expected = np.repeat(f0[:, None, :], 10, axis=1)
    res = w.fill(f0)
    prev_view, prev_expected = res.obs, expected.copy()
    for t in range(25):
        frames = rng.normal(size=(n, 15)).astype(np.float32)
        done = rng.random(n) < 0.05
        reset_frames = rng.normal(size=(n, 15)).astype(np.float32)
        idx = np.flatnonzero(done)
        rec = np.zeros(idx.size, dtype=RECORD_DTYPE)
        rec["env"] = idx
        rec["terminal_frame"][:, :15] = frames[idx]
        rec["reset_frame"][:, :15] = reset_frames[idx]
        sent = frames.copy()
        sent[idx] = reset_frames[idx]
        res = w.push(sent, None, done.astype(np.uint8), None, rec)
        expected = np.concatenate([expected[:, 1:], frames[:, None, :]], axis=1)
        np.testing.assert_array_equal(res.terminal_obs, expected[res.records["env"]])
        expected[idx] = reset_frames[idx][:, None, :]
        np.testing.assert_array_equal(res.obs, expected)
        np.testing.assert_array_equal(prev_view, prev_expected)
        prev_view, prev_expected = res.obs, expected.copy()
    w.close()


def test_window_is_a_view_not_a_copy():
    w = HostWindow(64, n_rings=1, pin=False)
    res = w.fill(np.zeros((64, 15), np.float32))
    assert not res.obs.flags.owndata
    assert res.obs.strides[0] == 60 and res.obs.strides[2] == 4 and res.obs.strides[1] % 4096 == 0
    # torch can wrap the strided view without copying (SB3's obs_as_tensor path)
    import torch
    t = torch.as_tensor(res.obs)
    assert tuple(t.shape) == (64, 10, 15)
    w.close()


def test_bad_arguments_are_reported():
    from f16_jsb_b200 import _lib
    with pytest.raises(_lib.F16Error):
        HostWindow(0, pin=False)
    with pytest.raises(_lib.F16Error):
        HostWindow(8, n_rings=3, pin=False)
    w = HostWindow(8, pin=False)
    rec = np.zeros(1, dtype=RECORD_DTYPE)
    rec["env"] = 99
    with pytest.raises(_lib.F16Error):
        w.push(np.zeros((8, 15), np.float32), None, None, None, rec)
    w.close()


def test_pinning_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from f16_jsb_b200 import _lib
    with pytest.raises(_lib.F16Error):
        HostWindow(8, pin=True)


@pytest.mark.parametrize("n_rings,alias", [(1, True), (2, True), (2, False)])
def test_gather_returns_an_owned_contiguous_copy_of_the_window(n_rings, alias):
    """f16_hostwin_gather (copy_obs=True of F16VecEnv: DummyVecEnv hands out owned arrays): the threaded copy equals the
    strided view it was taken from and does not change when the ring moves on."""
    n = 5000
    rng = np.random.default_rng(3)
    w = HostWindow(n, n_rings=n_rings, pin=False, alias=alias)
    res = w.fill(rng.normal(size=(n, 15)).astype(np.float32))
    for t in range(13):
        frames = rng.normal(size=(n, 15)).astype(np.float32)
        res = w.push(frames, None, None, None, np.empty(0, dtype=RECORD_DTYPE))
        got = w.gather(res.ring, res.first_slot)
        assert got.flags.owndata and got.flags.c_contiguous and got.shape == (n, 10, 15)
        np.testing.assert_array_equal(got, res.obs)
        if t == 5:
            kept, kept_expected = got, got.copy()
    np.testing.assert_array_equal(kept, kept_expected)
    w.close()
