"""The REFERENCE's own rollout loop against this repo's VecEnv boundary ("next" item 7).

`stable_baselines3.PPO`, `OnPolicyAlgorithm.collect_rollouts` (common/on_policy_algorithm.py:162-262), `RolloutBuffer`
(common/buffers.py:327-521), `VecEnv`, `DummyVecEnv` and `Monitor` are imported UNMODIFIED from /root/reference (over
oracle/refshim, which stands in for gymnasium / jsbsim / matplotlib) and run twice with the same stub policy and seeds:

  A. on the reference's own env stack - Monitor + DummyVecEnv around jsbsim_gym.py's JSBSim-v0 (FDM = the oracle);
  B. on a VecEnv that replays what F16VecEnv (CUDA, FP64 parity mode, carry-over reset) returned on the B200 for the same
     seeds and actions: tests/golden/f16vecenv_rollout_fp64.npz, recorded by tools/record_vecenv_rollout.py through the
     public F16VecEnv.step call.

What SB3 stores and computes from the two must agree: observations, actions, rewards (including the time-limit
bootstrap gamma * V(terminal_observation) of on_policy_algorithm.py:236-245), episode starts, returns, advantages, and
the Monitor episode statistics. The reference draws the goals of auto-resets from OS entropy; the test injects the goals the
GPU env drew (read from the recorded reset observations) so that both sides fly the same second episodes.
The off-policy side (train.py's SAC option) gets the same treatment: `stable_baselines3.SAC`, `OffPolicyAlgorithm.collect_rollouts`
/ `_store_transition` (common/off_policy_algorithm.py:436-600) and `ReplayBuffer` (common/buffers.py:176-325) run on both
stacks, and what the replay buffer holds - observations, next observations (the terminal stack for finished envs), actions,
rewards, dones, time-limit flags - must agree.
The reference tree does not exist on the GPU box, so this runs where it does (the build container) and skips elsewhere.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FIXTURE = os.path.join(ROOT, "tests", "golden", "f16vecenv_rollout_fp64.npz")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "stable_baselines3")), reason="needs the reference tree (/root/reference)")


@pytest.fixture(scope="module")
def ref():
    """The reference's modules, imported from the read-only mount over the stub packages."""
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), REF]
    try:
        import gymnasium as gym
        import jsbsim_gym.jsbsim_gym as jg
        import torch
        from stable_baselines3 import PPO
        from stable_baselines3.common.callbacks import BaseCallback
        from stable_baselines3.common.monitor import Monitor
        from stable_baselines3.common.vec_env import DummyVecEnv, VecEnv
        yield dict(gym=gym, jg=jg, torch=torch, PPO=PPO, BaseCallback=BaseCallback, Monitor=Monitor, DummyVecEnv=DummyVecEnv, VecEnv=VecEnv)
    finally:
        sys.path[:] = saved


@pytest.fixture(scope="module")
def rec():
    z = np.load(FIXTURE)
    return {k: z[k] for k in z.files}


def make_replay_env(ref, rec):
    """A stock-SB3 VecEnv that hands back, step by step, what F16VecEnv returned on the GPU."""
    VecEnv, spaces = ref["VecEnv"], ref["gym"].spaces

    class ReplayVecEnv(VecEnv):
        def __init__(self):
            self.t = 0
            self.render_mode = None
            obs_space = spaces.Box(low=rec["obs_space_low"], high=rec["obs_space_high"], shape=(10, 15), dtype=np.float32)
            act_space = spaces.Box(low=rec["act_space_low"], high=rec["act_space_high"], shape=(4,), dtype=np.float32)
            super().__init__(int(rec["n_envs"]), obs_space, act_space)
            self.done_at = {}
            for j, (s, e) in enumerate(zip(rec["done_step"], rec["done_env"])):
                self.done_at[(int(s), int(e))] = j

        def reset(self):
            return rec["reset_obs"].copy()

        def step_async(self, actions):
            np.testing.assert_array_equal(np.asarray(actions, np.float32), rec["actions"][self.t])   # the loop fed the recorded actions
            self._a = actions

        def step_wait(self):
            t = self.t
            self.t += 1
            infos = []
            for i in range(self.num_envs):
                j = self.done_at.get((t, i))
                if j is None:
                    infos.append({"TimeLimit.truncated": False})
                else:
                    infos.append({"TimeLimit.truncated": bool(rec["done_truncated"][j]), "terminal_observation": rec["done_terminal_obs"][j].copy(),
                                  "episode": {"r": float(rec["done_ep_r"][j]), "l": int(rec["done_ep_l"][j]), "t": 0.0}})
            return rec["obs"][t].copy(), rec["rewards"][t].copy(), rec["dones"][t].copy(), infos

        def close(self): pass
        def get_attr(self, name, indices=None): return [getattr(self, name, None)] * self.num_envs
        def set_attr(self, name, value, indices=None): setattr(self, name, value)
        def env_method(self, name, *a, indices=None, **k): raise NotImplementedError
        def env_is_wrapped(self, wrapper_class, indices=None): return [False] * self.num_envs

    return ReplayVecEnv()


def stub_policy(model, rec, torch):
    """Actions from the recorded table; values a fixed function of the observation, so that the time-limit bootstrap and
    GAE have something to chew on; no network, no sampling - both runs see exactly the same policy."""
    counter = {"t": 0}

    def value_of(obs):
        obs = obs.reshape(obs.shape[0], 10, 15)
        return (1e-3 * obs[:, -1, 2] + 0.5 * obs[:, -1, 3] - 1e-4 * (obs[:, -1, 12] - obs[:, -1, 0])).reshape(-1, 1).float()

    def forward(obs, deterministic=False):
        a = torch.as_tensor(rec["actions"][counter["t"]])
        counter["t"] += 1
        return a, value_of(obs), torch.zeros(obs.shape[0])

    model.policy.forward = forward
    model.policy.predict_values = value_of
    return counter


def run_collect(ref, rec, env, after_reset=None, after_first_step=None):
    PPO = ref["PPO"]
    T = int(rec["n_steps"])
    model = PPO("MlpPolicy", env, n_steps=T, batch_size=T, n_epochs=1, gamma=0.99, gae_lambda=0.95, seed=0, device="cpu", verbose=0)
    model.env.seed(int(rec["seed"]))

    class AfterFirstStep(ref["BaseCallback"]):
        def _on_step(self):
            if self.n_calls == 1 and after_first_step:
                after_first_step(self.training_env)
            return True

    _, callback = model._setup_learn(T * env.num_envs, AfterFirstStep(), True, "run", False)     # resets the env as PPO.learn does
    if after_reset:
        after_reset(model.env)
    stub_policy(model, rec, ref["torch"])
    assert model.collect_rollouts(model.env, callback, model.rollout_buffer, n_rollout_steps=T)
    return model


import contextlib


@contextlib.contextmanager
def reference_env_stack(ref, rec):
    """Stack A: Monitor + DummyVecEnv around the reference's JSBSim-v0 (FDM = the oracle), with the GPU env's auto-reset goals
    injected. Yields (venv, after_reset, after_first_step) hooks for the algorithm's loop."""
    gym, jg, Monitor, DummyVecEnv = ref["gym"], ref["jg"], ref["Monitor"], ref["DummyVecEnv"]
    N = int(rec["n_envs"])
    goals_after_done = {}
    for s, e in zip(rec["done_step"], rec["done_env"]):
        goals_after_done.setdefault(int(e), []).append(rec["obs"][int(s), int(e), -1, 12:15].astype(np.float64))
    orig_reset = jg.JSBSimEnv.reset

    def reset_with_injected_goal(self, seed=None, options=None):
        obs, info = orig_reset(self, seed=seed, options=options)
        q = getattr(self, "_injected_goals", None)
        if seed is None and q:
            self.goal[:] = q.pop(0)
            first = self._get_current_single_observation()
            self.obs_buffer.clear()
            for _ in range(self.num_stacked_frames):
                self.obs_buffer.append(np.copy(first))
            obs = np.array(self.obs_buffer, dtype=np.float32)
        return obs, info

    jg.JSBSimEnv.reset = reset_with_injected_goal
    try:
        venv = DummyVecEnv([lambda: Monitor(gym.make("JSBSim-v0")) for _ in range(N)])

        def inject(env):
            for i, q in goals_after_done.items():
                env.envs[i].unwrapped._injected_goals = list(q)

        def preset(env):        # as tools/record_vecenv_rollout.py does after the first step
            for i, st in zip(rec["preset_env"], rec["preset_step"]):
                w = env.envs[int(i)]
                w.unwrapped.current_step = int(st)            # jsbsim_gym.py:258-261 truncates on its own counter ...
                while hasattr(w, "env"):
                    if hasattr(w, "_elapsed_steps"):
                        w._elapsed_steps = int(st)           # ... and so does gymnasium's TimeLimit (registration, jsbsim_gym.py:537-545)
                    w = w.env

        yield venv, inject, preset
    finally:
        jg.JSBSimEnv.reset = orig_reset


def test_reference_collect_rollouts_on_recorded_f16vecenv_matches_reference_env_stack(ref, rec):
    N, T = int(rec["n_envs"]), int(rec["n_steps"])

    # ---- A: the reference's env stack, with the GPU env's auto-reset goals injected
    with reference_env_stack(ref, rec) as (venv, inject, preset):
        a = run_collect(ref, rec, venv, after_reset=inject, after_first_step=preset)

    # ---- B: the same loop fed by what F16VecEnv returned on the B200
    b = run_collect(ref, rec, make_replay_env(ref, rec))

    A, B = a.rollout_buffer, b.rollout_buffer
    assert A.observations.shape == B.observations.shape == (T, N, 10, 15)
    # episode structure: exact
    np.testing.assert_array_equal(A.episode_starts, B.episode_starts)
    assert A.episode_starts[1:].sum() == len(rec["done_step"]) == 3
    np.testing.assert_array_equal(A.actions, B.actions)
    # observations: the FP64 kernel against the oracle FDM under the reference's float32 cast chain. Bit-identical for
    # most frames; the stated bound is 1e-5 relative (floor 1e-2) per element over a 320-step free run
    err = np.abs(A.observations - B.observations) / np.maximum(np.abs(A.observations), 1e-2)
    assert err.max() < 1e-5, err.max()
    assert (A.observations == B.observations).mean() > 0.99
    print("reference loop: max relative observation difference %.2e, bit-identical elements %.4f, max reward difference %.2e"
          % (err.max(), (A.observations == B.observations).mean(), np.abs(A.rewards - B.rewards).max()))
    # rewards incl. the time-limit bootstrap (two truncated episodes), values, GAE
    assert np.abs(A.rewards - B.rewards).max() < 2e-5
    boot = [(int(s), int(e)) for s, e, tr in zip(rec["done_step"], rec["done_env"], rec["done_truncated"]) if tr]
    assert len(boot) == 2
    for s, e in boot:
        assert abs(B.rewards[s, e] - rec["rewards"][s, e]) > 0.1          # gamma * V(terminal_observation) was added
    np.testing.assert_allclose(A.values, B.values, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(A.advantages, B.advantages, rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(A.returns, B.returns, rtol=1e-4, atol=2e-4)
    # Monitor's episode statistics as collect_rollouts -> _update_info_buffer keeps them
    ea, eb = list(a.ep_info_buffer), list(b.ep_info_buffer)
    assert [x["l"] for x in ea] == [x["l"] for x in eb] == [51, 101, 249]
    assert np.allclose([x["r"] for x in ea], [x["r"] for x in eb], atol=1e-3)
    assert a.num_timesteps == b.num_timesteps == T * N


def test_replay_fixture_follows_dummy_vec_env_conventions(rec):
    """The recorded F16VecEnv returns: on done the observation is the reset stack (ten copies of one frame carrying the new
    goal), the terminal observation is the shifted previous stack, done envs carry episode stats."""
    for j, (s, e) in enumerate(zip(rec["done_step"], rec["done_env"])):
        o = rec["obs"][s, e]
        assert np.array_equal(o, np.repeat(o[:1], 10, axis=0))
        term = rec["done_terminal_obs"][j]
        assert np.array_equal(term[:-1], rec["obs"][s - 1, e][1:])
        assert not np.array_equal(o[-1, 12:15], term[-1, 12:15])      # new goal
        assert rec["done_ep_l"][j] in (51, 101, 249)
    assert rec["dones"].sum() == len(rec["done_step"])


def run_collect_off_policy(ref, rec, env, after_reset=None, after_first_step=None):
    """SAC's loop (OffPolicyAlgorithm.collect_rollouts, common/off_policy_algorithm.py:506-600, with _store_transition
    :436-504 and ReplayBuffer.add, common/buffers.py:233-290) for one rollout of T steps with the recorded actions."""
    import sys as _sys
    saved = list(_sys.path)
    _sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), REF]
    try:
        from stable_baselines3 import SAC
        from stable_baselines3.common.type_aliases import TrainFreq, TrainFrequencyUnit
    finally:
        _sys.path[:] = saved
    T, N = int(rec["n_steps"]), int(rec["n_envs"])
    model = SAC("MlpPolicy", env, buffer_size=T * N, learning_starts=0, train_freq=(T, "step"), seed=0, device="cpu", verbose=0,
                policy_kwargs=dict(net_arch=[8]))
    model.env.seed(int(rec["seed"]))

    class AfterFirstStep(ref["BaseCallback"]):
        def _on_step(self):
            if self.n_calls == 1 and after_first_step:
                after_first_step(self.training_env)
            return True

    _, callback = model._setup_learn(T * N, AfterFirstStep(), True, "run", False)
    if after_reset:
        after_reset(model.env)
    counter = {"t": 0}

    def sample_action(learning_starts, action_noise=None, n_envs=1):        # the recorded action table instead of the actor
        a = np.asarray(rec["actions"][counter["t"]], np.float32)
        counter["t"] += 1
        return a, model.policy.scale_action(a)

    model._sample_action = sample_action
    out = model.collect_rollouts(model.env, callback, TrainFreq(T, TrainFrequencyUnit.STEP), model.replay_buffer, learning_starts=0)
    assert out.continue_training and out.episode_timesteps == T * N
    return model


def test_reference_sac_collect_rollouts_on_recorded_f16vecenv_matches_reference_env_stack(ref, rec):
    """The off-policy side of the boundary (train.py's SAC option): what SB3's ReplayBuffer holds after the reference's own
    OffPolicyAlgorithm.collect_rollouts must be the same whether the loop ran on the reference env stack or on what
    F16VecEnv returned on the B200 - in particular next_observations, which take info["terminal_observation"] for
    finished envs (off_policy_algorithm.py:473-490), dones, and the time-limit flags (ReplayBuffer.timeouts)."""
    N, T = int(rec["n_envs"]), int(rec["n_steps"])
    with reference_env_stack(ref, rec) as (venv, inject, preset):
        a = run_collect_off_policy(ref, rec, venv, after_reset=inject, after_first_step=preset)
    b = run_collect_off_policy(ref, rec, make_replay_env(ref, rec))
    A, B = a.replay_buffer, b.replay_buffer
    assert A.pos == B.pos and A.full == B.full and A.observations.shape == B.observations.shape == (T, N, 10, 15)
    np.testing.assert_array_equal(A.dones, B.dones)
    np.testing.assert_array_equal(A.timeouts, B.timeouts)
    assert A.dones.sum() == len(rec["done_step"]) == 3 and A.timeouts.sum() == 2
    np.testing.assert_array_equal(A.actions, B.actions)
    for name in ("observations", "next_observations"):
        x, y = getattr(A, name), getattr(B, name)
        err = np.abs(x - y) / np.maximum(np.abs(x), 1e-2)
        assert err.max() < 1e-5, (name, err.max())
        assert (x == y).mean() > 0.99
    assert np.abs(A.rewards - B.rewards).max() < 2e-5
    # a finished env's next observation is its terminal stack, not the reset stack the VecEnv returned
    for j, (s, e) in enumerate(zip(rec["done_step"], rec["done_env"])):
        np.testing.assert_array_equal(B.next_observations[s, e], rec["done_terminal_obs"][j])
        assert not np.array_equal(B.next_observations[s, e], rec["obs"][s, e])
        if s + 1 < T:
            np.testing.assert_array_equal(B.observations[s + 1, e], rec["obs"][s, e])      # the next transition starts from the reset stack
    ea, eb = list(a.ep_info_buffer), list(b.ep_info_buffer)
    assert [x["l"] for x in ea] == [x["l"] for x in eb] == [51, 101, 249]
