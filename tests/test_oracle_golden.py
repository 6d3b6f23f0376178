"""The oracle's C env layer against the golden vectors produced by the REFERENCE's own Python env
file running on the same FDM (tools/make_golden.py): frames, flags and rewards bit-exact
(the distance reproduces NumPy float32 dot = BLAS sdot accumulation, see oracle dist3)."""
import numpy as np
import pytest


@pytest.mark.parametrize("name", ["random0", "random1", "random2", "random3", "gentle0", "gentle1"])
def test_oracle_env_layer_matches_reference_python(oracle, golden, name):
    t = golden[name]
    env = oracle.OracleEnv()
    obs = env.reset(t["goal"])
    assert np.array_equal(obs, t["reset_obs"])
    stacked_at = {0: 0, 4: 1, 11: 2}
    for k, a in enumerate(t["actions"]):
        obs, r, term, trunc = env.step(a)
        assert np.array_equal(obs[-1], t["frames"][k]), "frame differs at step %d" % k
        assert term == bool(t["terminated"][k]) and trunc == bool(t["truncated"][k])
        assert r == t["rewards"][k], "reward differs at step %d" % k   # bit-exact incl. the BLAS sdot accumulation order
        if k in stacked_at:
            assert np.array_equal(obs, t["stacked"][stacked_at[k]])
        if "states" in t:
            assert np.array_equal(env.fdm.pack_state(), t["states"][k + 1])
    assert bool(t["terminated"][-1]) or bool(t["truncated"][-1])


def test_golden_goals_are_the_reference_draws(oracle, golden):
    # goal of JSBSimEnv.reset(seed) = default_rng(seed) uniform x3 (jsbsim_gym.py:312-323)
    for name, seed in (("random0", 0), ("random1", 1), ("random2", 2), ("random3", 3), ("gentle0", 10), ("gentle1", 11)):
        assert np.array_equal(golden[name]["goal"], oracle.sample_goal(seed))


def test_truncation_at_1200_steps(golden):
    t = golden["gentle0"]
    assert len(t["frames"]) == 1200 and t["truncated"][-1] and not t["terminated"][-1]
    assert not t["truncated"][:-1].any()


def test_crash_reward_and_precedence(golden):
    t = golden["random0"]
    assert t["terminated"][-1] and t["frames"][-1][2] < 10.0
    assert -10.5 < t["rewards"][-1] < -9.5        # -10 plus the potential-shaping term (jsbsim_gym.py:506)
