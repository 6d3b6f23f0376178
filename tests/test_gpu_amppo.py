"""AM-PPO on the GPU env (SURVEY.md 8(f) row 3 / BASELINE configs[4]): the LMA extractor on device tensors
against the reference's outputs, and short end-to-end rollout + update runs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lma_extractor_on_device_matches_reference():
    """On CUDA tensors the per-frame transform is the f16_features17 kernel; features must still match the
    reference's StackedLMAFeaturesExtractor output (float32, <= 1e-4)."""
    from f16_jsb_b200.lma import LMAConfig, LMAExtractor
    g = torch.load(os.path.join(ROOT, "tests", "golden", "learner_golden.pt"), weights_only=False)["lma"]
    net = LMAExtractor(LMAConfig()).eval().cuda()
    net.load_state_dict(g["state_dict"])
    with torch.no_grad():
        out = net(g["obs"].cuda()).cpu()
    assert torch.allclose(out, g["features"], rtol=1e-4, atol=1e-4), float((out - g["features"]).abs().max())


@pytest.mark.parametrize("use_am_ppo,optimizer", [(True, "DAG"), (True, "Adam"), (False, "Adam")])
def test_rollout_and_update_run_on_device(use_am_ppo, optimizer):
    from f16_jsb_b200 import F16BatchedEnv
    from f16_jsb_b200.amppo import AMPPO, AMPPOConfig
    env = F16BatchedEnv(256, mode="fp32", seed=2)
    cfg = AMPPOConfig(n_steps=24, batch_size=1024, n_epochs=2, use_am_ppo=use_am_ppo, optimizer=optimizer, seed=4)
    algo = AMPPO(env, cfg)
    before = {k: v.clone() for k, v in algo.policy.state_dict().items()}
    algo.collect_rollouts()
    buf = algo.buffer
    assert buf.full and algo.num_timesteps == 24 * 256
    assert bool(torch.isfinite(buf.advantages).all()) and bool(torch.isfinite(buf.returns).all())
    # SB3 layout: first stored step starts an episode, stored actions are the unclipped samples
    assert bool((buf.episode_starts[0] == 1).all()) and float(buf.actions.abs().max()) > 1.0
    # returns = advantages + values (buffers.py:438)
    assert torch.allclose(buf.returns, buf.advantages + buf.values, atol=1e-5)
    algo.train()
    s = algo.last_stats
    assert all(np.isfinite(v) for v in s.values()), s
    assert s["n_updates"] == 2 and 0 <= s["clip_fraction"] <= 1
    changed = sum(int(not torch.equal(before[k], v)) for k, v in algo.policy.state_dict().items())
    assert changed >= len(before) - 2
    if use_am_ppo:
        assert s["alpha_A_ema"] != 1.0 and s["prev_saturation_A_ema"] != 0.10      # the controller moved once
    else:
        assert s["alpha_A_ema"] == 1.0
    env.close()


def test_learn_two_iterations_with_ring_layout():
    from f16_jsb_b200 import F16BatchedEnv
    from f16_jsb_b200.amppo import AMPPO, AMPPOConfig
    env = F16BatchedEnv(128, mode="fp32", seed=3, obs_layout="ring")
    algo = AMPPO(env, AMPPOConfig(n_steps=16, batch_size=512, n_epochs=1))
    algo.learn(2 * 16 * 128)
    assert algo.num_timesteps == 2 * 16 * 128 and algo.n_updates == 2
    assert np.isfinite(algo.last_stats["value_loss"])
    env.close()
