"""The rollout's policy forward as one kernel (csrc/f16_lma_policy.cu, include/f16_lma.h) against the torch modules it
replaces (LMAActorCritic.forward = the reference's ActorCriticPolicy.forward, stable_baselines3/common/policies.py:636-658,
whose extractor is pinned to the reference's own module outputs in tests/test_gpu_amppo.py and tests/test_learner.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _observations(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    u = lambda lo, hi, *s: lo + (hi - lo) * torch.rand(s, generator=g, device="cuda")      # noqa: E731
    obs = torch.empty((n, 10, 15), device="cuda")
    obs[..., 0:2] = u(-20000.0, 20000.0, n, 10, 2)
    obs[..., 2] = u(300.0, 12000.0, n, 10)
    obs[..., 3] = u(0.1, 1.4, n, 10)
    obs[..., 4:6] = u(-0.6, 0.6, n, 10, 2)
    obs[..., 6:9] = u(-2.0, 2.0, n, 10, 3)
    obs[..., 9:12] = u(-3.1, 3.1, n, 10, 3)
    obs[..., 12:14] = u(-10000.0, 10000.0, n, 1, 2).expand(n, 10, 2)
    obs[..., 14] = u(1000.0, 8000.0, n, 1).expand(n, 10)
    return obs.contiguous()


def _policy(seed):
    from f16_jsb_b200.lma import LMAActorCritic
    torch.manual_seed(seed)
    net = LMAActorCritic().cuda().eval()
    with torch.no_grad():                       # leave the initial zeros / ones: every parameter must matter
        for name, q in net.named_parameters():
            if q.ndim == 1:
                q.add_(0.2 * torch.randn_like(q))
        net.action_net.weight.mul_(30.0)        # the 0.01-gain head would hide errors of the mean
    return net


@pytest.mark.parametrize("n", [1, 16, 37, 4096, 5003])
def test_policy_forward_kernel_matches_the_modules(n):
    from f16_jsb_b200.constants import ACTION_HIGH, ACTION_LOW
    from f16_jsb_b200.lma import PolicyForwardKernel
    net = _policy(n)
    low, high = torch.as_tensor(ACTION_LOW).cuda(), torch.as_tensor(ACTION_HIGH).cuda()
    assert PolicyForwardKernel.supported(net)
    fused = PolicyForwardKernel(net, low, high)
    obs = _observations(n, n + 1)
    with torch.no_grad():
        feats_ref = net.features_extractor(obs)
        mean_ref, value_ref = net._heads(obs)
    # deterministic: actions are the mean, log-probability of the mean
    actions, values, log_probs, clipped, feats = [t.clone() for t in fused(obs, None, features=True)]
    assert torch.allclose(feats, feats_ref, rtol=1e-4, atol=2e-5), float((feats - feats_ref).abs().max())
    assert torch.allclose(actions, mean_ref, rtol=1e-4, atol=2e-5), float((actions - mean_ref).abs().max())
    assert torch.allclose(values, value_ref, rtol=1e-4, atol=2e-5), float((values - value_ref).abs().max())
    with torch.no_grad():
        assert torch.allclose(log_probs, net._log_prob(mean_ref, mean_ref), rtol=1e-5, atol=1e-5)
    assert torch.equal(clipped, torch.maximum(torch.minimum(actions, high), low))
    # sampled: actions = mean + exp(log_std) * noise, the Gaussian's log-probability, the clip
    noise = torch.randn((n, 4), device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    a2, v2, lp2, c2 = [t.clone() for t in fused(obs, noise)]
    with torch.no_grad():
        want = mean_ref + torch.exp(net.log_std) * noise
        assert torch.allclose(a2, want, rtol=1e-4, atol=3e-5), float((a2 - want).abs().max())
        assert torch.allclose(lp2, net._log_prob(actions, a2), rtol=1e-5, atol=1e-4), float((lp2 - net._log_prob(actions, a2)).abs().max())
    assert torch.equal(v2, values) and torch.equal(c2, torch.maximum(torch.minimum(a2, high), low))
    assert bool((c2 != a2).any()) or n < 16          # the box is narrower than the noise: the clip is exercised


@pytest.mark.parametrize("n", [5, 1000])
def test_policy_forward_kernel_matches_the_numpy_oracle(n):
    """The kernel against oracle/policy_forward_oracle.py (float64 NumPy restatement of the reference's forward pass, pinned on
    the CPU in tests/test_policy_oracle.py) on the very buffer the kernel reads: 2e-5 of the magnitude."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import policy_forward_oracle as oracle
    from f16_jsb_b200.constants import ACTION_HIGH, ACTION_LOW
    from f16_jsb_b200.lma import PolicyForwardKernel, PolicyPacker
    net = _policy(100 + n)
    fused = PolicyForwardKernel(net, torch.as_tensor(ACTION_LOW).cuda(), torch.as_tensor(ACTION_HIGH).cuda())
    obs = _observations(n, 7)
    noise = torch.randn((n, 4), device="cuda", generator=torch.Generator(device="cuda").manual_seed(8))
    actions, values, log_probs, clipped, feats = [t.cpu().numpy().astype(np.float64) for t in fused(obs, noise, features=True)]
    want = oracle.forward(obs.cpu().numpy(), fused.packed.cpu().numpy(), PolicyPacker.entries(), net.log_std.detach().cpu().numpy(),
                          noise.cpu().numpy(), ACTION_LOW, ACTION_HIGH)
    for name, got, tol in (("features", feats, 2e-5), ("actions", actions, 2e-5), ("values", values, 2e-5), ("log_probs", log_probs, 1e-4),
                           ("clipped", clipped, 2e-5)):
        err = float(np.abs(got - want[name]).max())
        assert err <= tol * max(1.0, float(np.abs(want[name]).max())), (name, err)


def test_policy_forward_kernel_follows_the_weights_and_the_golden_extractor():
    """refresh() re-packs in place; with the reference's recorded extractor weights the kernel's features equal the reference
    module's recorded outputs (tests/golden/learner_golden.pt, tools/make_golden_learner.py) to 1e-4."""
    from f16_jsb_b200.constants import ACTION_HIGH, ACTION_LOW
    from f16_jsb_b200.lma import LMAActorCritic, PolicyForwardKernel
    g = torch.load(os.path.join(ROOT, "tests", "golden", "learner_golden.pt"), weights_only=False)["lma"]
    torch.manual_seed(0)
    net = LMAActorCritic().cuda().eval()
    fused = PolicyForwardKernel(net, torch.as_tensor(ACTION_LOW).cuda(), torch.as_tensor(ACTION_HIGH).cuda())
    obs = g["obs"].cuda().float().contiguous()
    before = fused(obs, None, features=True)[4].clone()
    net.features_extractor.load_state_dict(g["state_dict"])
    ptr = fused.packed.data_ptr()
    fused.refresh()
    assert fused.packed.data_ptr() == ptr
    after = fused(obs, None, features=True)[4].cpu()
    assert not torch.allclose(before.cpu(), after)
    assert torch.allclose(after, g["features"], rtol=1e-4, atol=1e-4), float((after - g["features"]).abs().max())


def test_rollout_with_the_fused_forward_matches_the_module_rollout():
    """AMPPO.collect_rollouts with and without the fused kernel (same seeds, same noise stream, no graph): the same transitions
    to float rounding over a short rollout. Under the CUDA graph (whose random stream differs from eager mode's) the replayed
    kernel must agree with the modules on the observations it saw, before and after an update has moved the weights."""
    from f16_jsb_b200 import F16BatchedEnv
    from f16_jsb_b200.amppo import AMPPO, AMPPOConfig
    got = {}
    for fused in (True, False):
        env = F16BatchedEnv(512, mode="fp32", seed=3)
        algo = AMPPO(env, AMPPOConfig(n_steps=6, batch_size=1024, n_epochs=1, fused_policy_forward=fused, cuda_graph=False, seed=4))
        assert (algo._fused_act is not None) == fused
        torch.manual_seed(9)
        algo.collect_rollouts()
        b = algo.buffer
        got[fused] = [t.clone() for t in (b.actions, b.values, b.log_probs, b.rewards, b.advantages)]
        env.close()
    for x, y, tol in zip(got[True], got[False], (2e-5, 2e-5, 1e-4, 1e-4, 1e-3)):
        assert torch.allclose(x, y, rtol=1e-4, atol=tol), float((x - y).abs().max())

    env = F16BatchedEnv(512, mode="fp32", seed=3)
    algo = AMPPO(env, AMPPOConfig(n_steps=6, batch_size=1024, n_epochs=1, cuda_graph=True, seed=4))
    for it in range(2):
        algo.collect_rollouts()
        assert algo._graph is not None and algo._fused_act is not None
        with torch.no_grad():
            algo.policy.eval()
            mean_ref, v_ref = algo.policy._heads(algo._g_obs.clone())
        actions, values, log_probs, clipped = algo._g_out
        assert torch.allclose(values, v_ref, rtol=1e-4, atol=2e-5), float((values - v_ref).abs().max())
        with torch.no_grad():
            assert torch.allclose(log_probs, algo.policy._log_prob(mean_ref, actions), rtol=1e-4, atol=1e-4)
        assert float((actions - mean_ref).std()) > 0.5            # the graph draws fresh noise
        assert torch.equal(algo.buffer.values[-1], values)
        packed = algo._fused_act.packed.clone()
        algo.train()
        if it == 0:
            continue
        algo._fused_act.refresh()
        assert not torch.equal(packed, algo._fused_act.packed)
    env.close()
