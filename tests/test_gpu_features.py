"""Feature kernel vs a plain PyTorch float32 restatement of JSBSimFeatureExtractor.forward
(jsbsim_gym/features.py:37-67). Tolerance: 2e-6 absolute on the bounded outputs (sin/cos/normalised
values; both sides use 1-2 ulp float32 libm), exact on the pass-through columns."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def torch_reference(observations):          # features.py:37-67, operation for operation
    position = observations[:, :3]
    mach = observations[:, 3:4]
    alpha_beta = observations[:, 4:6]
    angular_rates = observations[:, 6:9]
    phi_theta = observations[:, 9:11]
    psi = observations[:, 11:12]
    goal = observations[:, 12:]
    displacement = goal - position
    distance = torch.sqrt(torch.sum(displacement[:, :2] ** 2, 1, True))
    dz = displacement[:, 2:3]
    altitude = position[:, 2:3]
    abs_bearing = torch.atan2(displacement[:, 1:2], displacement[:, 0:1])
    rel_bearing = abs_bearing - psi
    dist_norm = 1 / (1 + distance * 1e-3)
    dz_norm = dz / 15000
    alt_norm = altitude / 15000
    cab, sab = torch.cos(alpha_beta), torch.sin(alpha_beta)
    cpt, spt = torch.cos(phi_theta), torch.sin(phi_theta)
    cr, sr = torch.cos(rel_bearing), torch.sin(rel_bearing)
    return torch.concat([dist_norm, dz_norm, alt_norm, mach, angular_rates, cab, sab, cpt, spt, cr, sr], 1)


@pytest.mark.parametrize("n_envs", [1, 37, 4096])
def test_features_match_torch_reference_on_real_observations(n_envs):
    from f16_jsb_b200 import F16BatchedEnv
    from f16_jsb_b200.features import jsbsim_features
    env = F16BatchedEnv(n_envs, mode="fp32", seed=2)
    env.reset()
    for _ in range(25):
        obs, *_ = env.step(None, auto_reset=True)
    got = jsbsim_features(obs)                       # (N, 10, 17)
    want = torch_reference(obs.reshape(-1, 15)).reshape(n_envs, 10, 17)
    assert got.shape == (n_envs, 10, 17)
    assert torch.equal(got[..., 3:7], want[..., 3:7])
    assert float((got - want).abs().max()) < 2e-6


def test_features_synthetic_ranges_and_ragged_sizes():
    from f16_jsb_b200.features import jsbsim_features
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in (5, 33, 1000):
        x = torch.randn((n, 15), device="cuda", generator=g)
        x[:, :3] *= 5000.0
        x[:, 12:] *= 5000.0
        x[:, 9:12] = (torch.rand((n, 3), device="cuda", generator=g) * 2 - 1) * np.pi
        got, want = jsbsim_features(x), torch_reference(x)
        assert float((got - want).abs().max()) < 3e-6
    with pytest.raises(Exception):
        jsbsim_features(torch.zeros((4, 15)))        # CPU tensor: no CPU fallback
