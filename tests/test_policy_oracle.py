"""The NumPy checker of the fused policy forward (oracle/policy_forward_oracle.py) and the packed parameter layout of
include/f16_lma.h, without a GPU: the oracle reads the buffer the product's packer wrote - through the layout the header
documents, re-derived independently - and must agree with the torch modules (LMAActorCritic, pinned to the reference's
modules in tests/test_learner.py) and with the reference extractor's own recorded outputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import policy_forward_oracle as oracle  # noqa: E402

from f16_jsb_b200.constants import ACTION_HIGH, ACTION_LOW  # noqa: E402
from f16_jsb_b200.lma import LMAActorCritic, PolicyPacker  # noqa: E402


def _observations(n, seed):
    rng = np.random.default_rng(seed)
    obs = np.empty((n, 10, 15), dtype=np.float32)
    obs[..., 0:2] = rng.uniform(-20000, 20000, (n, 10, 2))
    obs[..., 2] = rng.uniform(300, 12000, (n, 10))
    obs[..., 3] = rng.uniform(0.1, 1.4, (n, 10))
    obs[..., 4:6] = rng.uniform(-0.6, 0.6, (n, 10, 2))
    obs[..., 6:9] = rng.uniform(-2, 2, (n, 10, 3))
    obs[..., 9:12] = rng.uniform(-3.1, 3.1, (n, 10, 3))
    obs[..., 12:14] = rng.uniform(-10000, 10000, (n, 1, 2))
    obs[..., 14] = rng.uniform(1000, 8000, (n, 1))
    return obs


def test_packed_layout_roundtrip_and_oracle_matches_the_modules():
    torch.manual_seed(3)
    net = LMAActorCritic().eval()
    with torch.no_grad():
        for q in net.parameters():
            if q.ndim == 1:
                q.add_(0.2 * torch.randn_like(q))
        net.action_net.weight.mul_(30.0)
    packer = PolicyPacker(net)
    ents = PolicyPacker.entries()
    # every parameter comes back out of the buffer exactly, through the header's index arithmetic
    P = oracle.unpack(packer.packed.numpy(), ents)
    core = net.features_extractor.lma_extractor
    assert np.array_equal(P[0][0], core.initial_transform.positions.numpy().astype(np.float64))
    lin = [core.initial_transform.input_embedding, core.initial_transform.embed_layer_2]
    for blk in core.lma_blocks:
        lin += [blk.ln_1, blk.attn.c_attn, blk.attn.c_proj, blk.ln_2, blk.mlp.c_fc, blk.mlp.c_proj]
    lin += [net.mlp_extractor.policy_net[0], net.mlp_extractor.policy_net[2], net.action_net, net.mlp_extractor.value_net[0],
            net.mlp_extractor.value_net[2], net.value_net]
    for (w, b), m in zip(P[1:], lin):
        assert np.array_equal(w, m.weight.detach().numpy().astype(np.float64)) and np.array_equal(b, m.bias.detach().numpy().astype(np.float64))
    # refresh() re-packs in place
    ptr = packer.packed.data_ptr()
    with torch.no_grad():
        net.value_net.bias.add_(1.0)
    packer.refresh()
    assert packer.packed.data_ptr() == ptr and oracle.unpack(packer.packed.numpy(), ents)[20][1][0] == float(net.value_net.bias.detach()[0])
    # the restated forward against the modules
    obs = _observations(64, 1)
    noise = np.random.default_rng(2).standard_normal((64, 4)).astype(np.float32)
    got = oracle.forward(obs, packer.packed.numpy(), ents, net.log_std.detach().numpy(), noise, ACTION_LOW, ACTION_HIGH)
    with torch.no_grad():
        t = torch.from_numpy(obs)
        feats = net.features_extractor(t)
        mean, value = net._heads(t)
        actions = mean + torch.exp(net.log_std) * torch.from_numpy(noise)
        logp = net._log_prob(mean, actions)
    for name, ref, tol in (("features", feats, 2e-5), ("mean", mean, 2e-5), ("values", value, 2e-5), ("actions", actions, 2e-5), ("log_probs", logp, 1e-4)):
        err = float(np.abs(got[name] - ref.numpy().astype(np.float64)).max())
        assert err <= tol * max(1.0, float(ref.abs().max())), (name, err)
    assert np.array_equal(got["clipped"], np.clip(got["actions"], ACTION_LOW.astype(np.float64), ACTION_HIGH.astype(np.float64)))


def test_oracle_reproduces_the_reference_extractor_outputs():
    """With the reference's recorded extractor weights the oracle's features are the reference module's recorded outputs
    (tests/golden/learner_golden.pt, written by tools/make_golden_learner.py from jsbsim_gym/LMA_features.py unmodified)."""
    g = torch.load(os.path.join(ROOT, "tests", "golden", "learner_golden.pt"), weights_only=False)["lma"]
    net = LMAActorCritic().eval()
    net.features_extractor.load_state_dict(g["state_dict"])
    packer = PolicyPacker(net)
    got = oracle.forward(g["obs"].numpy(), packer.packed.numpy(), PolicyPacker.entries(), net.log_std.detach().numpy())
    err = float(np.abs(got["features"] - g["features"].numpy().astype(np.float64)).max())
    assert err <= 2e-5 * max(1.0, float(g["features"].abs().max())), err


def test_philox_known_answers_and_keep_rate():
    """The generator of the dropout-mask checker against Random123's published Philox4x32-10 known-answer vectors (kat_vectors:
    zero counter and key; all ones; the digits of pi), and the keep rate / scale of the mask it builds."""
    import dropout_mask_oracle as dm
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = dm.philox4x32_10(*[np.uint32(c) for c in ctr], *key)
        assert tuple(int(x) for x in got) == want, [hex(int(x)) for x in got]
    m = dm.keep_factors(1 << 20, 0.1, 0x123456789abcdef)
    thr = round(0.1 * 65536)
    assert set(np.unique(m).tolist()) == {0.0, float(np.float32(65536.0) / np.float32(65536 - thr))}
    assert abs(float((m > 0).mean()) - (1 - thr / 65536)) < 2e-3
    assert not np.array_equal(m, dm.keep_factors(1 << 20, 0.1, 0x123456789abcdf0))
    assert bool((dm.keep_factors(64, 0.0, 5) == 1.0).all())
