"""Learner-side rows of SURVEY.md 8(f) against golden vectors produced by executing the reference's own Python
(tools/make_golden_learner.py -> tests/golden/learner_golden.pt): the LMA feature extractor
(jsbsim_gym/LMA_features.py), AM-PPO's advantage modulation (stable_baselines3/ppo/ppo.py:29-99) and the DAG
optimizer (stable_baselines3/ppo/optim/sgd.py:87-344). CPU tensors: no GPU needed."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "learner_golden.pt"), weights_only=False)


def test_closest_divisor_matches_reference(gold):
    from f16_jsb_b200.lma import LMAConfig, closest_divisor
    for total, target, want in gold["divisors"]:
        assert closest_divisor(total, target) == want, (total, target)
    c = LMAConfig()
    assert (c.l_new, c.c_new, c.d_new, c.features_dim) == (gold["lma"]["L_new"], gold["lma"]["C_new"], 32, gold["lma"]["features_dim"])


def test_lma_extractor_matches_reference_forward_and_backward(gold):
    """Same state_dict keys as the reference (checkpoints interchange); features within 2e-5 (float32, different
    but equivalent op order); gradients of a scalar loss within 1e-4 relative."""
    from f16_jsb_b200.lma import LMAConfig, LMAExtractor
    g = gold["lma"]
    net = LMAExtractor(LMAConfig(dropout=0.1)).eval()
    assert set(net.state_dict().keys()) == set(g["state_dict"].keys())
    net.load_state_dict(g["state_dict"])
    with torch.no_grad():
        out = net(g["obs"])
    assert out.shape == g["features"].shape == (24, 160)
    assert torch.allclose(out, g["features"], rtol=2e-5, atol=2e-5), float((out - g["features"]).abs().max())
    net0 = LMAExtractor(LMAConfig(dropout=0.0)).train()
    net0.load_state_dict(g["state_dict"])
    loss = (net0(g["obs"]) ** 2).mean()
    loss.backward()
    assert torch.allclose(loss, g["loss"], rtol=1e-5)
    named = dict(net0.named_parameters())
    for k, want in g["grads"].items():
        got = named[k].grad
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-6 * float(want.abs().max())), k


def test_actor_critic_shapes_and_distribution():
    from f16_jsb_b200.lma import LMAActorCritic
    torch.manual_seed(0)
    pol = LMAActorCritic().eval()           # dropout off: repeated forwards agree
    n_params = sum(p.numel() for p in pol.parameters())
    assert n_params == 30688 + (160 * 64 + 64 + 64 * 64 + 64) + (160 * 128 + 128 + 128 * 64 + 64) + (64 * 4 + 4) + (64 + 1) + 4
    obs = torch.randn(7, 10, 15)
    a, v, lp = pol(obs)
    assert a.shape == (7, 4) and v.shape == (7,) and lp.shape == (7,)
    v2, lp2, ent = pol.evaluate_actions(obs, a)
    dist = torch.distributions.Normal(pol.action_net(pol.mlp_extractor.policy_net(pol.features_extractor(obs))), torch.exp(pol.log_std))
    assert torch.allclose(lp2, dist.log_prob(a).sum(-1), atol=1e-5) and torch.allclose(ent, dist.entropy().sum(-1), atol=1e-6)
    assert torch.allclose(pol(obs, deterministic=True)[0], dist.mean, atol=1e-6)


def test_advantage_modulation_matches_reference(gold):
    from f16_jsb_b200.amppo import DynagoConfig, modulate_advantages
    d = gold["dynago"]
    assert DynagoConfig().params() == d["params"]
    for c in d["cases"]:
        alpha, sat = c["alpha_in"].clone(), c["sat_in"].clone()
        mod = modulate_advantages(c["adv"].clone(), d["params"], alpha, sat, c["kappa"], c["v_shift"], update_ema=c["update"])
        assert torch.allclose(mod, c["mod"], rtol=1e-6, atol=1e-7), c["adv"].numel()
        assert torch.allclose(alpha, c["alpha_out"], rtol=1e-6) and torch.allclose(sat, c["sat_out"], rtol=1e-6)


def test_dag_optimizer_follows_the_reference_trajectory(gold):
    from f16_jsb_b200.dag import DAG
    d = gold["dag"]
    model = torch.nn.Sequential(torch.nn.Linear(17, 32), torch.nn.Tanh(), torch.nn.Linear(32, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    model.load_state_dict(d["init"])
    opt = DAG(model.parameters(), lr=d["lr"], shrink=d["shrink"])
    assert abs(opt.h["kappa"] - d["kappa"]) < 1e-6
    for k, want in enumerate(d["traj"]):
        opt.zero_grad()
        loss = ((model(d["xs"][k]) - d["ys"][k]) ** 2).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
        assert torch.allclose(loss.detach(), want["loss"], rtol=1e-5), k
        assert abs(opt.s_t - want["s_t"]) < 1e-5, (k, opt.s_t, want["s_t"])
        for name, p in model.state_dict().items():
            assert torch.allclose(p, want["params"][name], rtol=1e-5, atol=1e-7), (k, name)
    assert opt.s_t < 0.6          # the RMS-shrink branch was exercised


def test_dag_checkpoint_resume_equals_uninterrupted(gold, tmp_path):
    """DAG's per-tensor alpha / saturation statistics, the RMS-shrink state and the step counter must survive
    state_dict() -> torch.save -> a NEW optimizer object's load_state_dict(): 6 steps + resume + 6 steps == 12 steps."""
    from f16_jsb_b200.dag import DAG
    d = gold["dag"]

    def make():
        m = torch.nn.Sequential(torch.nn.Linear(17, 32), torch.nn.Tanh(), torch.nn.Linear(32, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
        m.load_state_dict(d["init"])
        return m, DAG(m.parameters(), lr=d["lr"], shrink=d["shrink"], momentum=0.5)

    def run(m, opt, ks):
        for k in ks:
            opt.zero_grad()
            ((m(d["xs"][k]) - d["ys"][k]) ** 2).mean().backward()
            opt.step()

    n = len(d["traj"])
    ma, oa = make()
    run(ma, oa, range(n))                                  # uninterrupted
    mb, ob = make()
    run(mb, ob, range(n // 2))
    torch.save({"model": mb.state_dict(), "opt": ob.state_dict()}, tmp_path / "ck.pt")
    ck = torch.load(tmp_path / "ck.pt", weights_only=False)
    mc, oc = make()                                        # a new process would build these afresh
    mc.load_state_dict(ck["model"])
    oc.load_state_dict(ck["opt"])
    assert oc.global_step == n // 2 and oc.s_t == ob.s_t
    assert torch.equal(oc.state["_group_0"]["alpha"], ob.state["_group_0"]["alpha"]) and not torch.equal(oc.state["_group_0"]["alpha"], torch.ones(6))
    run(mc, oc, range(n // 2, n))
    for (name, pa), pc in zip(ma.state_dict().items(), mc.state_dict().values()):
        assert torch.equal(pa, pc), name
    assert oa.s_t == oc.s_t and oa.global_step == oc.global_step


def _modulate_worker(rank, world, port, adv, params, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from f16_jsb_b200.amppo import modulate_advantages
    alpha, sat = torch.tensor([1.0]), torch.tensor([0.1])
    n = adv.numel() // world
    mod = modulate_advantages(adv[rank * n:(rank + 1) * n].clone(), params, alpha, sat, 2.0, 0.0, update_ema=True, group=dist.group.WORLD)
    out[rank] = (mod, alpha.clone(), sat.clone())
    dist.destroy_process_group()


def test_sharded_modulation_uses_whole_rollout_moments():
    """Two gloo ranks, each holding half of the advantages: the all-reduced sums give every rank the moments
    of the whole rollout, so the result equals the single-process one."""
    import torch.multiprocessing as mp
    from f16_jsb_b200.amppo import DynagoConfig, modulate_advantages
    params = DynagoConfig().params()
    adv = torch.randn(4096, generator=torch.Generator().manual_seed(5)) * 0.7 + 0.1
    alpha, sat = torch.tensor([1.0]), torch.tensor([0.1])
    want = modulate_advantages(adv.clone(), params, alpha, sat, 2.0, 0.0, update_ema=True)
    with mp.Manager() as m:
        out = m.dict()
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_modulate_worker, args=(2, port, adv, params, out), nprocs=2, join=True)
        got = torch.cat([out[0][0], out[1][0]])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)
        for r in (0, 1):
            assert torch.allclose(out[r][1], alpha, rtol=1e-5) and torch.allclose(out[r][2], sat, rtol=1e-5)
