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


def _attention_reference(qkv, heads, mask=None):
    """Plain PyTorch float32: softmax(q k^T / sqrt(dh)) [* mask] v."""
    b, t, d3 = qkv.shape
    d = d3 // 3
    q, k, v = qkv.view(b, t, 3, heads, d // heads).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-2, -1) / (d // heads) ** 0.5, dim=-1)
    if mask is not None:
        p = p * mask
    return (p @ v).transpose(1, 2).reshape(b, t, d)


@pytest.mark.parametrize("batch,dropout", [(1, 0.0), (1000, 0.0), (4097, 0.1), (333, 0.5)])
def test_latent_attention_kernels_match_pytorch(batch, dropout):
    """Forward and backward of the hand-written latent-attention kernels (include/f16_lma.h) against a plain
    PyTorch float32 reference of the same op; with dropout the reference multiplies by the kernels' own keep mask
    (f16_lma_attention_mask). Tolerance 2e-5 relative to the largest value."""
    import ctypes as C

    from f16_jsb_b200 import _lib
    from f16_jsb_b200.lma import _LatentAttentionFn
    g = torch.Generator(device="cuda").manual_seed(batch)
    qkv = (torch.randn((batch, 5, 96), device="cuda", generator=g) * 1.5).requires_grad_(True)
    dy = torch.randn((batch, 5, 32), device="cuda", generator=g)
    torch.manual_seed(7)
    y = _LatentAttentionFn.apply(qkv, 4, dropout)
    seed = y.grad_fn.meta[2]
    mask = None
    if dropout > 0:
        mask = torch.empty((batch, 4, 5, 5), device="cuda")
        _lib.check(_lib.load().f16_lma_attention_mask(batch, 5, 4, C.c_void_p(mask.data_ptr()), float(dropout), seed,
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mask")
        kept = float((mask > 0).float().mean())
        assert set(mask.unique().tolist()) <= {0.0, float(mask.max())} and abs(float(mask.max()) - 1 / (1 - dropout)) < 1e-3
        if batch > 1000:
            assert abs(kept - (1 - dropout)) < 0.01
    (gq,) = torch.autograd.grad(y, qkv, dy)
    ref_in = qkv.detach().clone().requires_grad_(True)
    y_ref = _attention_reference(ref_in, 4, mask)
    (gq_ref,) = torch.autograd.grad(y_ref, ref_in, dy)
    assert torch.allclose(y, y_ref, rtol=2e-5, atol=2e-5 * float(y_ref.abs().max()))
    assert torch.allclose(gq, gq_ref, rtol=2e-5, atol=2e-5 * float(gq_ref.abs().max())), float((gq - gq_ref).abs().max())


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


@pytest.mark.parametrize("optimizer", ["dag", "adam"])
def test_amppo_checkpoint_restores_the_learner(optimizer, tmp_path):
    """train.py always saves its model; AMPPO.save / load must carry the policy, the optimizer (DAG statistics
    included), the two advantage-modulation EMAs, the sampling generator and the counters, so that a learner rebuilt in a
    new process continues exactly like the one that kept running: the same update on the same rollout."""
    from f16_jsb_b200 import F16BatchedEnv
    from f16_jsb_b200.amppo import AMPPO, AMPPOConfig
    cfg = AMPPOConfig(n_steps=16, batch_size=512, n_epochs=2, use_am_ppo=True, optimizer=optimizer, seed=4, cuda_graph=False)
    env = F16BatchedEnv(128, mode="fp32", seed=2)
    a = AMPPO(env, cfg)
    a.collect_rollouts()
    a.train()
    a.collect_rollouts()
    a.save(str(tmp_path / "amppo.pt"))
    env_b = F16BatchedEnv(128, mode="fp32", seed=2)
    b = AMPPO(env_b, cfg).load(str(tmp_path / "amppo.pt"))
    assert b.num_timesteps == a.num_timesteps and b.n_updates == a.n_updates
    assert torch.equal(b.alpha_state, a.alpha_state) and torch.equal(b.sat_state, a.sat_state) and float(a.alpha_state[0]) != 1.0
    for (k, va), vb in zip(a.policy.state_dict().items(), b.policy.state_dict().values()):
        assert torch.equal(va, vb), k
    # the same rollout in both buffers, then one update each
    for name in ("frames", "age", "actions", "rewards", "episode_starts", "values", "log_probs", "advantages", "returns"):
        getattr(b.buffer, name).copy_(getattr(a.buffer, name))
    b.buffer.pos, b.buffer.full = a.buffer.pos, a.buffer.full
    torch.manual_seed(77)          # dropout masks come from torch's global generator (not learner state, as in SB3)
    a.train()
    torch.manual_seed(77)
    b.train()
    # (the weight-gradient kernel merges its row slabs with float atomics, so two runs of the same update agree to a
    # few ulp of the gradient, not bit for bit: 1e-5 of a parameter step of ~1e-3)
    for (k, va), vb in zip(a.policy.state_dict().items(), b.policy.state_dict().values()):
        assert torch.allclose(va, vb, rtol=1e-4, atol=2e-6), (k, float((va - vb).abs().max()))
    assert abs(a.last_stats["alpha_A_ema"] - b.last_stats["alpha_A_ema"]) < 1e-6 and a.n_updates == b.n_updates
    if optimizer == "dag":
        assert a.optimizer.global_step == b.optimizer.global_step and abs(a.optimizer.s_t - b.optimizer.s_t) < 1e-9
    env.close(); env_b.close()


def test_time_limit_bootstrap_is_deferred_without_changing_rewards():
    """on_policy_algorithm.py:236-245 adds gamma * V(terminal_observation) to the reward of a truncated env inside the
    step loop. AMPPO parks the terminal observations and values them after the last step (no host synchronisation per
    env-step); the stored rewards must equal those of the in-loop form."""
    from f16_jsb_b200 import F16BatchedEnv, _lib
    from f16_jsb_b200.amppo import AMPPO, AMPPOConfig

    class InLoop(AMPPO):
        def _note_truncations(self, step, truncated):
            self._pending = None
            tr = truncated.nonzero().flatten()
            if tr.numel():
                self._pending = (tr, self.cfg.gamma * self.policy.predict_values(self.env.terminal_obs.index_select(0, tr)))
                self._hits = getattr(self, "_hits", 0) + int(tr.numel())

        def _bootstrap_truncations(self):
            pass

    results = []
    for cls in (AMPPO, InLoop):
        torch.manual_seed(11)
        env = F16BatchedEnv(64, mode="fp32", seed=6)
        algo = cls(env, AMPPOConfig(n_steps=10, batch_size=320, n_epochs=1, seed=9, cuda_graph=False))
        algo._obs = env.reset()
        algo._episode_starts = torch.ones(64, dtype=torch.uint8, device=env.device)
        for i in range(12):                      # envs 0..11 reach step 1200 during this rollout
            _lib.check(env.lib.f16_set_env_step(env._h, i, 1191 + (i % 9)), "f16_set_env_step")
        if cls is InLoop:
            orig_add = algo.buffer.add

            def add(obs, actions, rewards, es, values, log_probs, _orig=orig_add, _a=algo):
                rewards = rewards.clone()
                if _a._pending is not None:
                    rewards[_a._pending[0]] += _a._pending[1]
                _orig(obs, actions, rewards, es, values, log_probs)
            algo.buffer.add = add
        torch.manual_seed(12)
        algo.collect_rollouts()
        results.append((algo.buffer.rewards.clone(), algo.buffer.returns.clone(), getattr(algo, "_hits", None)))
        env.close()
    assert results[1][2] == 12
    # the value head sees the same terminal observations in batches of different shape (one per step against one per
    # rollout): float32 GEMM rounding, a few ulp of rewards of magnitude ~10
    assert torch.allclose(results[0][0], results[1][0], rtol=0, atol=5e-6), float((results[0][0] - results[1][0]).abs().max())
    assert float((results[0][0] - results[1][0]).abs().max()) < 5e-6 and float(results[0][0].abs().max()) > 0.5
    assert torch.allclose(results[0][1], results[1][1], rtol=0, atol=1e-5)


def test_learn_two_iterations_with_ring_layout():
    from f16_jsb_b200 import F16BatchedEnv
    from f16_jsb_b200.amppo import AMPPO, AMPPOConfig
    env = F16BatchedEnv(128, mode="fp32", seed=3, obs_layout="ring")
    algo = AMPPO(env, AMPPOConfig(n_steps=16, batch_size=512, n_epochs=1))
    algo.learn(2 * 16 * 128)
    assert algo.num_timesteps == 2 * 16 * 128 and algo.n_updates == 2
    assert np.isfinite(algo.last_stats["value_loss"])
    env.close()


@pytest.mark.parametrize("rows,bias", [(5, True), (37, True), (4096 * 5, True), (131072 * 5, False)])
def test_layernorm32_kernels_match_torch(rows, bias):
    """csrc/f16_lma_norm.cu against F.layer_norm (the reference's LayerNorm, jsbsim_gym/LMA_features.py:172-185),
    forward and all three gradients; ragged row counts exercise the tile tail."""
    import torch.nn.functional as F

    from f16_jsb_b200.lma import layer_norm32
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn((rows, 32), generator=g, device="cuda") * 3 + 0.7).requires_grad_(True)
    w = (1 + 0.3 * torch.randn(32, generator=g, device="cuda")).requires_grad_(True)
    b = (0.2 * torch.randn(32, generator=g, device="cuda")).requires_grad_(True) if bias else None
    dy = torch.randn((rows, 32), generator=g, device="cuda")
    y = layer_norm32(x, w, b)
    y.backward(dy)
    got = [y.detach(), x.grad.clone(), w.grad.clone()] + ([b.grad.clone()] if bias else [])
    x.grad = w.grad = None
    if bias:
        b.grad = None
    y_ref = F.layer_norm(x.double(), (32,), w.double(), b.double() if bias else None, 1e-5)
    y_ref.backward(dy.double())
    want = [y_ref.detach(), x.grad, w.grad] + ([b.grad] if bias else [])
    for a, r, tol in zip(got, want, (2e-6, 1e-5, 1e-5, 1e-5)):
        scale = float(r.abs().max())
        assert torch.allclose(a.double(), r.double(), rtol=0, atol=tol * max(1.0, scale)), float((a.double() - r.double()).abs().max())


@pytest.mark.parametrize("rows,fin,fout,bias", [(4096, 17, 64, True), (70001, 17, 64, True), (1310720, 17, 64, True), (5 * 4099, 128, 32, True), (131072, 32, 96, True), (65536, 32, 128, False),
                                                 (40000, 160, 64, True), (40000, 160, 128, True), (50000, 64, 4, True), (50001, 64, 1, True)])
def test_linear_weight_gradient_kernel_matches_torch(rows, fin, fout, bias):
    """csrc/f16_lma_wgrad.cu / f16_lma_wgrad_tc.cu through the autograd wiring of lma.Linear, against a float64 torch
    reference: every layer shape of the policy (train.py:21-32,84), ragged row counts; the 17-feature embedding at
    >= 65 536 rows goes to the tensor-core kernel through a zero-padded 32-wide copy."""
    from f16_jsb_b200.lma import Linear
    g = torch.Generator(device="cuda").manual_seed(5)
    lin = Linear(fin, fout, bias=bias).cuda()
    x = torch.randn((rows, fin), generator=g, device="cuda", requires_grad=True)
    dy = torch.randn((rows, fout), generator=g, device="cuda")
    y = lin(x)
    assert y.grad_fn is not None and "LinearFn" in type(y.grad_fn).__name__
    y.backward(dy)
    got = [x.grad.clone(), lin.weight.grad.clone()] + ([lin.bias.grad.clone()] if bias else [])
    xd = x.detach().double().requires_grad_(True)
    wd = lin.weight.detach().double().requires_grad_(True)
    bd = lin.bias.detach().double().requires_grad_(True) if bias else None
    torch.nn.functional.linear(xd, wd, bd).backward(dy.double())
    want = [xd.grad, wd.grad] + ([bd.grad] if bias else [])
    for a, r in zip(got, want):
        # sums of `rows` products of unit-variance numbers in float32: error ~ sqrt(rows) * 6e-8 * sqrt(rows)
        tol = 2e-6 * max(1.0, float(r.abs().max()))
        assert torch.allclose(a.double(), r, rtol=0, atol=tol), (float((a.double() - r).abs().max()), tol)


@pytest.mark.parametrize("shape,p", [((3, 10, 64), 0.0), ((4099, 10, 64), 0.1), ((70000, 5, 32), 0.1), ((1000, 5, 32), 0.5)])
def test_fused_dropout_kernels_match_torch(shape, p):
    """csrc/f16_lma_elementwise.cu through its autograd wiring (lma.dropout_add, lma._EmbedActFn) against plain torch
    float32 ops of the same step (jsbsim_gym/LMA_features.py:221-279 embedding activation, :386-407 residual dropout).
    The keep mask is the kernels' own (read off a forward of ones); with it the results must agree to one rounding, the
    backward must regenerate the same mask, and the keep rate must be 1 - p."""
    from f16_jsb_b200.lma import _DropoutAddFn, _EmbedActFn
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(shape, generator=g, device="cuda").requires_grad_(True)
    z = torch.randn(shape, generator=g, device="cuda").requires_grad_(True)
    dy = torch.randn(shape, generator=g, device="cuda")
    scale = 1.0 / (1.0 - round(p * 65536) / 65536)

    import ctypes as C

    from f16_jsb_b200 import _lib
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ones, zeros = torch.ones(shape, device="cuda"), torch.zeros(shape, device="cuda")

    def mask_of(seed):                     # keep factors (0 or 1 / (1 - p')) of a seed, read off 0 + keep * 1
        out = torch.empty_like(ones)
        _lib.check(_lib.load().f16_lma_dropout_add_forward(ones.numel(), C.c_void_p(ones.data_ptr()), C.c_void_p(zeros.data_ptr()),
                                                           C.c_void_p(out.data_ptr()), float(p), seed, stream), "mask probe")
        assert set(out.unique().tolist()) <= {0.0, float(np.float32(scale))}
        return out

    # residual dropout: y = z + keep * x
    y = _DropoutAddFn.apply(x, z, p)
    seed = y.grad_fn.meta[1]
    m = mask_of(seed)
    # the documented mask function (include/f16_lma.h), restated in NumPy and pinned to Random123's Philox known answers on the CPU
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dropout_mask_oracle
    assert np.array_equal(m.cpu().numpy().reshape(-1), dropout_mask_oracle.keep_factors(x.numel(), p, seed))
    # the kernel's multiply-add is one FMA, torch's two rounded operations: equal to one rounding of the product
    assert torch.allclose(y.detach(), z.detach() + x.detach() * m, rtol=2e-7, atol=2e-7)
    if p == 0:
        assert bool((m == 1).all())
    elif x.numel() > 100000:
        assert abs(float((m > 0).float().mean()) - (1 - p)) < 0.005
        assert not torch.equal(m, mask_of(seed + 1))
    gx, gz = torch.autograd.grad(y, (x, z), dy)
    assert torch.equal(gx, dy * m) and torch.equal(gz, dy)
    # embedding activation: y = keep * (relu(a) + pos[t]), da = keep * dy * (a > 0)
    t, ch = shape[1], shape[2]
    pos = torch.randn((t, ch), generator=g, device="cuda") + 3.0        # > 0 almost surely is not needed: the mask is read off ones
    a = x.detach().clone().requires_grad_(True)
    ye = _EmbedActFn.apply(a, pos, p)
    seed_e = ye.grad_fn.meta[1]
    probe = mask_of(seed_e)
    assert torch.equal(ye.detach(), (torch.relu(a.detach()) + pos) * probe)       # both kernels key the mask by (seed, element / 8)
    (ga,) = torch.autograd.grad(ye, a, dy)
    assert torch.equal(ga, torch.where(a.detach() > 0, dy * probe, torch.zeros_like(dy)))
    # head-stacked output (the permutation of LMA_features.py:255-270 folded into the store / the backward's load)
    for heads in (2, 4):
        if (ch // heads) % 8:
            continue
        stack = lambda v: v.view(shape[0], t, heads, ch // heads).permute(0, 2, 1, 3).reshape(shape[0], t * ch)      # noqa: E731
        a2 = x.detach().clone().requires_grad_(True)
        ys = _EmbedActFn.apply(a2, pos, p, heads)
        assert ys.shape == (shape[0], t * ch)
        m2 = mask_of(ys.grad_fn.meta[1])
        assert torch.equal(ys.detach(), stack((torch.relu(a2.detach()) + pos) * m2))
        dys = stack(dy).contiguous()
        (ga2,) = torch.autograd.grad(ys, a2, dys)
        assert torch.equal(ga2, torch.where(a2.detach() > 0, dy * m2, torch.zeros_like(dy)))


def test_block_with_fused_dropout_equals_unfused_at_p0_and_trains():
    """The extractor wired through the fused paths: at p = 0 both settings of the switch take the torch ops (dropout is
    the identity; outputs equal), at p = 0.1 the fused kernels run (their autograd nodes are in the graph), the outputs
    differ from eval mode and every gradient is finite."""
    from f16_jsb_b200.lma import LMAConfig, LMAExtractor, _LinearFn
    torch.manual_seed(0)
    obs = torch.randn((257, 10, 15), device="cuda")
    res = {}
    for fused in (True, False):
        _LinearFn.use_fused_elementwise = fused
        try:
            torch.manual_seed(1)
            net = LMAExtractor(LMAConfig(dropout=0.0)).cuda().train()
            out = net(obs)
            out.square().sum().backward()
            res[fused] = (out.detach().clone(), [q.grad.clone() for q in net.parameters() if q.grad is not None])
        finally:
            _LinearFn.use_fused_elementwise = True
    assert torch.equal(res[True][0], res[False][0])
    for ga, gb in zip(res[True][1], res[False][1]):
        assert torch.allclose(ga, gb, rtol=1e-5, atol=1e-6 * max(1.0, float(gb.abs().max())))
    net = LMAExtractor(LMAConfig(dropout=0.1)).cuda().train()
    out = net(obs)
    names, todo = set(), [out.grad_fn]
    while todo:
        f = todo.pop()
        if f is not None and f not in names:
            names.add(f)
            todo.extend(n for n, _ in f.next_functions)
    names = {type(f).__name__ for f in names}
    assert any("DropoutAddFn" in n for n in names) and any("EmbedActFn" in n for n in names), names
    out.square().sum().backward()
    assert all(bool(torch.isfinite(q.grad).all()) for q in net.parameters() if q.grad is not None)
    assert not torch.equal(out.detach(), net.eval()(obs).detach())
