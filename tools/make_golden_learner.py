#!/usr/bin/env python3
"""Golden vectors for the learner-side rows of SURVEY.md 8(f) (2: LMA feature extractor, 3: AM-PPO advantage
modulation and the DAG optimizer), produced by executing the REFERENCE's own Python from /root/reference in
this container (it cannot travel to the GPU box; the vectors can):

* jsbsim_gym/LMA_features.py + jsbsim_gym/features.py  (StackedLMAFeaturesExtractor, train.py:21-32 kwargs),
  imported unmodified with two stand-ins on sys.modules: `gymnasium.spaces` (oracle/refshim) and
  `stable_baselines3.common.torch_layers.BaseFeaturesExtractor` (the vendored SB3 package itself needs the real
  gymnasium to import; the stand-in is the 10-line nn.Module base class of common/torch_layers.py:14-32);
* stable_baselines3/ppo/ppo.py: the function `dynago_transform_advantages` (:29-99), extracted from the file's
  AST and executed as is (importing the module would pull the whole SB3 package);
* stable_baselines3/ppo/optim/sgd.py: class DAG (:87-344), same extraction, on top of torch.optim.Optimizer.

Writes tests/golden/learner_golden.pt. Usage: python tools/make_golden_learner.py
"""
import ast
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))     # gymnasium stand-in


def _stub_sb3():
    import torch.nn as nn

    class BaseFeaturesExtractor(nn.Module):
        def __init__(self, observation_space, features_dim: int = 0):
            super().__init__()
            assert features_dim > 0
            self._observation_space = observation_space
            self._features_dim = features_dim

        @property
        def features_dim(self):
            return self._features_dim

    for name in ("stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.torch_layers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["stable_baselines3.common.torch_layers"].BaseFeaturesExtractor = BaseFeaturesExtractor


def _load_ref_module(qualname, path, package):
    spec = importlib.util.spec_from_file_location(qualname, path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = package
    sys.modules[qualname] = mod
    spec.loader.exec_module(mod)
    return mod


def _extract(path, names, namespace):
    """exec the named top-level defs/classes of a reference file, unmodified, in `namespace`."""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, namespace)
    return namespace


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    out = {}

    # ---------------------------------------------------------------- LMA extractor
    _stub_sb3()
    pkg = types.ModuleType("jsbsim_gym")
    pkg.__path__ = [os.path.join(REF, "jsbsim_gym")]
    sys.modules["jsbsim_gym"] = pkg
    _load_ref_module("jsbsim_gym.features", os.path.join(REF, "jsbsim_gym", "features.py"), "jsbsim_gym")
    lma = _load_ref_module("jsbsim_gym.LMA_features", os.path.join(REF, "jsbsim_gym", "LMA_features.py"), "jsbsim_gym")
    from gymnasium import spaces
    low = np.full((10, 15), -np.inf, np.float32)
    obs_space = spaces.Box(low=low, high=-low, shape=(10, 15), dtype=np.float32)
    kw = dict(lma_embed_dim_d0=64, lma_num_heads_stacking=4, lma_num_heads_latent=4, lma_ff_latent_hidden=128,
              lma_num_layers=2, lma_dropout=0.1, lma_bias=True)                      # train.py:21-32
    net = lma.StackedLMAFeaturesExtractor(obs_space, **kw).eval()
    # observations shaped like the env's: positions ~1e3 m, altitude, mach, small angles/rates, goal
    g = torch.Generator().manual_seed(1)
    B = 24
    frame = torch.zeros(B, 10, 15)
    frame[..., 0:2] = torch.randn(B, 10, 2, generator=g) * 3000
    frame[..., 2] = 1500 + torch.randn(B, 10, generator=g) * 500
    frame[..., 3] = 0.8 + torch.randn(B, 10, generator=g) * 0.05
    frame[..., 4:9] = torch.randn(B, 10, 5, generator=g) * 0.1
    frame[..., 9:12] = (torch.rand(B, 10, 3, generator=g) * 2 - 1) * math.pi
    frame[..., 12:15] = torch.tensor([4000.0, -2500.0, 2500.0]) + torch.randn(B, 1, 3, generator=g) * 1000
    with torch.no_grad():
        feats = net(frame)
    out["lma"] = {"kwargs": kw, "state_dict": {k: v.clone() for k, v in net.state_dict().items()}, "obs": frame, "features": feats,
                  "features_dim": int(net.features_dim), "L_new": int(net.lma_extractor.lma_config.L_new),
                  "C_new": int(net.lma_extractor.lma_config.C_new)}
    # gradient of a scalar loss w.r.t. two parameters (training-mode parity without dropout: p = 0)
    net0 = lma.StackedLMAFeaturesExtractor(obs_space, **dict(kw, lma_dropout=0.0))
    net0.load_state_dict(net.state_dict())
    net0.train()
    loss = (net0(frame) ** 2).mean()
    loss.backward()
    out["lma"]["loss"] = loss.detach()
    out["lma"]["grads"] = {k: p.grad.clone() for k, p in net0.named_parameters()
                           if k in ("lma_extractor.initial_transform.input_embedding.weight", "lma_extractor.lma_blocks.1.attn.c_attn.weight",
                                    "lma_extractor.lma_blocks.0.ln_1.weight")}
    out["divisors"] = [(t, d, lma.find_closest_divisor(t, d)) for t, d in ((640, 5), (640, 7), (384, 3), (170, 5), (97, 10), (1000, 33))]

    # ---------------------------------------------------------------- AM-PPO advantage modulation
    ns = {"th": torch, "Dict": dict}
    _extract(os.path.join(REF, "stable_baselines3", "ppo", "ppo.py"), {"dynago_transform_advantages"}, ns)
    fn = ns["dynago_transform_advantages"]
    params = {"kappa": 2.0, "tau": 1.25, "p_star": 0.10, "eta": 0.3, "rho": 0.1, "eps": 1e-5, "alpha_min": 1e-12,
              "alpha_max": 1e12, "rho_sat": 0.98}                                  # train.py:165-176 defaults
    cases = []
    alpha, sat = torch.tensor([1.0]), torch.tensor([0.10])
    for i, (n, scale, update) in enumerate(((4096, 1.0, True), (256, 1.0, False), (256, 5.0, False), (4096, 0.01, True),
                                            (1000, 30.0, True), (2, 1.0, False), (1, 1.0, True), (257, 1.0, True))):
        adv = torch.randn(n, generator=g) * scale + (0.3 * scale if i % 2 else 0.0)
        a0, s0 = alpha.clone(), sat.clone()
        mod = fn(adv.clone(), params, alpha, sat, 2.0, 0.0 if i != 2 else 0.25, update_ema=update)
        cases.append({"adv": adv, "alpha_in": a0, "sat_in": s0, "update": update, "kappa": 2.0, "v_shift": 0.0 if i != 2 else 0.25,
                      "mod": mod.clone(), "alpha_out": alpha.clone(), "sat_out": sat.clone()})
    out["dynago"] = {"params": params, "cases": cases}

    # ---------------------------------------------------------------- DAG optimizer
    from torch.optim.optimizer import Optimizer
    ns = {"torch": torch, "math": math, "Optimizer": Optimizer, "Iterable": object, "Optional": __import__("typing").Optional,
          "Callable": __import__("typing").Callable, "Tensor": torch.Tensor, "List": list}
    _extract(os.path.join(REF, "stable_baselines3", "ppo", "optim", "sgd.py"), {"DAG", "_try_multi_tensor_std"}, ns)
    DAG = ns["DAG"]
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(17, 32), torch.nn.Tanh(), torch.nn.Linear(32, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    init = {k: v.clone() for k, v in model.state_dict().items()}
    opt = DAG(model.parameters(), lr=9e-5, shrink={"warmup_steps": 4, "lambda_rms": 2.0})     # short warm-up, strong shrink: s_t < 1
    xs = torch.randn(12, 64, 17, generator=g)
    ys = torch.randn(12, 64, 1, generator=g)
    traj = []
    for k in range(12):
        opt.zero_grad()
        loss = ((model(xs[k]) - ys[k]) ** 2).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
        traj.append({"loss": loss.detach().clone(), "s_t": float(opt.s_t),
                     "params": {kk: v.clone() for kk, v in model.state_dict().items()}})
    out["dag"] = {"init": init, "xs": xs, "ys": ys, "traj": traj, "lr": 9e-5, "shrink": {"warmup_steps": 4, "lambda_rms": 2.0}, "kappa": float(opt.h["kappa"])}

    path = os.path.join(ROOT, "tests", "golden", "learner_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; LMA features_dim", out["lma"]["features_dim"], "L_new", out["lma"]["L_new"],
          "C_new", out["lma"]["C_new"], "params", sum(v.numel() for v in out["lma"]["state_dict"].values()))


if __name__ == "__main__":
    main()
