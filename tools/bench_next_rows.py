#!/usr/bin/env python3
"""Measures the kernels of SURVEY 8(f) rows 1-2 (rollout store, GAE, gather, feature transform) with
CUDA events and reports achieved algorithmic GB/s against the measured HBM copy bandwidth."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from f16_jsb_b200 import F16BatchedEnv  # noqa: E402
from f16_jsb_b200.features import jsbsim_features  # noqa: E402
from f16_jsb_b200.rollout import GpuRolloutBuffer  # noqa: E402


def timed(fn, iters):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def main():
    peak = 6478.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    N, T = 65536, 256
    env = F16BatchedEnv(N, mode="fp32")
    obs = env.reset()
    buf = GpuRolloutBuffer(T, N, device=env.device, gae_lambda=0.95, gamma=0.99)
    act = torch.rand((N, 4), device="cuda")
    rew, val, lp = (torch.randn(N, device="cuda") for _ in range(3))
    es = torch.zeros(N, dtype=torch.uint8, device="cuda")
    out = {}
    for t in range(T):
        buf.add(obs, act, rew, es, val, lp)
    buf.pos = 1
    t_add = timed(lambda: (buf.add(obs, act, rew, es, val, lp), setattr(buf, "pos", 1)), 200)
    out["rollout_add"] = dict(s=t_add, bytes=N * (60 * 2 + 16 * 2 + 4 * 3 * 2 + 1 + 1 + 4), per="step of 65536 envs")
    buf.full = True
    t_gae = timed(lambda: buf.compute_returns_and_advantage(val, es), 50)
    out["rollout_gae"] = dict(s=t_gae, bytes=N * T * (4 * 3 + 4 * 2), per="65536 envs x 256 steps")
    idx = torch.randperm(N * T, device="cuda")[: 1 << 20]
    t_gat = timed(lambda: buf.gather(idx), 50)
    out["rollout_gather"] = dict(s=t_gat, bytes=(1 << 20) * (600 * 2 + 8 + 16 * 2 + 4 * 2 * 4 + 1), per="1M-sample minibatch (incl. torch.empty of the outputs)")
    big = torch.randn((1 << 20, 10, 15), device="cuda")
    t_feat = timed(lambda: jsbsim_features(big), 50)
    out["features17"] = dict(s=t_feat, bytes=(10 << 20) * (60 + 68), per="1M stacked observations = 10M frames")
    for k, v in out.items():
        v["GBps"] = v["bytes"] / v["s"] / 1e9
        v["frac_of_measured_hbm"] = v["GBps"] / peak
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
