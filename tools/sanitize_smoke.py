#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel for `compute-sanitizer --tool memcheck` (ragged batch sizes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f16_jsb_b200 import F16BatchedEnv  # noqa: E402
from f16_jsb_b200.features import jsbsim_features  # noqa: E402
from f16_jsb_b200.rollout import GpuRolloutBuffer  # noqa: E402

for mode, n in (("fp32", 1000 + 37), ("fp64", 77)):
    env = F16BatchedEnv(n, mode=mode, seed=1)
    obs = env.reset()
    buf = GpuRolloutBuffer(6, n, device=env.device, gae_lambda=0.95, gamma=0.99)
    es = torch.ones(n, dtype=torch.bool, device="cuda")
    for t in range(6):
        last = obs.clone()
        a = torch.rand((n, 4), device="cuda")
        o, r, d, tr = env.step(a if t % 2 else None, auto_reset=True)
        buf.add(last, a, r, es, r, r)
        es = d.bool()
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    mask[::3] = 1
    env.reset(mask=mask)
    buf.compute_returns_and_advantage(r, es)
    s = buf.gather(torch.randperm(6 * n, device="cuda")[:500])
    f = jsbsim_features(s.observations)
    st = env.pack_states()
    env.unpack_states(st)
    env.get_state(3)
    torch.cuda.synchronize()
    print(mode, n, float(f.abs().sum()), env.stats())
print("sanitize smoke done")
