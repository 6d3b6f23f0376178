#!/usr/bin/env python3
"""Exercise the ground-reaction builds (near-ground tiles first, cold redo, carry-over reset) under compute-sanitizer:
a ragged batch of envs diving into the ground with auto-reset, both precisions, ring and stacked layouts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f16_jsb_b200 import F16BatchedEnv  # noqa: E402

for mode, n, layout, reset in (("fp32", 2000 + 37, "ring", "carryover"), ("fp64", 300 + 5, "ring", "snapshot"), ("fp32", 999, "stacked", "snapshot")):
    env = F16BatchedEnv(n, mode=mode, seed=1, obs_layout=layout, ground_reactions=True, reset_mode=reset)
    env.reset()
    a = torch.zeros((n, 4), device="cuda")
    a[:, 1] = 0.9 * (torch.rand(n, device="cuda") > 0.3).float()      # most envs push over into the ground
    a[:, 0] = torch.rand(n, device="cuda") - 0.5
    a[:, 3] = 1.0
    for t in range(int(os.environ.get("STEPS", "420"))):
        env.step(a, auto_reset=True)
    torch.cuda.synchronize()
    st = env.stats()
    print(mode, n, layout, reset, {k: st[k] for k in ("episodes", "crashes", "ground_redos")})
    assert st["crashes"] > 0
    env.close()
print("sanitize ground done")
