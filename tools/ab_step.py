#!/usr/bin/env python3
"""Device-resident timing of one library build at steady state (A/B runs: tools/ab.sh).
    F16_B200_LIB=path/to/lib.so python tools/ab_step.py --mode fp32 --layout ring --steps 500
Prints one line: <lib> <mode> <layout> <ground> <reset> ms_per_step value."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from f16_jsb_b200 import F16BatchedEnv

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fp32")
ap.add_argument("--layout", default="ring")
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=500)
ap.add_argument("--preroll", type=int, default=600)
ap.add_argument("--ground", default="default")
ap.add_argument("--reset", default="snapshot")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
env = F16BatchedEnv(a.envs, device=dev, mode=a.mode, obs_layout=a.layout, reset_mode=a.reset,
                    ground_reactions={"default": None, "on": True, "off": False}[a.ground], with_terminal_obs=a.layout != "frame")
gen = torch.Generator(device=dev); gen.manual_seed(1)
lo = torch.tensor([-1, -1, -1, 0], dtype=torch.float32, device=dev); hi = torch.tensor([1, 1, 1, 1], dtype=torch.float32, device=dev)
acts = [lo + (hi - lo) * torch.rand((a.envs, 4), generator=gen, device=dev) for _ in range(8)]
env.reset()
for w in range(a.preroll):
    env.step(acts[w % 8])
torch.cuda.synchronize()
best = []
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        env.step(acts[k % 8])
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / a.steps)
ms = min(best)
print("%s %s %s ground=%s reset=%s ms_per_step %.4f value %.4e (reps %s)" % (
    os.path.basename(os.environ.get("F16_B200_LIB", "main")), a.mode, a.layout, a.ground, a.reset, ms, a.envs / ms * 1e3,
    " ".join("%.4f" % b for b in best)))
