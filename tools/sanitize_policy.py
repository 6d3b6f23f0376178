#!/usr/bin/env python3
"""Small runs of the fused policy forward and the elementwise dropout kernels for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_policy.py
    compute-sanitizer --tool racecheck python tools/sanitize_policy.py
(compute-sanitizer is closed on the shared B200 pool; run plainly the script is the small-case runner: it prints the worst
difference against the torch modules.) Ragged tiles (37 and 300 envs: partial CTAs, several tiles per CTA are covered by the grid-stride loop at 5 000)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from f16_jsb_b200.constants import ACTION_HIGH, ACTION_LOW  # noqa: E402
from f16_jsb_b200.lma import LMAActorCritic, PolicyForwardKernel, _DropoutAddFn, _EmbedActFn  # noqa: E402

torch.manual_seed(0)
net = LMAActorCritic().cuda().eval()
fused = PolicyForwardKernel(net, torch.as_tensor(ACTION_LOW).cuda(), torch.as_tensor(ACTION_HIGH).cuda())
worst = 0.0
for n in (37, 300, 5000):
    obs = torch.randn((n, 10, 15), device="cuda")
    noise = torch.randn((n, 4), device="cuda")
    actions, values, log_probs, clipped, feats = fused(obs, noise, features=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        worst = max(worst, float((feats - net.features_extractor(obs)).abs().max()), float((values - net.predict_values(obs)).abs().max()))
x = torch.randn((301, 10, 64), device="cuda", requires_grad=True)
pos = torch.randn((10, 64), device="cuda")
for heads in (1, 4):
    y = _EmbedActFn.apply(x, pos, 0.1, heads)
    y.sum().backward()
z = torch.randn((301, 5, 32), device="cuda", requires_grad=True)
w = torch.randn((301, 5, 32), device="cuda", requires_grad=True)
_DropoutAddFn.apply(z, w, 0.1).sum().backward()
torch.cuda.synchronize()
print("sanitize_policy: ok, worst abs difference against the modules %.3e" % worst)
