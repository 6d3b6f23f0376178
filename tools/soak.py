"""Soak run on a B200: random-action rollouts of several thousand steps, finite-value checks on observations, rewards and
states, episode statistics. usage: python tools/soak_carryover.py [carryover|snapshot]"""
import sys, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
RESET = sys.argv[1] if len(sys.argv) > 1 else 'carryover'
from f16_jsb_b200 import F16BatchedEnv
for mode, n, steps in (("fp32", 65536, 6000), ("fp64", 16384, 2500)):
    env = F16BatchedEnv(n, mode=mode, reset_mode=RESET, ground_reactions=True, seed=3)
    env.reset()
    bad = 0
    for k in range(steps):
        obs, rew, done, trunc = env.step(None, auto_reset=True)
        if k % 500 == 499:
            ok = bool(torch.isfinite(obs).all()) and bool(torch.isfinite(rew).all())
            bad += (not ok)
            st = env.pack_states()
            bad += (not bool(torch.isfinite(st).all()))
    s = env.stats()
    print(RESET, mode, n, steps, "non-finite checks failed:", bad, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in s.items()},
          "mean len %.1f mean ret %.2f" % (s["length_sum"] / max(1, s["episodes"]), s["return_sum"] / max(1, s["episodes"])))
    env.close()
