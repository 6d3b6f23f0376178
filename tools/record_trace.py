#!/usr/bin/env python3
"""Record one episode of the REFERENCE env into an .f16trace.npz (format: f16_jsb_b200/trace.py).

    python tools/record_trace.py --backend jsbsim --reference /path/to/F16_JSB --seed 0 --actions random \
        --out tests/golden/jsbsim_random0.f16trace.npz

--backend jsbsim   imports the real `jsbsim` and `gymnasium` packages and the reference's own
                   jsbsim_gym/jsbsim_gym.py from --reference: the file it writes PINS PARITY AGAINST REAL JSBSIM
                   (tests/test_traces.py picks up every tests/golden/*.f16trace.npz).
--backend oracle   same driver, but `jsbsim`, `gymnasium`, `pygame`, `moderngl` come from oracle/refshim (the CPU
                   restatement): what can be produced where JSBSim is not installable; pins the format and the
                   env layer only.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_actions(kind, seed, n):
    low, high = np.array([-1, -1, -1, 0], np.float32), np.array([1, 1, 1, 1], np.float32)
    rng = np.random.default_rng(seed)
    if kind == "random":
        return rng.uniform(low, high, size=(n, 4)).astype(np.float32)
    if kind == "gentle":
        a = np.stack([0.2 * rng.standard_normal(n), -0.1 + 0.2 * rng.standard_normal(n), 0.2 * rng.standard_normal(n),
                      0.6 + 0.2 * rng.standard_normal(n)], axis=1)
        return np.clip(a, low, high).astype(np.float32)
    if kind == "dive":     # push over and hold: a steep, fast impact (exercises the ground reactions of the last env-step)
        a = np.zeros((n, 4), np.float32)
        a[:, 1] = 0.5 + 0.5 * rng.random()
        a[:, 3] = 1.0
        a[:, 0] = (0.2 + 0.6 * rng.random()) * np.sin(np.arange(n) / (8.0 + 20.0 * rng.random()))
        a[:, 2] = 0.3 * rng.standard_normal()
        return a
    raise ValueError(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["jsbsim", "oracle"], default="oracle")
    ap.add_argument("--reference", default=os.environ.get("F16_REFERENCE", "/root/reference"))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--actions", choices=["random", "gentle", "dive"], default="random")
    ap.add_argument("--action-seed", type=int, default=None)
    ap.add_argument("--max-steps", type=int, default=1300)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.backend == "oracle":
        sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim")]
    sys.path[:0] = [a.reference, ROOT]
    import gymnasium as gym
    import jsbsim
    import jsbsim_gym.jsbsim_gym  # noqa: F401  (the reference's file: registers JSBSim-v0)
    from f16_jsb_b200.trace import record_episode, save_trace

    env = gym.make("JSBSim-v0", root=a.reference)
    fdm = env.unwrapped.simulation
    producer = "jsbsim %s" % getattr(jsbsim, "__version__", "?") if a.backend == "jsbsim" else "cpu-restatement"
    actions = make_actions(a.actions, 1000 + a.seed if a.action_seed is None else a.action_seed, a.max_steps)
    rec = record_episode(env, fdm, a.seed, actions)
    save_trace(a.out, producer=producer, notes="actions=%s" % a.actions, **rec)
    print("wrote %s: %d steps, terminated=%s truncated=%s, producer=%s" % (
        a.out, len(rec["frames"]), bool(rec["terminated"][-1]), bool(rec["truncated"][-1]), producer))


if __name__ == "__main__":
    main()
