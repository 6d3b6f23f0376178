#!/usr/bin/env python3
"""Generate tests/golden/ref_env_traces.npz by running the REFERENCE's own env file.

/root/reference/jsbsim_gym/jsbsim_gym.py is imported unmodified (oracle/refshim supplies stub
`jsbsim`, `gymnasium`, `pygame`, `moderngl` modules); its Python - float32 cast chain, angle wrap,
frame stack, reward, termination, PositionReward, TimeLimit(1200) - runs verbatim. The FDM underneath
is the CPU oracle (oracle/f16_oracle.cpp), because real JSBSim is not installable here, so these
vectors pin the env layer against the reference's code and the FDM against the restatement
("parity unpinned" w.r.t. real JSBSim).

Only runs in the build container (needs /root/reference). Usage: python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("F16_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), REF, ROOT]

import gymnasium as gym  # noqa: E402  (refshim)
import jsbsim_gym.jsbsim_gym  # noqa: E402,F401  (the reference's file: registers JSBSim-v0)

LOW = np.array([-1, -1, -1, 0], dtype=np.float32)
HIGH = np.array([1, 1, 1, 1], dtype=np.float32)


def actions_random(seed, n):
    rng = np.random.default_rng(1000 + seed)
    return rng.uniform(LOW, HIGH, size=(n, 4)).astype(np.float32)


def actions_gentle(seed, n, scale, elev, thr):
    rng = np.random.default_rng(2000 + seed)
    a = np.stack([scale * rng.standard_normal(n), elev + scale * rng.standard_normal(n),
                  scale * rng.standard_normal(n), thr + 0.2 * rng.standard_normal(n)], axis=1)
    return np.clip(a, LOW, HIGH).astype(np.float32)


def run(seed, actions, with_states):
    env = gym.make("JSBSim-v0", root=REF)
    obs, _ = env.reset(seed=seed)
    fdm = env.unwrapped.simulation._fdm
    rec = dict(goal=env.unwrapped.goal.copy(), reset_obs=obs.copy(), frames=[], rewards=[], terminated=[], truncated=[],
               stacked=[], states=[fdm.pack_state()] if with_states else None)
    for k, a in enumerate(actions):
        obs, r, term, trunc, _ = env.step(a)
        rec["frames"].append(obs[-1].copy())
        rec["rewards"].append(np.float32(r))
        rec["terminated"].append(term)
        rec["truncated"].append(trunc)
        if k in (0, 4, 11):
            rec["stacked"].append(obs.copy())
        if with_states:
            rec["states"].append(fdm.pack_state())
        if term or trunc:
            break
    n = len(rec["frames"])
    out = dict(actions=actions[:n], goal=rec["goal"], reset_obs=rec["reset_obs"], frames=np.stack(rec["frames"]),
               rewards=np.array(rec["rewards"], dtype=np.float32), terminated=np.array(rec["terminated"]),
               truncated=np.array(rec["truncated"]), stacked=np.stack(rec["stacked"]))
    if with_states:
        out["states"] = np.stack(rec["states"])
    return out


def run_multi(seeds, action_fn, max_steps=1300):
    """Several episodes on ONE env object, as train.py's DummyVecEnv drives it: reset() on a used env is run_ic() +
    set-running on top of whatever the previous episode left in the FCS / Auxiliary / Accelerations models
    (jsbsim_gym.py:305-306). The packed FDM state is recorded after every reset and every step."""
    env = gym.make("JSBSim-v0", root=REF)
    fdm = env.unwrapped.simulation._fdm
    out = {}
    for ep, seed in enumerate(seeds):
        obs, _ = env.reset(seed=seed)
        actions = action_fn(ep, max_steps)
        frames, rewards, term_l, trunc_l, states = [], [], [], [], [fdm.pack_state()]
        for a in actions:
            o, r, term, trunc, _ = env.step(a)
            frames.append(o[-1].copy()); rewards.append(np.float32(r)); term_l.append(term); trunc_l.append(trunc)
            states.append(fdm.pack_state())
            if term or trunc:
                break
        n = len(frames)
        pre = "ep%d/" % ep
        out[pre + "goal"] = env.unwrapped.goal.copy()
        out[pre + "reset_obs"] = obs.copy()
        out[pre + "actions"] = actions[:n]
        out[pre + "frames"] = np.stack(frames)
        out[pre + "rewards"] = np.array(rewards, dtype=np.float32)
        out[pre + "terminated"] = np.array(term_l)
        out[pre + "truncated"] = np.array(trunc_l)
        out[pre + "states"] = np.stack(states)
        print("one env object, episode", ep, "steps", n, "terminated", bool(term_l[-1]), "truncated", bool(trunc_l[-1]))
    return out


def multi_actions(ep, n):
    """Episode 0 random, 1 a hard dive (ends in a crash within ~250 steps), 2 gentle, 3 random."""
    if ep == 1:
        rng = np.random.default_rng(3100)
        a = np.stack([0.3 * rng.standard_normal(n), 0.9 + 0.05 * rng.standard_normal(n), 0.1 * rng.standard_normal(n),
                      0.8 + 0.1 * rng.standard_normal(n)], axis=1)
        return np.clip(a, LOW, HIGH).astype(np.float32)
    if ep == 2:
        return actions_gentle(7, n, 0.2, -0.1, 0.6)
    return actions_random(40 + ep, n)


def main():
    if "--multi" in sys.argv:
        flat = run_multi([100, 101, 102, 103], multi_actions)
        flat["numpy_version"] = np.array(np.__version__)
        path = os.path.join(ROOT, "tests", "golden", "ref_env_one_object_4_episodes.npz")
        np.savez_compressed(path, **flat)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")
        return
    traces = {}
    for seed in range(4):
        traces["random%d" % seed] = run(seed, actions_random(seed, 1300), with_states=(seed == 0))
    traces["gentle0"] = run(10, actions_gentle(0, 1300, 0.2, -0.1, 0.6), with_states=True)
    traces["gentle1"] = run(11, actions_gentle(1, 1300, 0.3, -0.1, 0.7), with_states=False)
    flat = {}
    for name, t in traces.items():
        print(name, "steps", len(t["frames"]), "terminated", bool(t["terminated"][-1]), "truncated", bool(t["truncated"][-1]),
              "return %.4f" % float(t["rewards"].sum()))
        for k, v in t.items():
            flat["%s/%s" % (name, k)] = v
    flat["numpy_version"] = np.array(np.__version__)
    path = os.path.join(ROOT, "tests", "golden", "ref_env_traces.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
