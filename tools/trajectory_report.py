#!/usr/bin/env python3
"""Trajectory-error report (BASELINE metric: "...; trajectory error"), configs[1]/[2] sizes.

4 096 envs, seeds 0..4095 for the goals (default_rng(seed) draws of JSBSimEnv.reset), 1 000 env-steps,
host-generated actions, no auto-reset: the CUDA env in FP64 and FP32 mode against the oracle's batch
trajectories. Reports, per checkpoint step, the median / p99 / max over the still-flying envs of the
relative error of the 12 observed quantities (|x - ref| / max(|ref|, 1e-2)), the position error in
metres, and how many envs end their episode at the same step as the oracle.
Two action distributions: "uniform" = action_space.sample()-like (aggressive, most envs crash within
~700 steps) and "gentle" (small stick inputs around a pitch-up bias, most envs survive).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from f16_jsb_b200 import F16BatchedEnv  # noqa: E402
from oracle import f16_oracle  # noqa: E402

N, T = 4096, 1000
CHECK = (1, 10, 30, 100, 300, 600, 1000)
LOW, HIGH = np.array([-1, -1, -1, 0], np.float32), np.array([1, 1, 1, 1], np.float32)


def actions_for(kind, rng):
    if kind == "uniform":
        return rng.uniform(LOW, HIGH, size=(T, N, 4)).astype(np.float32)
    a = np.stack([0.2 * rng.standard_normal((T, N)), -0.1 + 0.2 * rng.standard_normal((T, N)),
                  0.2 * rng.standard_normal((T, N)), 0.6 + 0.2 * rng.standard_normal((T, N))], axis=-1)
    return np.clip(a, LOW, HIGH).astype(np.float32)


def run(kind):
    rng = np.random.default_rng(7)
    goals = np.stack([f16_oracle.sample_goal(s) for s in range(N)])
    actions = actions_for(kind, rng)
    frames, rewards, flags = f16_oracle.batch_trajectory(goals, actions)
    ref_done_step = np.where((flags & 3).any(0), (flags & 3 != 0).argmax(0), T)
    out = {"envs": N, "steps": T, "oracle_episodes_finished": int((ref_done_step < T).sum())}
    for mode in ("fp64", "fp32"):
        env = F16BatchedEnv(N, mode=mode)
        env.reset(goals=torch.from_numpy(goals).cuda())
        a_dev = torch.from_numpy(actions).cuda()
        alive = np.ones(N, bool)
        done_step = np.full(N, T)
        rows = {}
        for k in range(T):
            obs, rew, done, trunc = env.step(a_dev[k], auto_reset=False)
            d = done.cpu().numpy().astype(bool)
            newly = alive & d
            done_step[newly] = k
            if (k + 1) in CHECK:
                fr = obs[:, -1, :12].cpu().numpy()
                both = alive & (ref_done_step > k)
                if both.sum() == 0:
                    break
                ref = frames[k][both, :12]
                e = (np.abs(fr[both] - ref) / np.maximum(np.abs(ref), 1e-2)).max(1)
                pos = np.abs(fr[both, :3] - ref[:, :3]).max(1)
                rows[str(k + 1)] = {"envs_flying": int(both.sum()), "rel_err_median": float(np.median(e)), "rel_err_p99": float(np.percentile(e, 99)),
                                    "rel_err_max": float(e.max()), "pos_err_m_median": float(np.median(pos)), "pos_err_m_p99": float(np.percentile(pos, 99)),
                                    "pos_err_m_max": float(pos.max()), "bit_identical_frames": int((e == 0).sum())}
            alive &= ~d
        same = int((done_step == ref_done_step).sum())
        within1 = int((np.abs(done_step - ref_done_step) <= 1).sum())
        out[mode] = {"checkpoints": rows, "episodes_ending_at_same_step": same, "episodes_ending_within_one_step": within1}
        env.close()
    return out


if __name__ == "__main__":
    rep = {kind: run(kind) for kind in ("uniform", "gentle")}
    print(json.dumps(rep, indent=1))
