#!/usr/bin/env python3
"""Trajectory-error report (BASELINE metric: "...; trajectory error") at the sizes of BASELINE configs[1] / [2].

N envs, seeds 0..N-1 for the goals (default_rng(seed) draws of JSBSimEnv.reset), T env-steps, host-generated actions,
no auto-reset: the CUDA env against the oracle's batch trajectories. Per checkpoint step: median / p99 / max over the
still-flying envs of the relative error of the 12 observed quantities (|x - ref| / max(|ref|, 1e-2)), the position
error in metres, how many envs end their episode at the same step as the oracle, and the step at which each env's
error first exceeds 1e-3 (its "divergence step": once actuators saturate the airframe is open-loop unstable, e-folding
~0.4 s, so round-off differences grow until a switch - TEF 250 kt, LEF alpha thresholds, Mach 0.9 - flips on one side).
Two action distributions: "uniform" = action_space.sample()-like (aggressive, most envs crash within ~700 steps) and
"gentle" (small stick inputs around a pitch-up bias, most envs survive).

`trajectory_stats` is what tests/test_gpu_full_horizon.py asserts the stated bounds on; run as a script it prints the
full table (profiles/r2_trajectory_error.json).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LOW, HIGH = np.array([-1, -1, -1, 0], np.float32), np.array([1, 1, 1, 1], np.float32)
DIVERGED = 1e-3


def actions_for(kind, rng, t, n):
    if kind == "uniform":
        return rng.uniform(LOW, HIGH, size=(t, n, 4)).astype(np.float32)
    a = np.stack([0.2 * rng.standard_normal((t, n)), -0.1 + 0.2 * rng.standard_normal((t, n)),
                  0.2 * rng.standard_normal((t, n)), 0.6 + 0.2 * rng.standard_normal((t, n))], axis=-1)
    return np.clip(a, LOW, HIGH).astype(np.float32)


def trajectory_stats(mode, kind, n, t, checkpoints, chunk=16384, seed=7):
    """CUDA env (mode 'fp64' / 'fp32') against the oracle on n envs x t steps; envs are processed `chunk` at a time so the
    oracle's (t, chunk, 15) frame array stays small. Returns a dict of per-checkpoint error statistics."""
    import torch

    from f16_jsb_b200 import F16BatchedEnv
    from oracle import f16_oracle

    per_ck = {c: {"e": [], "pos": []} for c in checkpoints}
    done_step_all, ref_done_all, div_step_all = [], [], []
    reward_err = 0.0
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        rng = np.random.default_rng(seed + lo)
        goals = np.stack([f16_oracle.sample_goal(s) for s in range(lo, lo + m)])
        actions = actions_for(kind, rng, t, m)
        frames, rewards, flags = f16_oracle.batch_trajectory(goals, actions)
        ref_done = np.where((flags & 3).any(0), (flags & 3 != 0).argmax(0), t)
        env = F16BatchedEnv(m, mode=mode)
        env.reset(goals=torch.from_numpy(goals).cuda())
        a_dev = torch.from_numpy(actions).cuda()
        newest = torch.empty((t, m, 15), dtype=torch.float32, device="cuda")
        rew_d = torch.empty((t, m), dtype=torch.float32, device="cuda")
        done_d = torch.empty((t, m), dtype=torch.uint8, device="cuda")
        for k in range(t):
            obs, rew, done, trunc = env.step(a_dev[k], auto_reset=False)
            newest[k].copy_(obs[:, -1, :])
            rew_d[k].copy_(rew)
            done_d[k].copy_(done)
        fr = newest.cpu().numpy()
        rw = rew_d.cpu().numpy()
        dn = done_d.cpu().numpy().astype(bool)
        env.close()
        done_step = np.where(dn.any(0), dn.argmax(0), t)
        # error of every frame while both sides are still flying
        steps = np.arange(t)[:, None]
        both = (steps <= np.minimum(done_step, ref_done)[None, :]) & (steps < t)
        e_all = (np.abs(fr[..., :12] - frames[..., :12]) / np.maximum(np.abs(frames[..., :12]), 1e-2)).max(-1)
        e_all = np.where(both, e_all, 0.0)
        big = e_all > DIVERGED
        div_step_all.append(np.where(big.any(0), big.argmax(0), t))
        flying = both & (steps < np.minimum(done_step, ref_done)[None, :])       # strictly before either side's last step
        reward_err = max(reward_err, float(np.abs(np.where(flying & ~big, rw - rewards, 0.0)).max()))
        for c in checkpoints:
            k = c - 1
            sel = (done_step > k) & (ref_done > k)
            if sel.any():
                per_ck[c]["e"].append(e_all[k, sel])
                per_ck[c]["pos"].append(np.abs(fr[k, sel, :3] - frames[k, sel, :3]).max(1))
        done_step_all.append(done_step)
        ref_done_all.append(ref_done)
    done_step, ref_done, div_step = np.concatenate(done_step_all), np.concatenate(ref_done_all), np.concatenate(div_step_all)
    rows = {}
    for c in checkpoints:
        if not per_ck[c]["e"]:
            continue
        e, pos = np.concatenate(per_ck[c]["e"]), np.concatenate(per_ck[c]["pos"])
        rows[str(c)] = {"envs_flying": int(e.size), "rel_err_median": float(np.median(e)), "rel_err_p99": float(np.percentile(e, 99)),
                        "rel_err_max": float(e.max()), "pos_err_m_median": float(np.median(pos)), "pos_err_m_p99": float(np.percentile(pos, 99)),
                        "pos_err_m_max": float(pos.max()), "bit_identical_frames": int((e == 0).sum())}
    diverged = div_step < np.minimum(np.minimum(done_step, ref_done), t)
    return {"mode": mode, "actions": kind, "envs": n, "steps": t, "checkpoints": rows,
            "oracle_episodes_finished": int((ref_done < t).sum()),
            "episodes_ending_at_same_step": int((done_step == ref_done).sum()),
            "episodes_ending_within_one_step": int((np.abs(done_step - ref_done) <= 1).sum()),
            "envs_diverged_before_their_end": int(diverged.sum()), "divergence_threshold": DIVERGED,
            "first_divergence_step_min": int(div_step[diverged].min()) if diverged.any() else None,
            "first_divergence_step_p01": float(np.percentile(np.where(diverged, div_step, t), 1)),
            "first_divergence_step_median_of_diverged": float(np.median(div_step[diverged])) if diverged.any() else None,
            "reward_abs_err_max_before_divergence": reward_err}


if __name__ == "__main__":
    rep = {}
    ck = (1, 10, 30, 100, 200, 300, 600, 1000)
    for kind in ("uniform", "gentle"):
        rep[kind] = {"fp64_4096x1000": trajectory_stats("fp64", kind, 4096, 1000, ck),
                     "fp32_4096x1000": trajectory_stats("fp32", kind, 4096, 1000, ck),
                     "fp32_65536x200": trajectory_stats("fp32", kind, 65536, 200, (1, 10, 30, 100, 200))}
    print(json.dumps(rep, indent=1))
