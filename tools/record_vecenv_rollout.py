#!/usr/bin/env python3
"""Record what F16VecEnv (the SB3 VecEnv boundary, CUDA engine) returns over one rollout - run on the GPU box:

    python tools/record_vecenv_rollout.py gpurun_out/f16vecenv_rollout_fp64.npz

The file is committed as tests/golden/f16vecenv_rollout_fp64.npz. tests/test_reference_loop.py (CPU, in the
container that has /root/reference) then lets the REFERENCE's own `OnPolicyAlgorithm.collect_rollouts`
(stable_baselines3/common/on_policy_algorithm.py:162-262) and `RolloutBuffer` consume these recorded returns
through a replaying VecEnv, and compares the buffer they fill with the one the same loop fills from the
reference's own env stack (Monitor + DummyVecEnv around jsbsim_gym.py, FDM = the oracle) on the same seeds
and actions. Six envs, 320 steps, FP64 parity mode, carry-over reset (what DummyVecEnv's env objects go through):
envs 0-1 fly gently for the whole rollout, 2-3 are pushed into the ground (crash -> auto-reset -> new episode),
4-5 have their current_step set to 1100 / 1150 after the first step and are cut by the time limit (TimeLimit.truncated + terminal_observation).
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from f16_jsb_b200 import F16VecEnv, _lib  # noqa: E402

N, T, SEED = 6, 320, 100
PRESET_STEPS = {4: 1100, 5: 1150}


def action_table():
    rng = np.random.default_rng(7)
    a = np.zeros((T, N, 4), np.float32)
    a[..., :3] = rng.normal(0.0, 0.12, size=(T, N, 3))
    a[..., 1] -= 0.05
    a[..., 3] = 0.7
    a[:, 2, 1] = 0.85            # stick forward / back: one of the two dives into the ground
    a[:, 3, 1] = -0.9
    a[:, 3, 0] = 0.5
    a[:, 2, 3] = 1.0
    return np.clip(a, [-1, -1, -1, 0], [1, 1, 1, 1]).astype(np.float32)


def main(out):
    venv = F16VecEnv(N, mode="fp64", seed=0, host_obs="window", lazy_infos=False, reset_mode="carryover")
    venv.seed(SEED)
    obs0 = np.array(venv.reset())
    acts = action_table()
    obs = np.zeros((T, N, 10, 15), np.float32)
    rew = np.zeros((T, N), np.float32)
    done = np.zeros((T, N), np.bool_)
    d_step, d_env, d_trunc, d_term, d_r, d_l = [], [], [], [], [], []
    for t in range(T):
        o, r, d, infos = venv.step(acts[t])
        obs[t], rew[t], done[t] = o, r, d
        if t == 0:
            # after the first step (the kernel recognises an env object's very first flight frame by current_step == 1:
            # it still carries the mass properties of the 1500-lb tanks' CG)
            for i, st in PRESET_STEPS.items():
                _lib.check(venv.env.lib.f16_set_env_step(venv.env._h, i, st), "f16_set_env_step")
        assert len(infos) == N
        for i in range(N):
            info = infos[i]
            if d[i]:
                assert set(info) >= {"TimeLimit.truncated", "terminal_observation", "episode"}, info.keys()
                d_step.append(t); d_env.append(i); d_trunc.append(bool(info["TimeLimit.truncated"]))
                d_term.append(np.array(info["terminal_observation"], np.float32))
                d_r.append(float(info["episode"]["r"])); d_l.append(int(info["episode"]["l"]))
            else:
                assert not info.get("TimeLimit.truncated", False) and "terminal_observation" not in info
    np.savez_compressed(out, n_envs=N, n_steps=T, seed=SEED, preset_env=np.array(list(PRESET_STEPS)), preset_step=np.array(list(PRESET_STEPS.values())),
                        actions=acts, reset_obs=obs0, obs=obs, rewards=rew, dones=done,
                        done_step=np.array(d_step), done_env=np.array(d_env), done_truncated=np.array(d_trunc),
                        done_terminal_obs=np.array(d_term, np.float32).reshape(-1, 10, 15), done_ep_r=np.array(d_r), done_ep_l=np.array(d_l),
                        obs_space_low=venv.observation_space.low, obs_space_high=venv.observation_space.high,
                        act_space_low=venv.action_space.low, act_space_high=venv.action_space.high,
                        meta=np.array("F16VecEnv(mode='fp64', host_obs='window', reset_mode='carryover'), libf16b200 %s" % venv.env.lib.f16_version().decode()))
    print("recorded %d steps x %d envs: %d episodes ended (%d by the time limit), lengths %s" % (T, N, len(d_step), sum(d_trunc), d_l))
    venv.close()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/f16vecenv_rollout_fp64.npz")
