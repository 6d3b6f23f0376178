#!/usr/bin/env python3
"""BASELINE configs[4]: end-to-end AM-PPO (n_steps 2048, LMA extractor) rollout + update consuming GPU env
observations, on one B200. Reports env-steps/s of the rollout alone (policy forward + env step + rollout store),
samples/s of the update (10 epochs of minibatches: gather with stack rebuild, LMA forward/backward, DAG step)
and the overall env-steps/s of one iteration. The reference runs one env with minibatches of 256
(train.py:145-160); a batched env needs proportionally larger minibatches, so both are parameters.

    python tools/bench_amppo.py --envs 1024 --n-steps 2048 --batch-size 32768 --iterations 2
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from f16_jsb_b200 import F16BatchedEnv  # noqa: E402
from f16_jsb_b200.amppo import AMPPO, AMPPOConfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--n-steps", type=int, default=2048)
    ap.add_argument("--batch-size", type=int, default=32768)
    ap.add_argument("--n-epochs", type=int, default=10)
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--optimizer", default="DAG")
    ap.add_argument("--no-am-ppo", action="store_true")
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fused-forward", action="store_true", help="rollout policy forward through the torch modules instead of f16_lma_policy_forward")
    args = ap.parse_args()
    env = F16BatchedEnv(args.envs, mode="fp32", seed=0)
    cfg = AMPPOConfig(n_steps=args.n_steps, batch_size=args.batch_size, n_epochs=args.n_epochs, optimizer=args.optimizer,
                      use_am_ppo=not args.no_am_ppo, tf32=args.tf32, cuda_graph=not args.no_graph,
                      fused_policy_forward=not args.no_fused_forward)
    algo = AMPPO(env, cfg)
    rows = []
    for it in range(args.iterations + 1):          # iteration 0 is the warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        algo.collect_rollouts()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        algo.train()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if it:
            rows.append((t1 - t0, t2 - t1))
    n = args.envs * args.n_steps
    roll = sum(r[0] for r in rows) / len(rows)
    upd = sum(r[1] for r in rows) / len(rows)
    st = env.stats()
    print(json.dumps({
        "workload": "BASELINE configs[4]: AM-PPO (n_steps %d, LMA extractor, %s) rollout + update on GPU env observations" % (args.n_steps, args.optimizer),
        "tf32": args.tf32, "cuda_graph": not args.no_graph, "fused_policy_forward": algo._fused_act is not None, "envs": args.envs, "n_steps": args.n_steps, "batch_size": args.batch_size, "n_epochs": args.n_epochs,
        "transitions_per_iteration": n, "rollout_s": roll, "update_s": upd,
        "rollout_env_steps_per_s": n / roll, "update_samples_per_s": n * args.n_epochs / upd,
        "overall_env_steps_per_s": n / (roll + upd), "last_stats": algo.last_stats,
        "episodes": st["episodes"], "mean_return": st["return_sum"] / max(1.0, st["episodes"])}))
    env.close()


if __name__ == "__main__":
    main()
