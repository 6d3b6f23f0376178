#!/usr/bin/env python3
"""The rollout's policy forward (obs -> actions, values, log-probabilities, clipped actions): the single kernel
f16_lma_policy_forward against the torch modules it replaces, eager and replayed as a CUDA graph. CUDA events, 200 calls.

    python tools/bench_policy_forward.py --envs 4096 65536
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from f16_jsb_b200.constants import ACTION_HIGH, ACTION_LOW  # noqa: E402
from f16_jsb_b200.lma import LMAActorCritic, PolicyForwardKernel  # noqa: E402


def timed(fn, iters=200):
    for _ in range(10):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3          # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[4096, 65536])
    args = ap.parse_args()
    torch.manual_seed(0)
    net = LMAActorCritic().cuda().eval()
    low, high = torch.as_tensor(ACTION_LOW).cuda(), torch.as_tensor(ACTION_HIGH).cuda()
    fused = PolicyForwardKernel(net, low, high)
    rows = []
    for n in args.envs:
        obs = torch.randn((n, 10, 15), device="cuda")
        noise = torch.randn((n, 4), device="cuda")

        def modules():
            with torch.no_grad():
                actions, values, log_probs = net(obs)
                return actions, values, log_probs, torch.maximum(torch.minimum(actions, high), low)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                modules()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            modules()
        flops = 2.0 * 200e3 * n
        t_k = timed(lambda: fused(obs, noise))
        rows.append({"envs": n, "kernel_us": t_k, "modules_eager_us": timed(modules), "modules_cuda_graph_us": timed(graph.replay),
                     "kernel_tflops_fp32": flops / t_k * 1e-6})
    print(json.dumps({"what": "policy forward per call, CUDA events", "rows": rows}))


if __name__ == "__main__":
    main()
