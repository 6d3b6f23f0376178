#!/usr/bin/env python3
"""The extractor's fused elementwise kernels (csrc/f16_lma_elementwise.cu) at the shapes of one AM-PPO minibatch of 131 072
samples, against the torch ops they replace; CUDA events, algorithmic bytes / time.

    python tools/bench_elementwise.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from f16_jsb_b200.lma import _DropoutAddFn, _EmbedActFn, sinusoidal_positions  # noqa: E402


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    B = 131072
    rows = []
    a = torch.randn((B, 10, 64), device="cuda", requires_grad=True)
    pos = sinusoidal_positions(10, 64).cuda()
    dy = torch.randn((B, 640), device="cuda")
    mb = a.numel() * 4 / 1e6

    def fused_embed():
        y = _EmbedActFn.apply(a, pos, 0.1, 4)
        y.backward(dy)

    def torch_embed():
        y = F.dropout(F.relu(a) + pos, 0.1, True)
        y = y.view(B, 10, 4, 16).permute(0, 2, 1, 3).reshape(B, 640)
        y.backward(dy)

    t_f, t_t = timed(fused_embed), timed(torch_embed)
    rows.append({"op": "embedding activation + head stacking, forward + backward", "tensor_mb": mb, "fused_us": t_f, "torch_us": t_t,
                 "fused_algorithmic_gbs": 5 * mb / t_f * 1e3})          # fwd: read a, write y; bwd: read a, dy, write da
    x = torch.randn((B, 5, 32), device="cuda", requires_grad=True)
    z = torch.randn((B, 5, 32), device="cuda", requires_grad=True)
    dz = torch.randn((B, 5, 32), device="cuda")
    mb2 = x.numel() * 4 / 1e6

    def fused_res():
        _DropoutAddFn.apply(x, z, 0.1).backward(dz)

    def torch_res():
        (z + F.dropout(x, 0.1, True)).backward(dz)

    t_f, t_t = timed(fused_res), timed(torch_res)
    rows.append({"op": "residual dropout z + drop(x), forward + backward", "tensor_mb": mb2, "fused_us": t_f, "torch_us": t_t,
                 "fused_algorithmic_gbs": 5 * mb2 / t_f * 1e3})        # fwd: read x, z, write y; bwd: read dy, write dx (+ autograd's grad accumulation, not counted)
    print(json.dumps({"what": "fused elementwise kernels against torch ops, one minibatch of 131 072 samples", "rows": rows}))


if __name__ == "__main__":
    main()
