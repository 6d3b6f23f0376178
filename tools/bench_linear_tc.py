#!/usr/bin/env python3
"""Times the tensor-core Linear forward (f16_lma_linear_forward) against torch's F.linear on the policy's layer shapes at
one AM-PPO minibatch (131 072 samples): us per call, achieved GB/s of the algorithmic bytes (x read + y written)."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from f16_jsb_b200 import _lib  # noqa: E402

L = _lib.load()
B = 131072
LAYERS = [("embed 17->64", 10 * B, 17, 64), ("latent 128->32", 5 * B, 128, 32), ("qkv 32->96", 5 * B, 32, 96), ("proj 32->32", 5 * B, 32, 32),
          ("fc 32->128", 5 * B, 32, 128), ("fc2 128->32", 5 * B, 128, 32), ("pi0 160->64", B, 160, 64), ("pi1 64->64", B, 64, 64),
          ("vf1 128->64", B, 128, 64),
          ("d latent 32->128", 5 * B, 32, 128), ("d qkv 96->32", 5 * B, 96, 32), ("d pi0 64->160", B, 64, 160)]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


out = []
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, rows, k, n in LAYERS:
    if only and only not in name:
        continue
    xs = [torch.randn((rows, k), device="cuda") for _ in range(3)]           # > L2 in rotation
    w = torch.randn((n, k), device="cuda") * 0.1
    b = torch.randn((n,), device="cuda")
    ys = [torch.empty((rows, n), device="cuda") for _ in range(3)]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    i = [0]

    def mine():
        j = i[0] % 3
        i[0] += 1
        _lib.check(L.f16_lma_linear_forward(rows, k, n, C.c_void_p(xs[j].data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                            C.c_void_p(ys[j].data_ptr()), st), "linear")

    def lib():
        j = i[0] % 3
        i[0] += 1
        torch.nn.functional.linear(xs[j], w, b, ) if False else torch.addmm(b, xs[j], w.t(), out=ys[j])

    t_mine, t_lib = timeit(mine), timeit(lib)
    gb = rows * (k + n) * 4 / 1e9
    out.append({"layer": name, "rows": rows, "tc_us": round(t_mine, 1), "torch_us": round(t_lib, 1), "tc_GBps": round(gb / (t_mine * 1e-6), 0),
                "torch_GBps": round(gb / (t_lib * 1e-6), 0)})
    print(out[-1], flush=True)
print(json.dumps({"minibatch": B, "layers": out, "sum_tc_us": sum(o["tc_us"] for o in out), "sum_torch_us": sum(o["torch_us"] for o in out)}))
