#!/usr/bin/env python3
"""Times the tensor-core weight-gradient kernel (f16_lma_linear_wgrad_tc) against the FP32 slab kernel (f16_lma_linear_wgrad)
on the policy's layer shapes at one AM-PPO minibatch (131 072 samples): us per call, GB/s of the algorithmic bytes (x + dy read)."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from f16_jsb_b200 import _lib  # noqa: E402

L = _lib.load()
B = 131072
LAYERS = [("latent 128->32", 5 * B, 128, 32), ("qkv 32->96", 5 * B, 32, 96), ("proj 32->32", 5 * B, 32, 32), ("fc 32->128", 5 * B, 32, 128),
          ("fc2 128->32", 5 * B, 128, 32), ("pi0 160->64", B, 160, 64), ("pi1 64->64", B, 64, 64), ("vf1 128->64", B, 128, 64)]


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
    xs = [torch.randn((rows, k), device="cuda") for _ in range(3)]
    dys = [torch.randn((rows, n), device="cuda") for _ in range(3)]
    dw = torch.empty((n, k), device="cuda")
    db = torch.empty((n,), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    i = [0]

    def run(fn):
        def f():
            j = i[0] % 3
            i[0] += 1
            _lib.check(getattr(L, fn)(rows, k, n, C.c_void_p(xs[j].data_ptr()), C.c_void_p(dys[j].data_ptr()), C.c_void_p(dw.data_ptr()),
                                      C.c_void_p(db.data_ptr()), st), fn)
        return f

    t_tc, t_slab = timeit(run("f16_lma_linear_wgrad_tc")), timeit(run("f16_lma_linear_wgrad"))
    gb = rows * (k + n) * 4 / 1e9
    out.append({"layer": name, "rows": rows, "tc_us": round(t_tc, 1), "slab_us": round(t_slab, 1), "tc_GBps": round(gb / (t_tc * 1e-6), 0),
                "slab_GBps": round(gb / (t_slab * 1e-6), 0)})
    print(out[-1], flush=True)
print(json.dumps({"minibatch": B, "layers": out, "sum_tc_us": sum(o["tc_us"] for o in out), "sum_slab_us": sum(o["slab_us"] for o in out)}))
