#!/usr/bin/env python3
"""Small runs of the tensor-core Linear / weight-gradient kernels for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_tc.py
Ragged row counts (tail tiles / chunks), the vector and the 17-feature paths, both orientations of the weight gradient."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from f16_jsb_b200 import _lib  # noqa: E402

L = _lib.load()
st = None
worst = 0.0
for rows, k, n in [(1000, 17, 64), (777, 32, 96), (1300, 128, 32), (515, 160, 64), (2049, 64, 160)]:
    x = torch.randn((rows, k), device="cuda"); w = torch.randn((n, k), device="cuda"); b = torch.randn((n,), device="cuda")
    y = torch.empty((rows, n), device="cuda")
    _lib.check(L.f16_lma_linear_forward(rows, k, n, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(y.data_ptr()), st), "fwd")
    torch.cuda.synchronize()
    worst = max(worst, float((y - torch.nn.functional.linear(x, w, b)).abs().max()))
for rows, k, n in [(1000, 32, 96), (777, 128, 32), (1300, 160, 64), (515, 64, 64), (2049, 32, 32)]:
    x = torch.randn((rows, k), device="cuda"); dy = torch.randn((rows, n), device="cuda")
    dw = torch.empty((n, k), device="cuda"); db = torch.empty((n,), device="cuda")
    _lib.check(L.f16_lma_linear_wgrad_tc(rows, k, n, C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dw.data_ptr()), C.c_void_p(db.data_ptr()), st), "wgrad")
    torch.cuda.synchronize()
    worst = max(worst, float((dw - dy.t() @ x).abs().max()), float((db - dy.sum(0)).abs().max()))
print("sanitize_tc: ok, worst abs difference against torch %.3e" % worst)
