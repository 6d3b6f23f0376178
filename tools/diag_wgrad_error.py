import ctypes as C, os, sys, torch
sys.path.insert(0, ".")
from f16_jsb_b200 import _lib
L = _lib.load()
def run(fn, x, dy):
    n, k = dy.shape[1], x.shape[1]
    dw = torch.zeros((n, k), device="cuda"); db = torch.zeros((n,), device="cuda")
    _lib.check(getattr(L, fn)(x.shape[0], k, n, C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dw.data_ptr()), C.c_void_p(db.data_ptr()), None), fn)
    torch.cuda.synchronize(); return dw
for (rows, k, n) in [(40000, 160, 64), (655360, 32, 128), (655360, 128, 32), (131072, 64, 64)]:
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((rows, k), generator=g, device="cuda"); dy = torch.randn((rows, n), generator=g, device="cuda")
    ref = dy.double().t() @ x.double()
    out = []
    for fn in ("f16_lma_linear_wgrad_tc", "f16_lma_linear_wgrad"):
        dw = run(fn, x, dy)
        out.append(float((dw.double() - ref).abs().max() / ref.abs().max()))
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    run("f16_lma_linear_wgrad_tc", x, dy); t0.record()
    for _ in range(10): run("f16_lma_linear_wgrad_tc", x, dy)
    t1.record(); torch.cuda.synchronize()
    print("flush=%s rows=%d k=%d n=%d: max err / max|dW|: tc %.2e  slab %.2e   (%.1f us incl. sync)" % (os.environ.get("F16_WG_FLUSH", "64 (default)"), rows, k, n, out[0], out[1], t0.elapsed_time(t1) * 100))
