"""Tensor-core weight / bias gradients (csrc/f16_lma_wgrad_tc.cu, include/f16_lma.h) against float64 and against the FP32
slab kernel on the shapes of the reference's LMA policy (jsbsim_gym/LMA_features.py:221-279,315-385; SB3 heads)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

# (in_features, out_features)
SHAPES = [(128, 32), (32, 96), (32, 32), (32, 128), (160, 64), (64, 64), (128, 64), (96, 32), (64, 32), (96, 64), (64, 96)]


def _call(fn, x, dy, bias=True):
    from f16_jsb_b200 import _lib
    L = _lib.load()
    dw = torch.full((dy.shape[1], x.shape[1]), 7.0, dtype=torch.float32, device=x.device)       # must be overwritten
    db = torch.full((dy.shape[1],), 7.0, dtype=torch.float32, device=x.device) if bias else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(getattr(L, fn)(x.shape[0], x.shape[1], dy.shape[1], C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()),
                              C.c_void_p(dw.data_ptr()), C.c_void_p(db.data_ptr() if bias else 0), st), fn)
    return dw, db


@pytest.mark.parametrize("k,n", SHAPES)
@pytest.mark.parametrize("rows", [1, 15, 16, 17, 255, 256, 257, 4096 + 37, 300_000])
def test_wgrad_tc_matches_float64(k, n, rows):
    from f16_jsb_b200 import _lib
    assert _lib.load().f16_lma_linear_wgrad_tc_supported(k, n) == 1
    g = torch.Generator(device="cuda").manual_seed(1000 * k + n + rows)
    x = torch.randn((rows, k), device="cuda", generator=g) * 2.0 + 0.3           # a non-zero mean: sums that grow
    dy = torch.randn((rows, n), device="cuda", generator=g) * 0.5 + 0.1
    dw, db = _call("f16_lma_linear_wgrad_tc", x, dy)
    dw32, db32 = _call("f16_lma_linear_wgrad", x, dy)
    torch.cuda.synchronize()
    ref = dy.double().t() @ x.double()
    scale = dy.abs().double().t() @ x.abs().double() + 1.0
    err = ((dw.double() - ref).abs() / scale).max().item()
    err32 = ((dw32.double() - ref).abs() / scale).max().item()
    # FP32-accurate: a few float32 ulp of the sum's magnitude, comparable with the FP32 slab kernel on the same data
    assert err < 2e-6, (k, n, rows, err, err32)
    assert err < 6 * err32 + 5e-7, (k, n, rows, err, err32)
    refb = dy.double().sum(0)
    scaleb = dy.abs().double().sum(0) + 1.0
    assert ((db.double() - refb).abs() / scaleb).max().item() < 2e-6
    dw2, _ = _call("f16_lma_linear_wgrad_tc", x, dy, bias=False)
    assert torch.allclose(dw2, dw, rtol=1e-5, atol=1e-5 * float(dw.abs().max()))      # atomics: order varies, values agree


def test_wgrad_tc_refuses_unsupported_shapes():
    from f16_jsb_b200 import _lib
    L = _lib.load()
    for k, n in [(17, 64), (64, 4), (64, 1), (160, 128), (128, 128), (32, 160), (192, 32)]:
        assert L.f16_lma_linear_wgrad_tc_supported(k, n) == 0, (k, n)
    x = torch.zeros((8, 17), device="cuda")
    dy = torch.zeros((8, 64), device="cuda")
    dw = torch.zeros((64, 17), device="cuda")
    rc = L.f16_lma_linear_wgrad_tc(8, 17, 64, C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dw.data_ptr()), None, None)
    assert rc != 0 and b"unsupported shape" in L.f16_last_error()


def test_wgrad_tc_random_shapes_and_row_counts():
    """Seeded sweep over the supported (in, out) pairs and ragged row counts, against float64; the FP32 slab kernel on the
    same data is the yardstick for the error."""
    import random
    from f16_jsb_b200 import _lib
    L = _lib.load()
    rnd = random.Random(11)
    done = 0
    while done < 50:
        k, n = rnd.choice([32, 64, 96, 128, 160]), rnd.choice([32, 64, 96, 128])
        if not L.f16_lma_linear_wgrad_tc_supported(k, n):
            continue
        rows = rnd.choice([1, 31, 32, 33, 63, 64, 65, 1000, 4097, rnd.randrange(1, 120000)])
        g = torch.Generator(device="cuda").manual_seed(100 + done)
        x = torch.randn((rows, k), device="cuda", generator=g) * rnd.choice([1e-2, 1.0, 20.0])
        dy = torch.randn((rows, n), device="cuda", generator=g) * rnd.choice([1e-3, 1.0])
        dw, db = _call("f16_lma_linear_wgrad_tc", x, dy)
        dw32, _ = _call("f16_lma_linear_wgrad", x, dy)
        torch.cuda.synchronize()
        ref = dy.double().t() @ x.double()
        scale = dy.abs().double().t() @ x.abs().double() + 1e-30
        err = ((dw.double() - ref).abs() / scale).max().item()
        err32 = ((dw32.double() - ref).abs() / scale).max().item()
        assert err < 2e-6 and err < 6 * err32 + 5e-7, (k, n, rows, err, err32)
        refb = dy.double().sum(0)
        assert ((db.double() - refb).abs() / (dy.abs().double().sum(0) + 1e-30)).max().item() < 2e-6, (k, n, rows)
        done += 1
