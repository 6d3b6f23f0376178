"""Tensor-core Linear forward (csrc/f16_lma_linear.cu, include/f16_lma.h) against float64 and against torch's FP32
F.linear on the shapes of the reference's LMA policy (jsbsim_gym/LMA_features.py:221-279,315-385; SB3 heads)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

# (in_features, out_features): forward shapes of the policy, then the input-gradient shapes (weight = W^T)
SHAPES = [(17, 64), (128, 32), (32, 96), (32, 32), (32, 128), (160, 64), (64, 64), (128, 64),
          (96, 32), (64, 160), (64, 128), (8, 32), (24, 64), (32, 256), (160, 128), (128, 160), (64, 256)]


def _call(x, w, b):
    from f16_jsb_b200 import _lib
    L = _lib.load()
    y = torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.f16_lma_linear_forward(x.shape[0], x.shape[1], w.shape[0], C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()),
                                        C.c_void_p(b.data_ptr() if b is not None else 0), C.c_void_p(y.data_ptr()), st), "f16_lma_linear_forward")
    return y


@pytest.mark.parametrize("k,n", SHAPES)
@pytest.mark.parametrize("rows", [1, 127, 128, 129, 4096 + 37, 200_000])
def test_linear_tc_matches_float64(k, n, rows):
    from f16_jsb_b200 import _lib
    assert _lib.load().f16_lma_linear_supported(k, n) == 1
    g = torch.Generator(device="cuda").manual_seed(1000 * k + n + rows)
    x = torch.randn((rows, k), device="cuda", generator=g) * 3.0
    w = torch.randn((n, k), device="cuda", generator=g) * 0.5
    b = torch.randn((n,), device="cuda", generator=g)
    for bias in (b, None):
        y = _call(x, w, bias)
        torch.cuda.synchronize()
        ref = x.double() @ w.double().t() + (bias.double() if bias is not None else 0.0)
        y32 = torch.nn.functional.linear(x, w, bias)
        scale = (x.abs().double() @ w.abs().double().t()) + 1.0         # magnitude of the sum each output is made of
        err = ((y.double() - ref).abs() / scale).max().item()
        err32 = ((y32.double() - ref).abs() / scale).max().item()
        # FP32-accurate: a few float32 ulp of the sum's magnitude, and no worse than 4x the library's FP32 GEMM
        # tensor-core accumulation truncates where the FP32 FMA pipe rounds: a few float32 ulp of the sum's magnitude,
        # within 6x of the library's FP32 GEMM on the same data
        assert err < 1.2e-6, (k, n, rows, err, err32)
        assert err < 6 * err32 + 3e-7, (k, n, rows, err, err32)


def test_linear_tc_refuses_unsupported_shapes():
    from f16_jsb_b200 import _lib
    L = _lib.load()
    assert L.f16_lma_linear_supported(64, 4) == 0 and L.f16_lma_linear_supported(64, 1) == 0
    assert L.f16_lma_linear_supported(40, 32) == 0 and L.f16_lma_linear_supported(32, 48) == 0
    assert L.f16_lma_linear_supported(160, 320) == 0          # more than four column groups
    x = torch.zeros((8, 64), device="cuda")
    w = torch.zeros((4, 64), device="cuda")
    y = torch.zeros((8, 4), device="cuda")
    rc = L.f16_lma_linear_forward(8, 64, 4, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), None, C.c_void_p(y.data_ptr()), None)
    assert rc != 0 and b"unsupported shape" in L.f16_last_error()


def test_linear_tc_random_shapes_and_row_counts():
    """Seeded sweep: every supported (in, out) pair drawn from the sizes the kernel is built for, ragged row counts, with
    and without bias, non-trivial magnitudes; guard bytes after y must stay untouched."""
    import random
    from f16_jsb_b200 import _lib
    L = _lib.load()
    rnd = random.Random(7)
    ins = [1, 5, 8, 17, 24, 31, 32, 64, 96, 128, 160]
    outs = [32, 64, 96, 128, 160, 192, 224, 256]
    done = 0
    while done < 60:
        k, n = rnd.choice(ins), rnd.choice(outs)
        if not L.f16_lma_linear_supported(k, n):
            continue
        rows = rnd.choice([1, 2, 127, 128, 129, 255, 1000, 4097, rnd.randrange(1, 70000)])
        g = torch.Generator(device="cuda").manual_seed(done)
        x = torch.randn((rows, k), device="cuda", generator=g) * rnd.choice([1e-3, 1.0, 50.0])
        w = torch.randn((n, k), device="cuda", generator=g) * rnd.choice([0.05, 1.0])
        b = torch.randn((n,), device="cuda", generator=g) if rnd.random() < 0.7 else None
        ybuf = torch.full((rows * n + 64,), 123.0, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.f16_lma_linear_forward(rows, k, n, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()),
                                            C.c_void_p(b.data_ptr() if b is not None else 0), C.c_void_p(ybuf.data_ptr()), st), "f16_lma_linear_forward")
        torch.cuda.synchronize()
        assert bool((ybuf[rows * n:] == 123.0).all()), (k, n, rows)
        y = ybuf[:rows * n].view(rows, n)
        ref = x.double() @ w.double().t() + (b.double() if b is not None else 0.0)
        scale = (x.abs().double() @ w.abs().double().t()) + (b.abs().double() if b is not None else 0.0) + 1e-30
        err = ((y.double() - ref).abs() / scale).max().item()
        assert err < 1.2e-6, (k, n, rows, err)
        done += 1
