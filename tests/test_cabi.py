"""The C-ABI library loads on a machine without a GPU, exports every symbol include/f16_b200.h
declares, and refuses - loudly, with a message - to do any work without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="f16_b200.h"):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(f16_[a-z0-9_]+)\s*\(", txt)))


def test_every_header_in_include_is_exported():
    from f16_jsb_b200 import _lib
    L = _lib.load()
    n = 0
    for header in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not header.endswith(".h"):
            continue
        for s in declared_symbols(header):
            assert hasattr(L, s), "%s declares %s but libf16b200.so does not export it" % (header, s)
            n += 1
    assert n >= 24


def test_header_symbols_are_exported():
    from f16_jsb_b200 import _lib
    L = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(L, s), "libf16b200.so does not export %s" % s
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms
    assert b"sm_100a" in L.f16_version()
    assert L.f16_num_state_fields() == 53


def test_state_field_enum_matches_library(state_fields):
    from f16_jsb_b200 import _lib
    assert len(state_fields) == _lib.load().f16_num_state_fields()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU box")
    from f16_jsb_b200 import _lib
    L = _lib.load()
    h = C.c_void_p()
    rc = L.f16_create(C.byref(h), 16, 0, 0)
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in L.f16_last_error()
    with pytest.raises(_lib.F16Error):
        from f16_jsb_b200 import F16BatchedEnv
        F16BatchedEnv(4)


def test_argument_validation_without_device():
    from f16_jsb_b200 import _lib
    L = _lib.load()
    assert L.f16_create(None, 4, 0, 0) != 0
    h = C.c_void_p()
    assert L.f16_create(C.byref(h), 0, 0, 0) != 0 and b"n_envs" in L.f16_last_error()
    assert L.f16_create(C.byref(h), 4, 0, 7) != 0 and b"mode" in L.f16_last_error()
    assert L.f16_step(None, None, 0, None) != 0
    assert L.f16_state_bytes(None) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "f16_jsb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f in ("f16_model.cuh", "f16_env.cuh", "f16_host_setup.h"), f
                assert "f16_oracle" not in src and "libf16oracle" not in src and "hostsim.so" not in src, f


def test_policy_forward_parameter_layout():
    """f16_lma_policy_entry (include/f16_lma.h): the packed buffer is tiled exactly by the entries, in the documented order, with
    the reference run's layer shapes (train.py:21-32,84); bad arguments are refused before any device work."""
    from f16_jsb_b200 import _lib
    L = _lib.load()
    n = L.f16_lma_policy_entries()
    cursor, shapes = 0, []
    for i in range(n):
        fin, fout, wo, bo = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
        assert L.f16_lma_policy_entry(i, C.byref(fin), C.byref(fout), C.byref(wo), C.byref(bo)) == 0
        shapes.append((fin.value, fout.value))
        assert wo.value == cursor
        if bo.value < 0:                      # the position table
            cursor += fin.value * fout.value
        elif fout.value == 0:                 # LayerNorm: weight, bias
            assert bo.value == cursor + fin.value
            cursor += 2 * fin.value
        else:                                 # Linear: transposed weight, bias
            assert bo.value == cursor + fin.value * fout.value
            cursor += fin.value * fout.value + fout.value
    assert cursor == L.f16_lma_policy_packed_size()
    block = [(32, 0), (32, 96), (32, 32), (32, 0), (32, 128), (128, 32)]
    assert shapes == [(10, 64), (17, 64), (128, 32)] + block + block + [(160, 64), (64, 64), (64, 4), (160, 128), (128, 64), (64, 1)]
    assert L.f16_lma_policy_entry(n, None, None, None, None) != 0
    assert L.f16_lma_policy_forward(0, *([None] * 2), 0, *([None] * 10)) != 0
    assert L.f16_lma_policy_forward(4, *([None] * 2), L.f16_lma_policy_packed_size() - 1, *([None] * 10)) != 0
    assert b"packed_len" in L.f16_last_error()
