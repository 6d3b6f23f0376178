"""NumPy restatement of the keep mask of the fused dropout kernels (TEST INFRASTRUCTURE ONLY).

include/f16_lma.h: "the keep mask is a function of (seed, element index)": Philox4x32-10 (Salmon et al., SC'11; the counter-based
generator of Random123) with counter (g lo, g hi, 0x454c, 0x4d57) and key (seed lo, seed hi), g = element index / 8; element
8 g + i takes bits 16 (i % 2) .. 16 (i % 2) + 15 of output word i / 2 and is DROPPED when that 16-bit draw is below
round(p * 65536); kept elements are scaled by 65536 / (65536 - round(p * 65536)) (csrc/f16_lma_elementwise.cu: keep8, make_keep).
The generator is pinned to Random123's published known-answer vectors in tests/test_policy_oracle.py. Only tests/ import this.
"""
import numpy as np

M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Arrays (or scalars) of uint32 -> four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0, p1 = M0 * c0.astype(np.uint64), M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def keep_factors(n, p, seed):
    """float32 keep factor (0 or the scale) of elements 0 .. n - 1 (n a multiple of 8)."""
    assert n % 8 == 0
    g = np.arange(n // 8, dtype=np.uint64)
    w = philox4x32_10((g & np.uint64(0xffffffff)).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32), np.uint32(0x454c), np.uint32(0x4d57),
                      np.uint32(seed & 0xffffffff), np.uint32((seed >> 32) & 0xffffffff))
    thr = int(min(65535, max(0, round(float(np.float32(p) * np.float32(65536.0))))))
    scale = np.float32(65536.0) / (np.float32(65536.0) - np.float32(thr))
    draws = np.stack([(w[i >> 1] >> np.uint32(16 * (i & 1))) & np.uint32(0xffff) for i in range(8)], axis=1).reshape(-1)
    return np.where(draws < thr, np.float32(0.0), scale).astype(np.float32)
