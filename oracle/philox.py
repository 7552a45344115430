"""Oracle: Philox4x32-10 counter-based generator (TEST INFRASTRUCTURE ONLY).

Published algorithm (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC11).  The reference
draws from torch's CPU generator (pyprob/distributions/distribution.py:31-36); bit-exact agreement with
that stream is neither possible nor required — this oracle pins OUR integer stream so the
raw 32-bit words of the CUDA samplers are checked bit-exactly, and the float transforms to tolerance.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(seed, idx, offset):
    """seed, offset: python ints; idx: uint64 array -> [n,4] uint32."""
    idx = np.asarray(idx, dtype=np.uint64)
    c0 = idx & MASK
    c1 = idx >> np.uint64(32)
    c2 = np.full_like(idx, offset & 0xFFFFFFFF)
    c3 = np.full_like(idx, (offset >> 32) & 0xFFFFFFFF)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def u01(x):
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def u01_open0(x):
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)


def std_normal(a, b):
    u1, u2 = u01_open0(a).astype(np.float64), u01(b).astype(np.float64)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)
