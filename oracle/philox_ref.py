"""NumPy model of the counter-based generator the GPU path uses -- TEST INFRASTRUCTURE.

Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
1, 2, 3", SC'11; Random123 v1.14 `philox4x32_R(10, ...)`): multipliers 0xD2511F53 / 0xCD9E8D57,
Weyl key increments 0x9E3779B9 / 0xBB67AE85.  Pinned by the Random123 known-answer vectors in
tests/test_philox.py.  The counter/stream layout mirrors rvmc_b200/csrc/philox.cuh and
rvmc_constants.h so tests can regenerate, on the CPU, exactly the uniform draws a kernel used
("the same exported uniform draws" of the parity protocol).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_ORBIT_A, STREAM_ORBIT_B, STREAM_SIGMA_DYN, STREAM_CLUSTER, STREAM_MIX = 0, 1, 2, 3, 4
STREAM_BOOT, STREAM_SLOT, STREAM_EPOCH, STREAM_SINGLE, STREAM_MASS = 5, 6, 7, 8, 9


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays; returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def philox_block(seed, item, stream, sub=0):
    """ctr = (item_lo, item_hi, stream, sub), key = (seed_lo, seed_hi) -> (n,4) uint32."""
    item = np.asarray(item, dtype=np.uint64)
    w = philox4x32_10(item & MASK, item >> np.uint64(32), stream, sub, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=-1)


def u01(w):
    """[0,1) 24-bit uniform, exact in fp32 and fp64 (philox.cuh u01_f32/u01_f64)."""
    return (np.asarray(w, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) * 2.0 ** -24


def u01_open(w):
    """(0,1] variant used under the logarithm of Box-Muller."""
    return ((np.asarray(w, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24


def box_muller(w1, w2):
    """rvmc_math.cuh box_muller_*: z0 = r cos(theta), z1 = r sin(theta)."""
    r = np.sqrt(-2.0 * np.log(u01_open(w1)))
    th = 2 * np.pi * u01(w2)
    return r * np.cos(th), r * np.sin(th)


def noise6(w):
    """rvmc_math.cuh noise6_*: six normals from one Philox block [n, 4] -- three Box-Muller pairs, each
    from a 22-bit radius uniform in (0, 1] and a 20-bit angle.  Returns [6, n] (z0 = r cos, z1 = r sin)."""
    w = np.asarray(w, dtype=np.uint64)
    x, y, z, ww = w[:, 0], w[:, 1], w[:, 2], w[:, 3]
    u = [x >> 10, y & 0x3FFFFF, (((z << 10) | (ww >> 22)) & 0x3FFFFF)]
    a = [(((x << 10) | (y >> 22)) & 0xFFFFF), z >> 12, (ww >> 2) & 0xFFFFF]
    out = np.empty((6, len(w)))
    for p in range(3):
        r = np.sqrt(-2.0 * np.log((u[p].astype(np.float64) + 1.0) * 2.0 ** -22))
        th = a[p].astype(np.float64) * 5.9921124526782858e-06
        out[2 * p] = r * np.cos(th)
        out[2 * p + 1] = r * np.sin(th)
    return out

