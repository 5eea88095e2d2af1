"""Philox4x32-10 (Salmon et al. 2011, Random123) in numpy: the point generator of BASELINE config 5 (SURVEY.md §8d:
"1e9 points uniform in the C3 root box, Philox4x32-10, seed 0x5DF0C7EE, counter = point index, 53-bit mantissa -> f64").
The device generator (hpsdf_uniform_points_device, csrc/points_kernel.cuh) must give these bits."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr, key):
    """ctr: (n, 4) uint32, key: (2,) uint32 -> (n, 4) uint32."""
    c = [np.ascontiguousarray(ctr[:, i], np.uint32) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return np.stack(c, 1)


def unit_double(hi, lo):
    """53 random bits -> [0, 1): (hi >> 5) * 2^26 + (lo >> 6), scaled by 2^-53."""
    return ((hi >> np.uint32(5)).astype(np.float64) * 67108864.0 + (lo >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def uniform_points(seed, first, n, lo, hi):
    """Points first .. first+n-1 of the stream `seed`: x, y from block (index, 0), z from block (index, 1)."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    ctr = np.zeros((n, 4), np.uint32)
    ctr[:, 0] = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (idx >> np.uint64(32)).astype(np.uint32)
    key = (np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    a = philox4x32_10(ctr, key)
    ctr[:, 2] = 1
    b = philox4x32_10(ctr, key)
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    u = np.stack([unit_double(a[:, 0], a[:, 1]), unit_double(a[:, 2], a[:, 3]), unit_double(b[:, 0], b[:, 1])], 1)
    return lo + u * (hi - lo)
