"""Philox4x32-10 in numpy: the framework's counter-based stream, for tests only.

sequential stream of path (pixel, sample): philox(ctr=(0, 0, sample, pixel), key=seed) seeds the drand48 recurrence
keyed medium draw:                         philox(ctr=(leaf, 1+depth, sample, pixel), key=seed)[0]
u01(x) = float32(x >> 8) * 2^-24
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)


def philox4x32_10(ctr, key):
    c = [np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for x in ctr]
    k0 = np.uint64(key[0]); k1 = np.uint64(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        n0 = (p1 >> np.uint64(32)) ^ c[1] ^ k0
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c[3] ^ k1
        n3 = p0 & mask
        c = [n0 & mask, n1, n2 & mask, n3]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return [x.astype(np.uint32) for x in c]


def u01(x):
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def seq_draws(seed, pixel, sample, n):
    """first n sequential draws of one path as float32: Philox(0, 0, sample, pixel) seeds the drand48 recurrence
    X <- (0x5DEECE66D X + 0xB) mod 2^48, a draw is the top 24 bits of X"""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    b = philox4x32_10((0, 0, sample, pixel), key)
    x = ((int(b[1]) & 0xFFFF) << 32) | int(b[0])
    out = []
    for _ in range(n):
        x = (x * 0x5DEECE66D + 0xB) & 0xFFFFFFFFFFFF
        out.append(np.float32(x >> 24) * np.float32(1.0 / 16777216.0))
    return np.array(out, dtype=np.float32)
