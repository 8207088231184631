"""Host mirror of the device permutation generator (``csrc/common.cuh``).

Permutation ``p`` under ``seed`` is a keyed bijection of ``[0, n)``: eight rounds of a generalised
Feistel network on ``Z_a x Z_b`` (``b = 2^s ~ sqrt(n)``, ``a = ceil(n/b)``, additive combine,
murmur3-finalizer round function), round keys drawn from Philox4x32-10 with counter
``(p_lo, p_hi, block, 0x5C0B200)`` and key ``(seed_lo, seed_hi)``, cycle-walked into range (the
domain exceeds ``n`` by less than ``b``, so the walk almost never iterates).  No index array is ever stored on the device; this
numpy implementation reproduces it bit for bit so a Philox-mode run can be replayed on the CPU
oracle (``tests/``) or exported.
"""

from __future__ import annotations

import numpy as np

FEISTEL_ROUNDS = 8
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = (int(x) & 0xFFFFFFFF for x in counter)
    k0, k1 = (int(x) & 0xFFFFFFFF for x in key)
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def round_keys(seed: int, perm_index: int):
    keys = []
    for b in range(FEISTEL_ROUNDS // 4):
        keys.extend(
            philox4x32_10(
                (perm_index & 0xFFFFFFFF, (perm_index >> 32) & 0xFFFFFFFF, b, 0x5C0B200),
                (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF),
            )
        )
    return keys


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & _M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & _M32
    x ^= x >> np.uint64(16)
    return x


def _domain(n: int):
    """(a, s): Z_a x Z_{2^s} with 2^s ~ sqrt(n) and a = ceil(n / 2^s)."""
    bits = 1
    while bits < 32 and (1 << bits) < n:
        bits += 1
    s = (bits + 1) // 2
    a = max(1, (n + (1 << s) - 1) >> s)
    return a, s


def _feistel_once(v: np.ndarray, a: int, s: int, keys) -> np.ndarray:
    bmask = np.uint64((1 << s) - 1)
    L = v >> np.uint64(s)
    R = v & bmask
    for r in range(0, FEISTEL_ROUNDS, 2):
        h = _mix32((R * np.uint64(0x9E3779B1) + np.uint64(keys[r])) & _M32)
        f = (h * np.uint64(a)) >> np.uint64(32)
        t = L + f
        t = np.where(t >= a, t - np.uint64(a), t)
        L, R = R, t
        h2 = _mix32((R * np.uint64(0x9E3779B1) + np.uint64(keys[r + 1])) & _M32)
        f2 = h2 >> np.uint64(32 - s)
        t2 = (L + f2) & bmask
        L, R = R, t2
    return (L << np.uint64(s)) | R


def permutation(seed: int, perm_index: int, n: int) -> np.ndarray:
    """π_p as an int32 array: ``out[i] = π_p(i)`` — identical to ``sc_philox_permutation``."""
    keys = round_keys(seed, perm_index)
    a, s = _domain(n)
    v = _feistel_once(np.arange(n, dtype=np.uint64), a, s, keys)
    bad = v >= n
    while bad.any():
        v[bad] = _feistel_once(v[bad], a, s, keys)
        bad = v >= n
    return v.astype(np.int32)
