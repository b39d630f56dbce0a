"""NumPy restatement of the part of ``jax.random`` the reference's examples use
(TEST INFRASTRUCTURE).

The reference's stored integration goldens (tests/integration_tests.py:99-129) are produced
from data drawn with ``jax.random`` (examples/regression.py:43,66-70).  jax is not installed
here, so the published Threefry-2x32 generator (Salmon et al., SC'11; the default
``jax_default_prng_impl='threefry2x32'`` with ``jax_threefry_partitionable=True``, the
default since jax 0.5; the reference pins jax 0.7.1 in uv.lock:1159-1176) is restated so
those goldens become reachable.  The restatement is self-validating: if it were wrong the
goldens would not reproduce to 1e-8 (tests/test_oracle_goldens.py).
"""
from __future__ import annotations

import numpy as np
from scipy.special import erfinv

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, r):
    return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & _M32


def threefry2x32(k1: int, k2: int, c1: np.ndarray, c2: np.ndarray):
    """20-round Threefry-2x32 on uint32 lanes (carried in uint64 and masked)."""
    ks = [np.uint64(k1), np.uint64(k2), np.uint64(k1 ^ k2 ^ 0x1BD11BDA)]
    x0 = (np.asarray(c1, np.uint64) + ks[0]) & _M32
    x1 = (np.asarray(c2, np.uint64) + ks[1]) & _M32
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r)
            x1 = x1 ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M32
        x1 = (x1 + ks[(i + 2) % 3] + np.uint64(i + 1)) & _M32
    return x0, x1


def key(seed: int):
    """``jax.random.key(seed)`` -> (hi, lo) words of the 64-bit seed."""
    return (int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF


def _iota_2x32(n: int):
    idx = np.arange(n, dtype=np.uint64)
    return idx >> np.uint64(32), idx & _M32


def split(k, num: int = 2):
    """``jax.random.split`` (fold-like, partitionable threefry)."""
    hi, lo = _iota_2x32(num)
    b1, b2 = threefry2x32(k[0], k[1], hi, lo)
    return [(int(a), int(b)) for a, b in zip(b1, b2)]


def random_bits64(k, shape):
    n = int(np.prod(shape))
    hi, lo = _iota_2x32(n)
    b1, b2 = threefry2x32(k[0], k[1], hi, lo)
    return ((b1 << np.uint64(32)) | b2).reshape(shape)


def uniform(k, shape, minval=0.0, maxval=1.0):
    """``jax.random.uniform`` for float64: 52 mantissa bits OR'd onto 1.0, minus 1."""
    bits = random_bits64(k, shape)
    fb = (bits >> np.uint64(64 - 52)) | np.float64(1.0).view(np.uint64)
    floats = fb.view(np.float64) - 1.0
    minval = np.float64(minval)
    maxval = np.float64(maxval)
    return np.maximum(minval, floats * (maxval - minval) + minval)


def normal(k, shape):
    """``jax.random.normal`` for float64: sqrt(2) * erfinv(U(-1+eps, 1))."""
    lo = np.nextafter(np.float64(-1.0), np.float64(0.0))
    u = uniform(k, shape, lo, 1.0)
    return np.sqrt(2.0) * erfinv(u)


# ---- the pre-"partitionable" stream (jax < 0.5 default), which the reference's stored goldens use ----
def _threefry_original(k, counts):
    counts = np.asarray(counts, np.uint64)
    n = len(counts)
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, np.uint64)])
    h = len(counts) // 2
    a, b = threefry2x32(k[0], k[1], counts[:h], counts[h:])
    return np.concatenate([a, b])[:n]


def split_original(k, num: int = 2):
    out = _threefry_original(k, np.arange(2 * num)).reshape(num, 2)
    return [(int(r[0]), int(r[1])) for r in out]


def random_bits64_original(k, n: int):
    out = _threefry_original(k, np.arange(2 * n))
    return (out[:n] << np.uint64(32)) | out[n:]
