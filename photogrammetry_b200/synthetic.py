"""Seeded synthetic descriptor sets for the BASELINE configs.

A counter-based splitmix64 stream: the idx-th output of stream ``seed`` is
``mix(seed + (idx+1)*GAMMA)``, so the generator vectorises in numpy; the test
suite cross-checks it bit-for-bit against an independent C implementation of
the same stream.

Distributions:
  U  i.i.d. uniform W-bit descriptors.
  C  "noisy copy": train = permuted copy of query, each bit flipped with
     probability floor(256*flip_p)/256, then a fraction outlier_p of the rows
     replaced by uniform noise.
"""
from __future__ import annotations

import numpy as np

from .descriptors import stride_for_bits

_GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def _sm64(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return _mix(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + (np.asarray(idx, dtype=np.uint64) + np.uint64(1)) * _GAMMA)


def _derive(seed: int, k: int) -> int:
    with np.errstate(over="ignore"):
        return int(_mix(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(k) * _GAMMA))


def uniform_descriptors(seed: int, n: int, desc_bits: int = 256, stride: int | None = None) -> np.ndarray:
    """``uint8[n, stride]`` i.i.d. uniform ``desc_bits``-bit descriptors."""
    stride = stride or stride_for_bits(desc_bits)
    nw = (desc_bits + 63) // 64
    idx = np.arange(n * nw, dtype=np.uint64)
    words = _sm64(seed, idx).reshape(n, nw)
    top = desc_bits - 64 * (nw - 1)
    if top < 64:
        words[:, nw - 1] &= np.uint64((1 << top) - 1)
    out = np.zeros((n, stride), dtype=np.uint8)
    raw = words.astype("<u8").view(np.uint8).reshape(n, nw * 8)
    out[:, :min(stride, nw * 8)] = raw[:, :min(stride, nw * 8)]
    return out


def noisy_copy_descriptors(seed: int, query: np.ndarray, desc_bits: int = 256,
                           flip_p: float = 0.10, outlier_p: float = 0.30) -> np.ndarray:
    """Distribution C: permuted, bit-flipped copy of ``query`` with outlier rows."""
    query = np.ascontiguousarray(query, dtype=np.uint8)
    n, stride = query.shape
    s_perm, s_flip, s_out, s_noise = (_derive(seed, k) for k in (1, 2, 3, 4))
    rows = np.arange(n, dtype=np.uint64)
    perm = np.argsort(_sm64(s_perm, rows), kind="stable")
    out = query[perm].copy()
    nbytes = (desc_bits + 7) // 8
    thr = int(flip_p * 256.0)
    chunk = max(1, (1 << 22) // max(nbytes, 1))
    for r0 in range(0, n, chunk):
        r1 = min(n, r0 + chunk)
        idx = (np.arange(r0, r1, dtype=np.uint64)[:, None] * np.uint64(nbytes)
               + np.arange(nbytes, dtype=np.uint64)[None, :])
        draws = _sm64(s_flip, idx).astype("<u8").view(np.uint8).reshape(r1 - r0, nbytes, 8)
        bits = (draws < thr)
        bitpos = np.arange(nbytes)[:, None] * 8 + np.arange(8)[None, :]
        bits &= (bitpos < desc_bits)[None, :, :]
        mask = (bits.astype(np.uint8) << np.arange(8, dtype=np.uint8)[None, None, :]).sum(axis=2).astype(np.uint8)
        out[r0:r1, :nbytes] ^= mask
    u = (_sm64(s_out, rows) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    outlier = u < outlier_p
    if outlier.any():
        noise = uniform_descriptors(s_noise, n, desc_bits, stride)
        out[outlier] = noise[outlier]
    return out


def config2_pair(n: int = 8192, dist: str = "U", desc_bits: int = 256):
    """BASELINE configs[1]: one synthetic pair (query, train)."""
    if dist == "U":
        return uniform_descriptors(1234, n, desc_bits), uniform_descriptors(5678, n, desc_bits)
    if dist == "C":
        q = uniform_descriptors(1234, n, desc_bits)
        return q, noisy_copy_descriptors(42, q, desc_bits)
    raise ValueError(f"unknown distribution {dist!r}")
