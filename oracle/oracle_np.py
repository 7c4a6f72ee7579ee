"""Pure-Python / numpy restatements of the reference matcher -- TEST INFRASTRUCTURE ONLY.

Independent of oracle/pgm_oracle.c so the two can check each other:

* ``match_literal_py``  follows dotnet_src/ImageProcessing/KeypointMatching.cs:14-69
  with Python ints standing in for BigInteger, dict-of-dicts for the distance
  table and ascending lists for the two HashSets (small n only).
* ``match_literal_np``  same loop with a masked row-major argmin
  (row-major first minimum == ascending (i, j) scan with strict '<').
* ``hamming_py``        python_src/photogrammetry/image_processing/keypoint_matching.py:38-40.
"""
from __future__ import annotations

import numpy as np

INT_MAX = 2147483647


def desc_to_int(row) -> int:
    """uint8[stride] little-endian bytes -> non-negative int (the BigInteger)."""
    return int.from_bytes(bytes(bytearray(np.asarray(row, dtype=np.uint8))), "little")


def int_to_desc(value: int, stride: int = 32) -> np.ndarray:
    return np.frombuffer(int(value).to_bytes(stride, "little"), dtype=np.uint8).copy()


def ints_to_desc(values, stride: int = 32) -> np.ndarray:
    out = np.zeros((len(values), stride), dtype=np.uint8)
    for k, v in enumerate(values):
        out[k] = int_to_desc(v, stride)
    return out


def count_ones_py(value: int) -> int:
    """KeypointMatching.cs:71-82."""
    n = 0
    while value != 0:
        value &= value - 1
        n += 1
    return n


def hamming_py(a: int, b: int) -> int:
    """keypoint_matching.py:38-40."""
    return bin(a ^ b).count("1")


def match_literal_py(desc1, desc2):
    """KeypointMatching.cs:14-69 on lists of Python ints; returns [(i, j, dist)] * len(desc1)."""
    table = {}
    for i, a in enumerate(desc1):                       # :20-31
        table[i] = {j: count_ones_py(a ^ b) for j, b in enumerate(desc2)}
    pairs = []
    avail1 = list(range(len(desc1)))                    # :35 (HashSet enumerates ascending)
    avail2 = list(range(len(desc2)))                    # :36
    while len(pairs) < len(desc1):                      # :38
        smallest, si, sj = INT_MAX, 0, 0                # :40-42
        for i in avail1:                                # :44
            row = table[i]
            for j in avail2:                            # :47
                if smallest <= row[j]:                  # :49
                    continue
                smallest, si, sj = row[j], i, j         # :51-53
        if not desc2:
            raise IndexError("keypoints2[0] on an empty list (KeypointMatching.cs:61)")
        pairs.append((si, sj, smallest))                # :57-62
        if si in avail1:
            avail1.remove(si)                           # :64
        if sj in avail2:
            avail2.remove(sj)                           # :65
    return pairs


def distance_matrix_np(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    x = q[:, None, :] ^ t[None, :, :]
    return np.bitwise_count(x).sum(axis=2, dtype=np.int32)


def match_literal_np(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """int32[n1,3]; masked row-major argmin per iteration, with the n1>n2 tail."""
    n1, n2 = len(q), len(t)
    out = np.zeros((n1, 3), dtype=np.int32)
    if n1 == 0:
        return out
    if n2 == 0:
        raise IndexError("keypoints2[0] on an empty list (KeypointMatching.cs:61)")
    big = np.int32(1 << 30)
    d = distance_matrix_np(q, t)
    work = d.copy()
    for k in range(n1):
        if k >= n2:
            out[k] = (0, 0, INT_MAX)
            continue
        flat = int(np.argmin(work))          # first minimum in row-major order
        i, j = divmod(flat, n2)
        out[k] = (i, j, d[i, j])
        work[i, :] = big
        work[:, j] = big
    return out
