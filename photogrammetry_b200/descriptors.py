"""Descriptor marshalling between the reference's integer descriptors and the
packed byte rows the CUDA matcher consumes.

Reference format: ``Keypoint.BriefDescriptor`` is a non-negative
``System.Numerics.BigInteger`` < 2**W with W = ``NumGaussianPairs`` (256)
(dotnet_src/ImageProcessing.Abstractions/Keypoint.cs:14,29-57); the Python
generation uses a plain ``int`` (python_src/photogrammetry/models/keypoint.py:32-50).

Device format: ``uint8[n][stride]`` = the little-endian bytes of that integer
(what ``BigInteger.ToByteArray(isUnsigned: true, isBigEndian: false)`` yields),
zero padded to ``stride`` bytes, ``stride`` a multiple of 16 so every row is
one or more aligned 128-bit loads.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

MAX_DESC_BITS = 512


def stride_for_bits(desc_bits: int) -> int:
    """Smallest multiple of 16 bytes that holds ``desc_bits`` bits."""
    if desc_bits <= 0 or desc_bits > MAX_DESC_BITS:
        raise ValueError(f"desc_bits must be in 1..{MAX_DESC_BITS}, got {desc_bits}")
    return ((desc_bits + 127) // 128) * 16


def pack_descriptors(values: Sequence[int] | Iterable[int], desc_bits: int = 256) -> np.ndarray:
    """Python ints (BigInteger values) -> ``uint8[n, stride]`` rows."""
    stride = stride_for_bits(desc_bits)
    values = list(values)
    buf = bytearray(len(values) * stride)
    limit = 1 << desc_bits
    for k, v in enumerate(values):
        v = int(v)
        if v < 0 or v >= limit:
            raise ValueError(f"descriptor {k} does not fit in {desc_bits} unsigned bits")
        buf[k * stride:(k + 1) * stride] = v.to_bytes(stride, "little")
    return np.frombuffer(bytes(buf), dtype=np.uint8).reshape(len(values), stride).copy()


def unpack_descriptors(rows: np.ndarray) -> list[int]:
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    return [int.from_bytes(r.tobytes(), "little") for r in rows]


def as_descriptor_rows(desc, desc_bits: int | None = None) -> tuple[np.ndarray, int]:
    """Accept ``uint8[n, stride]`` rows or a sequence of ints; return (rows, desc_bits)."""
    if isinstance(desc, np.ndarray) and desc.dtype == np.uint8 and desc.ndim == 2:
        rows = np.ascontiguousarray(desc)
        bits = desc_bits or rows.shape[1] * 8
        if rows.shape[1] % 16 != 0:
            padded = np.zeros((rows.shape[0], stride_for_bits(max(bits, rows.shape[1] * 8))), dtype=np.uint8)
            padded[:, :rows.shape[1]] = rows
            rows = padded
        return rows, bits
    bits = desc_bits or 256
    return pack_descriptors(desc, bits), bits
