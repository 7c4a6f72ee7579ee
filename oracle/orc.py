"""ctypes binding of oracle/liborc.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module (see pgm_oracle.h).  The
product package ``photogrammetry_b200`` never does.

Each wrapper takes/returns numpy arrays; descriptors are ``uint8[n, stride]``
(little-endian bytes of the BigInteger, zero padded).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

ORC_OK = 0
ORC_E_INVALID_ARG = -1
ORC_E_NOMEM = -2
ORC_E_EMPTY_TRAIN = -5
TAIL_DISTANCE = 2147483647


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "pgm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        u8p, i32p, i64p, f32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float))
        _lib.orc_count_ones_kernighan.argtypes = [u8p, u8p, C.c_int]
        _lib.orc_hamming.argtypes = [u8p, u8p, C.c_int]
        _lib.orc_distance_matrix.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, i32p]
        _lib.orc_match_literal.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, i32p, i32p, i32p]
        _lib.orc_match_sweep.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p, i32p]
        _lib.orc_match_rounds.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p, i32p, i32p]
        _lib.orc_knn2.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p, i32p, i32p]
        _lib.orc_match_ratio_crosscheck.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                                    i32p, i32p, i32p, i32p]
        _lib.orc_python_twin.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, i64p]
        _lib.orc_l2_knn2.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, i32p, f32p, i32p, f32p]
        _lib.orc_gen_uniform.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, u8p]
        _lib.orc_gen_uniform.restype = None
        _lib.orc_gen_noisy_copy.argtypes = [C.c_uint64, u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, u8p]
        _lib.orc_gen_noisy_copy.restype = None
        _lib.orc_num_threads.argtypes = []
    return _lib


class EmptyTrainError(IndexError):
    """The reference throws ArgumentOutOfRangeException at KeypointMatching.cs:61."""


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i32(n):
    a = np.empty(max(int(n), 0), dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def _check(rc):
    if rc == ORC_E_EMPTY_TRAIN:
        raise EmptyTrainError("keypoints2 is empty")
    if rc != ORC_OK:
        raise RuntimeError(f"oracle error {rc}")


def _shape(q, t):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    stride = q.shape[1] if q.ndim == 2 and q.shape[0] else (t.shape[1] if t.ndim == 2 else 32)
    return q, t, int(q.shape[0]), int(t.shape[0]), int(stride)


def hamming(a, b, kernighan=False):
    a, pa = _u8(a)
    b, pb = _u8(b)
    f = lib().orc_count_ones_kernighan if kernighan else lib().orc_hamming
    return int(f(pa, pb, int(a.size)))


def distance_matrix(q, t, kernighan=False):
    q, t, n1, n2, stride = _shape(q, t)
    out = np.empty((n1, n2), dtype=np.int32)
    _check(lib().orc_distance_matrix(_u8(q)[1], n1, _u8(t)[1], n2, stride, int(kernighan),
                                     out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out


def _match(fn, q, t, *extra):
    q, t, n1, n2, stride = _shape(q, t)
    (qi, pqi), (tj, ptj), (dd, pdd) = _i32(n1), _i32(n1), _i32(n1)
    _check(fn(_u8(q)[1], n1, _u8(t)[1], n2, stride, *extra, pqi, ptj, pdd))
    return np.stack([qi, tj, dd], axis=1)


def match_literal(q, t, kernighan=True):
    """int32[n1, 3] rows (qi, tj, dist) in the reference's output order."""
    return _match(lib().orc_match_literal, q, t, int(kernighan))


def match_sweep(q, t):
    return _match(lib().orc_match_sweep, q, t)


def match_rounds(q, t, return_rounds=False):
    q, t, n1, n2, stride = _shape(q, t)
    (qi, pqi), (tj, ptj), (dd, pdd) = _i32(n1), _i32(n1), _i32(n1)
    rounds = C.c_int32(0)
    _check(lib().orc_match_rounds(_u8(q)[1], n1, _u8(t)[1], n2, stride, pqi, ptj, pdd, C.byref(rounds)))
    out = np.stack([qi, tj, dd], axis=1)
    return (out, rounds.value) if return_rounds else out


def knn2(q, t):
    q, t, n1, n2, stride = _shape(q, t)
    (bj, pbj), (bd, pbd), (sj, psj), (sd, psd) = _i32(n1), _i32(n1), _i32(n1), _i32(n1)
    _check(lib().orc_knn2(_u8(q)[1], n1, _u8(t)[1], n2, stride, pbj, pbd, psj, psd))
    return bj, bd, sj, sd


def match_ratio_crosscheck(q, t, ratio=0.8, cross_check=True, max_dist=-1):
    q, t, n1, n2, stride = _shape(q, t)
    (qi, pqi), (tj, ptj), (dd, pdd) = _i32(n1), _i32(n1), _i32(n1)
    cnt = C.c_int32(0)
    _check(lib().orc_match_ratio_crosscheck(_u8(q)[1], n1, _u8(t)[1], n2, stride, float(ratio), int(cross_check),
                                            int(max_dist), pqi, ptj, pdd, C.byref(cnt)))
    c = cnt.value
    return np.stack([qi[:c], tj[:c], dd[:c]], axis=1)


def python_twin(q, t):
    q, t, n1, n2, stride = _shape(q, t)
    out = np.empty((n1, n2, 2), dtype=np.int64)
    _check(lib().orc_python_twin(_u8(q)[1], n1, _u8(t)[1], n2, stride, out.ctypes.data_as(C.POINTER(C.c_int64))))
    return out


def l2_knn2(q, t):
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    n1, n2, dim = q.shape[0], t.shape[0], (q.shape[1] if q.shape[0] else t.shape[1])
    (bj, pbj), (sj, psj) = _i32(n1), _i32(n1)
    bd = np.empty(n1, dtype=np.float32)
    sd = np.empty(n1, dtype=np.float32)
    fp = C.POINTER(C.c_float)
    _check(lib().orc_l2_knn2(q.ctypes.data_as(fp), n1, t.ctypes.data_as(fp), n2, dim,
                             pbj, bd.ctypes.data_as(fp), psj, sd.ctypes.data_as(fp)))
    return bj, bd, sj, sd


def gen_uniform(seed, n, desc_bits=256, stride=None):
    stride = stride or ((desc_bits + 127) // 128) * 16
    out = np.zeros((n, stride), dtype=np.uint8)
    lib().orc_gen_uniform(C.c_uint64(seed), n, desc_bits, stride, _u8(out)[1]) if n else None
    return out


def gen_noisy_copy(seed, query, desc_bits=256, flip_p=0.10, outlier_p=0.30):
    query = np.ascontiguousarray(query, dtype=np.uint8)
    n, stride = query.shape
    out = np.zeros((n, stride), dtype=np.uint8)
    if n:
        lib().orc_gen_noisy_copy(C.c_uint64(seed), _u8(query)[1], n, desc_bits, stride, flip_p, outlier_p, _u8(out)[1])
    return out


def num_threads():
    return int(lib().orc_num_threads())
