"""include/pgmatch.hpp -- the C++ host-side mirror of ImageProcessing.KeypointMatching over the C ABI.

A C++17 program written like a caller of the reference's C# class (tests/cpp/consumer.cpp) is compiled with
-Wall -Wextra -Werror -pedantic, linked against libpgmatch.so and run on the frozen lego descriptors; the expected
triples come from the oracle.  Without a GPU the constructor must refuse (PGM_E_NO_DEVICE), never fall back."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from oracle import orc
from photogrammetry_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    cxx = shutil.which("g++") or shutil.which("c++")
    if not cxx:
        pytest.skip("no C++ compiler")
    libdir = os.path.dirname(_lib.LIB_PATH)
    _lib.load()                                   # builds the library if it is missing
    exe = tmp_path / "consumer"
    subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "consumer.cpp"), "-o", str(exe), "-L", libdir, "-lpgmatch",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


def _case(tmp_path, q, t):
    exp = orc.match_literal(q, t).astype(np.int32)
    p = tmp_path / "case.bin"
    with open(p, "wb") as f:
        f.write(struct.pack("<iii", len(q), len(t), q.shape[1]))
        f.write(np.ascontiguousarray(q).tobytes()); f.write(np.ascontiguousarray(t).tobytes())
        f.write(struct.pack("<i", len(exp))); f.write(np.ascontiguousarray(exp).tobytes())
    return p


def _lego(n1, n2):
    g = os.path.join(ROOT, "tests", "golden")
    return np.load(os.path.join(g, "lego_left.npz"))["desc"][:n1], np.load(os.path.join(g, "lego_right.npz"))["desc"][:n2]


def test_cpp_mirror_compiles_and_refuses_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    out = subprocess.run([str(exe), str(_case(tmp_path, *_lego(60, 40)))], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert (out.returncode, out.stdout.strip()) == (0, "ok"), (out.stdout, out.stderr)
    else:
        assert (out.returncode, out.stdout.strip()) == (2, "no-device"), (out.stdout, out.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("n1,n2", [(400, 250), (250, 400), (1, 1)])
def test_cpp_mirror_matches_reference_semantics(tmp_path, n1, n2):
    exe = _build(tmp_path)
    out = subprocess.run([str(exe), str(_case(tmp_path, *_lego(n1, n2)))], capture_output=True, text=True, timeout=300)
    assert (out.returncode, out.stdout.strip()) == (0, "ok"), (out.stdout, out.stderr)
