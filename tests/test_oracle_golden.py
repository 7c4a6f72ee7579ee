"""The oracle against the golden vectors (CPU).  See tests/golden/make_golden.py
for what each fixture pins and where it came from."""
import hashlib

import numpy as np
import pytest

from oracle import oracle_np as onp
from oracle import orc


def test_distances_match_reference_hamming_distance(lego):
    d = orc.distance_matrix(lego["left"], lego["right"])
    g = lego["dist"]
    assert d.shape == (2175, 1285)
    assert hashlib.sha256(np.ascontiguousarray(d, dtype="<i4").tobytes()).digest() == g["sha256"].tobytes()
    assert (d[:64] == g["first_rows"]).all()
    assert (d.min(axis=1) == g["row_min"]).all() and (d.argmin(axis=1) == g["row_argmin"]).all()
    assert (d.min(axis=0) == g["col_min"]).all() and (d.argmin(axis=0) == g["col_argmin"]).all()
    assert d.min() == 89 and d.max() == 172


def test_kernighan_count_ones_equals_popcount(lego):
    # CountOnes (KeypointMatching.cs:71-82) restated literally vs. the hardware popcount
    d1 = orc.distance_matrix(lego["left"][:40], lego["right"][:50], kernighan=True)
    d2 = orc.distance_matrix(lego["left"][:40], lego["right"][:50], kernighan=False)
    assert (d1 == d2).all() and (d1 == lego["dist"]["first_rows"][:40, :50]).all()


def test_python_twin_matches_reference_match_keypoints(lego):
    got = orc.python_twin(lego["left"][:48], lego["right"][:40])
    ref = lego["twin"]
    assert got.shape == ref.shape == (48, 40, 2)
    # distances per rank are pinned; the index order among equal distances is
    # unspecified upstream (numpy's default argsort is not stable)
    assert (got[:, :, 1] == ref[:, :, 1]).all()
    for i in range(48):
        assert sorted(got[i, :, 0].tolist()) == list(range(40)) == sorted(ref[i, :, 0].tolist())
        for d in np.unique(ref[i, :, 1]):
            assert set(got[i, got[i, :, 1] == d, 0]) == set(ref[i, ref[i, :, 1] == d, 0])


@pytest.mark.parametrize("direction", ["l2r", "r2l"])
def test_greedy_assignment_all_formulations(lego, direction):
    q, t = (lego["left"], lego["right"]) if direction == "l2r" else (lego["right"], lego["left"])
    exp = lego[direction]
    assert (orc.match_sweep(q, t) == exp).all()
    assert (orc.match_rounds(q, t) == exp).all()
    assert (orc.match_literal(q, t, kernighan=False) == exp).all()


def test_literal_with_kernighan_on_subblock(lego):
    q, t = lego["left"][:300], lego["right"][:200]
    a = orc.match_literal(q, t, kernighan=True)
    assert (a == orc.match_sweep(q, t)).all()
    assert (a == onp.match_literal_np(q, t)).all()


def test_anchor_triples(lego):
    # SURVEY.md section 4.4 / BASELINE.md section 4
    got = orc.match_rounds(lego["left"], lego["right"])
    assert got[:5].tolist() == [[336, 108, 89], [685, 255, 89], [185, 453, 91], [612, 1066, 91], [880, 1278, 91]]
    assert got[1284].tolist() == [1407, 408, 108]
    assert (got[1285:] == [0, 0, 2147483647]).all() and len(got) == 2175
    assert (np.diff(got[:1285, 2]) >= 0).all()
