#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the reference tree.

Run in the build container only (``/root/reference`` does not exist on the GPU
box; nothing in tests/, smoke() or bench.py reads it at run time):

    python tests/golden/make_golden.py [--reference /root/reference]

What it pins, and with what:

* ``lego_left.npz`` / ``lego_right.npz`` -- the two frozen keypoint sets the
  reference ships (data/feature_matching_test/lego_space_1_from_{left,right}_keypoints.dat),
  repacked as uint8[n,32] descriptor rows + int32[n,2] coordinates.
* ``lego_distances.npz`` -- Hamming distances computed BY THE REFERENCE'S OWN
  ``hamming_distance`` (python_src/photogrammetry/image_processing/keypoint_matching.py:38-40,
  imported from the reference tree): the first 64 rows of the 2175x1285
  matrix, the per-row minimum/first-argmin of every row, and a sha256 of the
  whole int32 matrix.
* ``lego_python_twin.npz`` -- the output of the reference's own
  ``match_keypoints`` (keypoint_matching.py:7-33) on the first 48x40 block.
* ``lego_l2r_expected.npy`` / ``lego_r2l_expected.npy`` -- the greedy
  assignment (C#-only, cannot be executed here) from the numpy restatement
  ``oracle/oracle_np.match_literal_np``, cross-checked in this script against
  the pure-Python dict/BigInt restatement on a sub-block and against the
  anchor triples recorded in SURVEY.md section 4.4.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    ref = args.reference

    from oracle import oracle_np as onp
    from photogrammetry_b200.descriptors import pack_descriptors
    from photogrammetry_b200.keypoint_cache import load_keypoint_dat

    sys.path.insert(0, os.path.join(ref, "python_src"))
    from photogrammetry.image_processing import keypoint_matching as ref_km  # the reference's own code

    sets = {}
    for side in ("left", "right"):
        kps = load_keypoint_dat(os.path.join(ref, "data", "feature_matching_test",
                                             f"lego_space_1_from_{side}_keypoints.dat"))
        ints = [k.BriefDescriptor for k in kps]
        desc = pack_descriptors(ints, 256)
        coord = np.array([k.coord for k in kps], dtype=np.int32)
        np.savez(os.path.join(HERE, f"lego_{side}.npz"), desc=desc, coord=coord)
        sets[side] = (ints, desc)
        print(side, desc.shape, "mean popcount", float(np.bitwise_count(desc).sum(axis=1).mean()))

    li, ld = sets["left"]
    ri, rd = sets["right"]

    # --- distances by the reference's hamming_distance -------------------------
    t0 = time.time()
    full = np.empty((len(li), len(ri)), dtype=np.int32)
    for i, a in enumerate(li):
        full[i] = [ref_km.hamming_distance(a, b) for b in ri]
    print("reference hamming_distance over %dx%d: %.1fs" % (*full.shape, time.time() - t0))
    assert (full == onp.distance_matrix_np(ld, rd)).all(), "numpy restatement disagrees with the reference"
    np.savez(os.path.join(HERE, "lego_distances.npz"),
             first_rows=full[:64].astype(np.int16),
             row_min=full.min(axis=1).astype(np.int16),
             row_argmin=full.argmin(axis=1).astype(np.int32),
             col_min=full.min(axis=0).astype(np.int16),
             col_argmin=full.argmin(axis=0).astype(np.int32),
             sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(full, dtype="<i4").tobytes()).digest(),
                                  dtype=np.uint8))
    print("distance matrix: min %d mean %.1f max %d" % (full.min(), full.mean(), full.max()))

    # --- the reference's match_keypoints on a sub-block ------------------------
    class _KP:  # match_keypoints only reads `.descriptor` (keypoint_matching.py:10-12)
        def __init__(self, d):
            self.descriptor = d

    twin = ref_km.match_keypoints([_KP(d) for d in li[:48]], [_KP(d) for d in ri[:40]], -1)
    np.savez(os.path.join(HERE, "lego_python_twin.npz"), rows=np.asarray(twin, dtype=np.int64))

    # --- greedy assignment from the independent restatements -------------------
    small = onp.match_literal_py(li[:90], ri[:60])
    assert (np.array(small, dtype=np.int32) == onp.match_literal_np(ld[:90], rd[:60])).all()
    t0 = time.time()
    l2r = onp.match_literal_np(ld, rd)
    r2l = onp.match_literal_np(rd, ld)
    print("numpy literal restatement both directions: %.1fs" % (time.time() - t0))
    # anchors recorded in SURVEY.md section 4.4 / BASELINE.md section 4
    assert l2r[:5].tolist() == [[336, 108, 89], [685, 255, 89], [185, 453, 91], [612, 1066, 91], [880, 1278, 91]]
    assert l2r[1284].tolist() == [1407, 408, 108]
    assert (l2r[1285:] == np.array([0, 0, 2147483647])).all() and len(l2r) - 1285 == 890
    np.save(os.path.join(HERE, "lego_l2r_expected.npy"), l2r)
    np.save(os.path.join(HERE, "lego_r2l_expected.npy"), r2l)
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    main()
