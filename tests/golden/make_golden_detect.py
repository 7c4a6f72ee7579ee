#!/usr/bin/env python
"""Golden vectors for the keypoint producer (FAST + BRIEF), made BY THE REFERENCE'S OWN PYTHON CODE.

Run in the build container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden_detect.py [--reference /root/reference]

Writes ``tests/golden/star_detect.npz``:

* ``gray0`` / ``gray1`` -- uint8[383, 451]: ``ImageDB.get_bw_image`` (storage/image_db.py:32-37) of
  data/feature_matching_test/15pt_star.png and 15pt_star_shifted_150.png (the pair
  scripts/match_keypoints.py is run on upstream);
* ``pairs`` -- int64[256, 2, 2]: ``generate_gaussian_pairs(stdev=50)`` (models/keypoint.py:52-57) after
  ``np.random.seed(20231018)`` (the reference never seeds; the table must be frozen to be reproducible);
* ``uv0`` / ``uv1`` -- int32[n, 2]: keypoint coordinates (row, column) returned by
  ``FASTKeypointDetector(50, image_db).detect_points`` (image_processing/keypoint_detection.py:147-175);
* ``desc0`` / ``desc1`` -- uint8[n, 32]: ``KeyPoint.descriptor`` (models/keypoint.py:32-50) of each, packed
  little-endian;
* ``twin_nearest`` -- int64[n0, 2]: column 0 of ``match_keypoints(k0, k1)`` (keypoint_matching.py:7-33),
  i.e. nearest neighbour distance per keypoint (index dropped where the minimum is tied: the
  reference's argsort is not stable).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    ref = args.reference
    sys.path.insert(0, os.path.join(ref, "python_src"))
    import cv2
    from photogrammetry.image_processing.keypoint_detection import FASTKeypointDetector   # the reference's own
    from photogrammetry.image_processing.keypoint_matching import match_keypoints
    from photogrammetry.models import keypoint as ref_kp
    from photogrammetry.storage.image_db import ImageDB

    from photogrammetry_b200.descriptors import pack_descriptors

    data = os.path.join(ref, "data", "feature_matching_test")
    imgs = [cv2.imread(os.path.join(data, f)) for f in ("15pt_star.png", "15pt_star_shifted_150.png")]
    h, w = imgs[0].shape[:2]
    db = ImageDB(h, w)
    ids = [db.add_image(im) for im in imgs]

    np.random.seed(20231018)
    det = FASTKeypointDetector(50, db)                  # draws the pair table from the seeded global generator
    pairs = np.array(det._gaussian_pairs, dtype=np.int64)
    out = {"pairs": pairs}
    kps = []
    for k, i in enumerate(ids):
        kp = det.detect_points(i)
        kps.append(kp)
        out[f"gray{k}"] = db.get_bw_image(i).astype(np.uint8)
        assert (out[f"gray{k}"].astype(np.int16) == db.get_bw_image(i)).all()
        out[f"uv{k}"] = np.array([p.coord for p in kp], dtype=np.int32).reshape(-1, 2)
        out[f"desc{k}"] = pack_descriptors([int(p.descriptor) for p in kp], 256)
        print(f"image {k}: {len(kp)} keypoints")
    twin = match_keypoints(kps[0], kps[1], -1)
    out["twin_nearest_dist"] = twin[:, 0, 1].astype(np.int64)
    out["twin_sorted_dists"] = twin[:, :, 1].astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "star_detect.npz"), **out)
    print("wrote star_detect.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
