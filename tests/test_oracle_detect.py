"""The detection oracle (oracle/detect_np.py) against what the reference itself pins (CPU only).

* the reference's three xUnit facts for FAST (ImageProcessing.Tests/KeypointDetectionTests.cs:10-50);
* the golden vectors produced by the reference's own Python detector and BRIEF code
  (tests/golden/make_golden_detect.py -> star_detect.npz);
* host-side pieces of the product that need no GPU (pair tables, grayscale conversion)."""
import os

import numpy as np
import pytest

from oracle import detect_np as D
from photogrammetry_b200 import keypoint_detection as kd
from photogrammetry_b200.descriptors import pack_descriptors

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def star():
    return np.load(os.path.join(GOLDEN, "star_detect.npz"))


# ---- KeypointDetectionTests.cs -------------------------------------------------------------
def test_xunit_dim_center_is_detected():
    img = np.zeros((7, 7), dtype=np.float32)          # img[y, x]; the C# indexer is [x, y]
    for x, y in ((3, 0), (0, 3), (3, 6), (6, 3)):
        img[y, x] = 1.0
    assert D.is_potential_keypoint(img, np.float32(0), 3, 3, 0.5) is True          # :10-28


def test_xunit_consistent_brightness_is_not_detected():
    img = np.full((7, 7), 0.5, dtype=np.float32)
    assert D.is_potential_keypoint(img, np.float32(0.5), 3, 3, 0.5) is False       # :30-40
    assert D.intensity_value_if_keypoint(img, 3, 3, 0.5) is None                   # :42-50
    assert len(D.detect(img, 0.5)[0]) == 0 and len(D.detect_vectorised(img, 0.5)[0]) == 0


# ---- golden vectors from the reference's Python generation ---------------------------------
@pytest.mark.parametrize("k", [0, 1])
def test_python_generation_detector_matches_reference_output(star, k):
    gray = star[f"gray{k}"].astype(np.float32)
    xy, _ = D.detect_vectorised(gray, 50, python_generation=True)
    assert (xy[:, ::-1] == star[f"uv{k}"]).all() and len(xy) == len(star[f"uv{k}"])
    pairs = D.py_pairs_to_xy(star["pairs"])
    ints = D.brief_descriptors(gray, xy, pairs, lsb_first=True)
    assert (pack_descriptors(ints, 256) == star[f"desc{k}"]).all()


@pytest.fixture(scope="module")
def lego_crop():
    return np.load(os.path.join(GOLDEN, "lego_crop_detect.npz"))


@pytest.mark.parametrize("k", [0, 1])
def test_python_generation_detector_matches_reference_output_on_a_natural_image(lego_crop, k):
    # tests/golden/make_golden_detect_lego.py: the reference's own detector and BRIEF on a 720 x 960 crop of the
    # lego photographs it ships (2101 / 2141 keypoints at threshold 25)
    g = lego_crop
    gray = g[f"gray{k}"].astype(np.float32)
    xy, _ = D.detect_vectorised(gray, float(g["threshold"]), python_generation=True)
    assert len(xy) == len(g[f"uv{k}"]) and (xy[:, ::-1] == g[f"uv{k}"]).all()
    pairs = D.py_pairs_to_xy(g["pairs"])
    ints = D.brief_descriptors(gray, xy, pairs, lsb_first=True)
    assert (pack_descriptors(ints, 256) == g[f"desc{k}"]).all()


def test_python_twin_distances_match_reference_output_on_a_natural_image(lego_crop):
    # match_keypoints of the reference (keypoint_matching.py:7-33) on those descriptors: every row's sorted distances
    from oracle import orc
    rows = orc.python_twin(lego_crop["desc0"], lego_crop["desc1"])
    assert (rows[:, :, 1] == lego_crop["twin_sorted_dists"]).all()
    assert (rows[:, 0, 1] == lego_crop["twin_nearest_dist"]).all()


def test_python_pixel_test_equals_the_shared_segment_test(star):
    gray = star["gray0"].astype(np.float32)
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, size=(40, 50)).astype(np.float32)
    for img, th in ((gray[150:230, 180:280], 50.0), (noise, 90.0), (noise, 120.0)):
        h, w = img.shape
        for y in range(3, h - 3):
            for x in range(3, w - 3):
                a = D.py_is_keypoint(img, x, y, th)
                b = D.intensity_value_if_keypoint(img, x, y, th, python_generation=True) is not None
                assert a == b


@pytest.mark.parametrize("python_generation", [False, True])
def test_scalar_and_vectorised_detectors_agree(star, python_generation):
    rng = np.random.default_rng(11)
    noise = rng.random((37, 61), dtype=np.float32)
    crop = (star["gray0"][120:200, 150:260].astype(np.float32) / np.float32(255))
    for img, th in ((noise, 0.05), (noise, 0.12), (crop, 0.1)):
        a = D.detect(img, th, python_generation)
        b = D.detect_vectorised(img, th, python_generation)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    assert len(D.detect_vectorised(noise, 0.12, python_generation)[0]) > 0


def test_ring_typo_changes_results(star):
    # the C# table repeats {-3, 1} (KeypointDetection.cs:18): a pixel pattern that tells the two tables apart
    img = np.zeros((7, 7), dtype=np.float32)
    img[3, 3] = 1.0
    img[3 - 1, 3 - 3] = 1.0          # (dx, dy) = (-3, -1): on the Python ring only
    img[3 + 3, 3 + 0] = 1.0
    img[3 + 3, 3 + 1] = 1.0
    img[3 + 3, 3 - 1] = 1.0
    img[3 + 2, 3 + 2] = 1.0          # five inside-threshold ring pixels for Python, four for the C# table
    assert D.intensity_value_if_keypoint(img, 3, 3, 0.5, python_generation=True) is None
    assert D.intensity_value_if_keypoint(img, 3, 3, 0.5, python_generation=False) == 12


# ---- NMS -------------------------------------------------------------------------------------
def test_nms_known_answer_and_properties():
    coords = np.array([[0, 0], [3, 4], [10, 0], [0, 5], [100, 100], [6, 8]], dtype=np.int32)
    scores = np.array([12, 16, 12, 16, 12, 13], dtype=np.int32)
    # order by score desc, stable: 1, 3, 5, 0, 2, 4.  radius 5: 1 kept; 3 is sqrt(9+1) from 1 -> dropped;
    # 5 is exactly 5 from 1 -> dropped (<= radius); 0 exactly 5 -> dropped; 2 is sqrt(49+16) -> kept; 4 kept
    assert D.eliminate_redundant(coords, scores, 5).tolist() == [1, 2, 4]
    assert D.eliminate_redundant(coords, scores, 0).tolist() == [1, 3, 5, 0, 2, 4]
    rng = np.random.default_rng(5)
    c = rng.integers(0, 200, size=(400, 2)).astype(np.int32)
    s = rng.integers(12, 17, size=400).astype(np.int32)
    kept = D.eliminate_redundant(c, s, 20)
    kc = c[kept].astype(np.int64)
    d2 = ((kc[:, None, :] - kc[None, :, :]) ** 2).sum(-1)
    assert (d2[~np.eye(len(kept), dtype=bool)] > 400).all()
    assert (np.diff(s[kept]) <= 0).all()
    # the vectorised form used at large sizes is the same function
    for n, radius, span in [(400, 20, 200), (900, 3, 60), (300, 0, 10), (500, 1000, 100)]:
        c = rng.integers(-span, span, size=(n, 2)).astype(np.int32)
        s = rng.integers(-2, 17, size=n).astype(np.int32)
        assert D.eliminate_redundant_vectorised(c, s, radius).tolist() == D.eliminate_redundant(c, s, radius).tolist()


# ---- host-side product logic (no GPU) ---------------------------------------------------------
def test_product_pair_tables_match(star):
    np.random.seed(20231018)                      # the seed make_golden_detect.py gave the reference
    assert (kd.generate_gaussian_pairs(stdev=50) == star["pairs"]).all()
    u = kd.Utils(99)
    mine = np.array([[[a.X, a.Y], [b.X, b.Y]] for a, b in (u.NextGaussianPair(50) for _ in range(64))])
    assert (mine == D.gaussian_pairs(99, 64, 50)).all()
    assert mine.min() >= 0 and mine.max() > 50        # upstream's sampler only yields non-negative offsets


def test_product_grayscale_matches_oracle():
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, size=(9, 13, 4), dtype=np.uint8)
    a, b = kd.grayscale_from_rgb8(rgb), D.grayscale_from_rgb8(rgb)
    assert a.dtype == np.float32 and (a == b).all()
    assert a.max() <= 1.0 and kd.grayscale_from_rgb8(np.full((1, 1, 3), 255, np.uint8))[0, 0] == 1.0
