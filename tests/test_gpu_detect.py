"""FAST-12 / BRIEF / NMS kernels against the oracle and the reference's own golden vectors, through the
C ABI (pgm_fast_detect, pgm_brief_describe, pgm_nms).  Bit-exact: coordinates, order, scores, descriptor
bytes, survivor order."""
import os

import numpy as np
import pytest

from oracle import detect_np as D
from oracle import orc
from photogrammetry_b200 import keypoint_detection as kd
from photogrammetry_b200.descriptors import pack_descriptors
from photogrammetry_b200.keypoint_matching import KeypointMatching, match_keypoints

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def star():
    return np.load(os.path.join(GOLDEN, "star_detect.npz"))


# ---- the reference's own outputs (Python generation) -------------------------------------------
@pytest.mark.parametrize("k", [0, 1])
def test_python_generation_golden(matcher, star, k):
    det = kd.FASTKeypointDetector(50, star[f"gray{k}"].astype(np.int16), gaussian_pairs=star["pairs"], matcher=matcher)
    uv, desc = det.detect_arrays()
    assert uv.shape == star[f"uv{k}"].shape and (uv == star[f"uv{k}"]).all()
    assert (desc == star[f"desc{k}"]).all()
    pts = det.detect_points()
    assert [p.coord.tolist() for p in pts] == star[f"uv{k}"].tolist()
    assert pack_descriptors([p.descriptor for p in pts], 256).tobytes() == star[f"desc{k}"].tobytes()


def test_python_generation_detect_then_match_golden(matcher, star):
    k0 = kd.FASTKeypointDetector(50, star["gray0"].astype(np.int16), star["pairs"], matcher=matcher).detect_points()
    k1 = kd.FASTKeypointDetector(50, star["gray1"].astype(np.int16), star["pairs"], matcher=matcher).detect_points()
    rows = match_keypoints(k0, k1)
    assert rows.shape == (128, 100, 2)
    assert (rows[:, :, 1] == star["twin_sorted_dists"]).all()           # the reference's match_keypoints output
    assert (rows[:, 0, 1] == star["twin_nearest_dist"]).all()


@pytest.mark.parametrize("k", [0, 1])
def test_python_generation_golden_natural_image(matcher, k):
    # the second golden: the reference's own detector + BRIEF on a crop of its lego photographs (2101 / 2141 keypoints)
    g = np.load(os.path.join(GOLDEN, "lego_crop_detect.npz"))
    det = kd.FASTKeypointDetector(int(g["threshold"]), g[f"gray{k}"].astype(np.int16), gaussian_pairs=g["pairs"], matcher=matcher)
    uv, desc = det.detect_arrays()
    assert uv.shape == g[f"uv{k}"].shape and (uv == g[f"uv{k}"]).all()
    assert (desc == g[f"desc{k}"]).all()
    if k == 0:
        rows = matcher.match_keypoints_sorted(g["desc0"], g["desc1"], 256)
        assert (rows[:, :, 1] == g["twin_sorted_dists"]).all()


# ---- C# generation vs the oracle -----------------------------------------------------------------
def _images(star):
    rng = np.random.default_rng(2024)
    out = [(kd.grayscale_from_rgb8(np.repeat(star["gray0"][:, :, None], 3, axis=2)), 0.1),
           (kd.grayscale_from_rgb8(np.repeat(star["gray1"][:, :, None], 3, axis=2)), 0.1),
           (rng.random((97, 131), dtype=np.float32), 0.08),
           (rng.random((64, 33), dtype=np.float32), 0.12),
           (rng.integers(0, 4, size=(150, 150)).astype(np.float32) / np.float32(3), 0.2),   # many exact ties at the bounds
           (rng.random((7, 7), dtype=np.float32), 0.05),
           (rng.random((6, 40), dtype=np.float32), 0.05),            # no interior: empty result
           (rng.random((1, 1), dtype=np.float32), 0.05)]
    return out


@pytest.mark.parametrize("python_generation", [False, True])
def test_fast_detect_against_oracle(matcher, star, python_generation):
    total = 0
    for img, th in _images(star):
        exp_xy, exp_sc = D.detect_vectorised(img, th, python_generation) if min(img.shape) >= 7 else (
            np.zeros((0, 2), np.int32), np.zeros(0, np.int32))
        xy, sc = matcher.fast_detect(img, th, python_generation=python_generation)
        assert xy.shape == exp_xy.shape and (xy == exp_xy).all() and (sc == exp_sc).all()
        total += len(xy)
    assert total > 1000


def test_fast_detect_xunit_facts(matcher):
    # KeypointDetectionTests.cs:42-50: constant brightness yields nothing
    assert len(matcher.fast_detect(np.full((7, 7), 0.5, np.float32), 0.5)[0]) == 0
    # a dim centre among bright compass points AND a bright ring is a keypoint with the full-ring score
    img = np.ones((7, 7), dtype=np.float32)
    img[3, 3] = 0.0
    xy, sc = matcher.fast_detect(img, 0.5)
    assert xy.tolist() == [[3, 3]] and sc.tolist() == [16]


def test_fast_detect_capacity_regrow(matcher):
    rng = np.random.default_rng(1)
    img = rng.random((300, 400), dtype=np.float32)
    xy, sc = matcher.fast_detect(img, 0.02)                 # > 4096 keypoints: the wrapper regrows its buffers
    exp_xy, exp_sc = D.detect_vectorised(img, 0.02)
    assert len(exp_xy) > 4096 and (xy == exp_xy).all() and (sc == exp_sc).all()


@pytest.mark.parametrize("n_pairs", [1, 31, 100, 256, 300, 512])
def test_brief_against_oracle(matcher, n_pairs):
    rng = np.random.default_rng(n_pairs)
    img = rng.integers(0, 6, size=(60, 80)).astype(np.float32)          # ties: strict < matters
    xy = np.stack([rng.integers(0, 80, size=70), rng.integers(0, 60, size=70)], axis=1).astype(np.int32)
    xy[:4] = [[0, 0], [79, 59], [0, 59], [79, 0]]
    pairs = rng.integers(-45, 46, size=(n_pairs, 2, 2)).astype(np.int32)
    for lsb in (False, True):
        exp = pack_descriptors(D.brief_descriptors(img, xy, pairs, lsb_first=lsb), max(n_pairs, 1))
        got = matcher.brief_describe(img, xy, pairs.reshape(-1, 4), python_generation=lsb)
        assert got.shape == exp.shape and (got == exp).all()


def test_brief_upstream_sampler_pairs(matcher, star):
    gray = kd.grayscale_from_rgb8(np.repeat(star["gray0"][:, :, None], 3, axis=2))
    det = kd.KeypointDetection(kd.KeypointDetectionOptions(Threshold=0.1), seed=7, matcher=matcher)
    kps = det.Detect(gray)
    pairs = D.gaussian_pairs(7, 256, 50)
    exp_xy, exp_sc = D.detect_vectorised(gray, 0.1)
    assert [(k.Coordinate.X, k.Coordinate.Y) for k in kps] == [tuple(c) for c in exp_xy.tolist()]
    assert [k.FastScore for k in kps] == exp_sc.tolist()
    assert [k.BriefDescriptor for k in kps] == D.brief_descriptors(gray, exp_xy, pairs)
    assert all(k.Value == gray[k.Coordinate.Y, k.Coordinate.X] for k in kps)


@pytest.mark.parametrize("mode", ["dense", "binned"])
@pytest.mark.parametrize("n,radius,span", [(1, 5, 10), (2, 0, 1), (50, 3, 20), (400, 20, 200), (3000, 50, 4000),
                                           (2500, 7, 60), (700, 1000, 100)])
def test_nms_against_oracle(matcher, monkeypatch, mode, n, radius, span):
    monkeypatch.setenv("PGM_NMS_MODE", mode)
    rng = np.random.default_rng(n + radius)
    xy = rng.integers(0, span, size=(n, 2)).astype(np.int32)
    sc = rng.integers(12, 17, size=n).astype(np.int32)
    exp = D.eliminate_redundant(xy, sc, radius)
    got = matcher.nms(xy, sc, radius)
    assert got.tolist() == exp.tolist()


@pytest.mark.parametrize("n,radius,span,lo", [(30000, 6, 2000, 0), (20000, 1, 500, -250), (30000, 0, 100, 0),
                                              (20000, 40, 2_000_000_000, -1_000_000_000)])
def test_nms_large_default_mode(matcher, n, radius, span, lo):
    """Sizes where the default picks the spatially binned rounds; negative and extreme coordinates, radius 0
    (only exact duplicates suppress each other), many duplicates."""
    rng = np.random.default_rng(n)
    xy = (rng.integers(0, span, size=(n, 2)) + lo).astype(np.int32)
    sc = rng.integers(-3, 17, size=n).astype(np.int32)
    exp = D.eliminate_redundant_vectorised(xy, sc, radius)
    got = matcher.nms(xy, sc, radius)
    assert got.tolist() == exp.tolist()


def test_nms_empty_and_chain(matcher):
    assert matcher.nms(np.zeros((0, 2), np.int32), np.zeros(0, np.int32), 5).tolist() == []
    # a chain where each decision depends on the previous one: worst case for the parallel rounds
    xy = np.stack([np.arange(300) * 3, np.zeros(300)], axis=1).astype(np.int32)
    sc = np.full(300, 12, np.int32)
    assert matcher.nms(xy, sc, 4).tolist() == D.eliminate_redundant(xy, sc, 4).tolist() == list(range(0, 300, 2))


# ---- configs[0]/[2] proxy: image pair -> detector -> NMS -> matcher, all through the product ------------
def test_image_pair_pipeline_against_oracle_chain(matcher, star):
    opts = kd.KeypointDetectionOptions(Threshold=0.1, GaussianStandardDeviation=50, NumGaussianPairs=256)
    det = kd.KeypointDetection(opts, seed=42, matcher=matcher)
    nms = kd.RedundantKeypointEliminator(kd.RedundantKeypointEliminationOptions(SuppressionRadius=10), matcher=matcher)
    pairs = D.gaussian_pairs(42, 256, 50)
    lists, exp_desc = [], []
    for k in (0, 1):
        gray = kd.grayscale_from_rgb8(np.repeat(star[f"gray{k}"][:, :, None], 3, axis=2))
        kept = nms.EliminateRedundantKeypoints(det.Detect(gray))
        exy, esc = D.detect_vectorised(gray, 0.1)
        ek = D.eliminate_redundant(exy, esc, 10)
        assert [(p.Coordinate.X, p.Coordinate.Y) for p in kept] == [tuple(c) for c in exy[ek].tolist()]
        lists.append(kept)
        exp_desc.append(pack_descriptors(D.brief_descriptors(gray, exy[ek], pairs), 256))
    got = KeypointMatching().MatchKeypoints(lists[0], lists[1])
    exp = orc.match_literal(exp_desc[0], exp_desc[1], kernighan=True)
    assert len(got) == len(lists[0]) == len(exp)
    idx0 = {id(p): i for i, p in enumerate(lists[0])}
    idx1 = {id(p): i for i, p in enumerate(lists[1])}
    assert [[idx0[id(p.Keypoint1)], idx1[id(p.Keypoint2)], p.Distance] for p in got] == exp.tolist()


# ---- the same chain with the image and every intermediate on the device ---------------------------------------
@pytest.mark.parametrize("radius", [-1, 0, 10])
@pytest.mark.parametrize("python_generation", [False, True])
def test_device_resident_chain_equals_host_calls(matcher, star, radius, python_generation):
    import torch
    gray = kd.grayscale_from_rgb8(np.repeat(star["gray0"][:, :, None], 3, axis=2))
    pairs = D.gaussian_pairs(5, 256, 50).reshape(-1, 4)
    xy, sc = matcher.fast_detect(gray, 0.1, python_generation)
    if radius >= 0:
        kept = matcher.nms(xy, sc, radius)
        xy, sc = xy[kept], sc[kept]
    desc = matcher.brief_describe(gray, xy, pairs, python_generation=python_generation)
    dev = torch.device("cuda", matcher.device)
    d_gray = torch.from_numpy(np.ascontiguousarray(gray, dtype=np.float32)).to(dev)
    for cap in (8192, 7):                                     # the second forces the capacity regrowth path
        gxy, gsc, gdesc = matcher.detect_describe_dev(d_gray, 0.1, pairs, radius, python_generation=python_generation,
                                                      capacity=cap)
        assert (gxy.cpu().numpy() == xy).all() and (gsc.cpu().numpy() == sc).all()
        assert gdesc.shape == desc.shape and (gdesc.cpu().numpy() == desc).all()
    assert len(xy) > 5


def test_device_resident_chain_into_the_matcher(matcher, star):
    # image pair -> device chain -> pgm_match_hamming_greedy_dev on the device descriptors == oracle chain
    import torch
    pairs = D.gaussian_pairs(9, 256, 50).reshape(-1, 4)
    dev = torch.device("cuda", matcher.device)
    outs, exp_desc = [], []
    for key in ("gray0", "gray1"):
        gray = kd.grayscale_from_rgb8(np.repeat(star[key][:, :, None], 3, axis=2))
        d_gray = torch.from_numpy(np.ascontiguousarray(gray, dtype=np.float32)).to(dev)
        outs.append(matcher.detect_describe_dev(d_gray, 0.1, pairs, 10))
        exy, esc = D.detect_vectorised(gray, 0.1)
        ek = D.eliminate_redundant(exy, esc, 10)
        assert (outs[-1][0].cpu().numpy() == exy[ek]).all()
        exp_desc.append(pack_descriptors(D.brief_descriptors(gray, exy[ek], pairs.reshape(-1, 2, 2)), 256))
        assert (outs[-1][2].cpu().numpy() == exp_desc[-1]).all()
    d1, d2 = outs[0][2].contiguous(), outs[1][2].contiguous()
    n1, n2 = int(d1.shape[0]), int(d2.shape[0])
    o = torch.empty((3, n1), dtype=torch.int32, device=dev)
    matcher.match_greedy_dev(d1.data_ptr(), n1, d2.data_ptr(), n2, 256, 32, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), n1)
    torch.cuda.synchronize()
    assert (o.T.cpu().numpy() == orc.match_sweep(exp_desc[0], exp_desc[1])).all()
