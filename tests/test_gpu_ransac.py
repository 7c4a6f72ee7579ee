"""RANSAC hypothesis scoring (pgm_ransac_score, CameraPoseEstimation.cs:41-88) against the numpy oracle.
Same float32 operation order on both sides, so counts, winner and inlier mask must agree exactly."""
import numpy as np
import pytest

from oracle import ransac_np as R
from photogrammetry_b200.camera_pose_estimation import CameraPoseEstimation, InvalidOperationException
from photogrammetry_b200.keypoint import Coordinate, Keypoint, KeypointPair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_hyp,n,seed", [(1, 1, 0), (7, 33, 1), (300, 1285, 2), (2000, 2175, 3), (64, 0, 4)])
def test_score_against_oracle(matcher, n_hyp, n, seed):
    rng = np.random.default_rng(seed)
    F = (rng.standard_normal((n_hyp, 3, 3)) * np.array([1e-6, 1e-6, 1e-3, 1e-6, 1e-6, 1e-3, 1e-3, 1e-3, 1.0]).reshape(3, 3)).astype(np.float32)
    xy1 = rng.integers(0, 4000, size=(n, 2)).astype(np.int32)
    xy2 = rng.integers(0, 4000, size=(n, 2)).astype(np.int32)
    valid = (rng.random(n_hyp) < 0.7).astype(np.uint8)
    for v in (None, valid):
        for thr in (0.001, 0.0, -0.5):
            ec, eb, em = R.score(F, v, xy1, xy2, thr)
            gc, gb, gm = matcher.ransac_score(F, xy1, xy2, thr, v)
            assert (gc == ec).all() and gb == eb and (gm == em).all()


def test_ties_pick_the_first_hypothesis(matcher):
    F = np.zeros((5, 3, 3), dtype=np.float32)
    F[:, 2, 2] = [1.0, -1.0, -1.0, 1.0, -1.0]            # residual == F[2][2]: hypotheses 1, 2, 4 accept every pair
    xy = np.zeros((10, 2), dtype=np.int32)
    counts, best, mask = matcher.ransac_score(F, xy, xy, 0.0)
    assert counts.tolist() == [0, 10, 10, 0, 10] and best == 1 and mask.all()
    counts, best, mask = matcher.ransac_score(F, xy, xy, -2.0)
    assert counts.tolist() == [0] * 5 and best == -1 and not mask.any()


def _pairs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 3000, size=(n, 2))
    b = a + rng.integers(-40, 41, size=(n, 2))
    return [KeypointPair(Keypoint(Coordinate(int(x1), int(y1)), 0), Keypoint(Coordinate(int(x2), int(y2)), 0), 0)
            for (x1, y1), (x2, y2) in zip(a, b)]


def test_get_fundamental_matrix_mirror(matcher):
    pairs = _pairs(500, 9)
    est = CameraPoseEstimation(matcher=matcher, seed=3, require_rank2=False)
    sample, F = est.GetFundamentalMatrix(pairs, 200, 32, 0.001)
    Fs, valid, counts = est.last_hypotheses
    xy1 = np.array([[p.Keypoint1.Coordinate.X, p.Keypoint1.Coordinate.Y] for p in pairs], dtype=np.int32)
    xy2 = np.array([[p.Keypoint2.Coordinate.X, p.Keypoint2.Coordinate.Y] for p in pairs], dtype=np.int32)
    ec, eb, em = R.score(Fs, valid, xy1, xy2, 0.001)
    assert (counts == ec).all() and (F == Fs[eb]).all()
    assert [id(p) for p in sample] == [id(p) for p, k in zip(pairs, em.tolist()) if k]      # the caller's own objects
    # every estimate annihilates (approximately) its own sample: the 8-point construction is consistent
    assert np.isfinite(Fs).all() and Fs.shape == (200, 3, 3)
    with pytest.raises(InvalidOperationException):
        est.GetFundamentalMatrix(pairs, 10, 7, 0.001)
    with pytest.raises(InvalidOperationException):
        est.GetFundamentalMatrix(pairs[:20], 10, 32, 0.001)
    # upstream's rank gate (:46-51): generic noisy samples are full rank, so nothing survives and upstream throws
    strict = CameraPoseEstimation(matcher=matcher, seed=3, require_rank2=True)
    try:
        strict.GetFundamentalMatrix(pairs, 50, 32, 0.001)
    except Exception as e:
        assert "Failed computing the best fundamental matrix" in str(e)
