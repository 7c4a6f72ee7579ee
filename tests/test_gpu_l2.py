"""Float-descriptor extension: squared-L2 top-2 on tcgen05 vs the exact oracle.

Tolerances (north_star): indices identical except documented near-ties; distances within 1e-4 relative."""
import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _check(matcher, q, t):
    bj, bd, sj, sd = matcher.knn2_l2(q, t)
    ej, ed, fj, fd = orc.l2_knn2(q, t)
    n2 = len(t)
    d = ((q[:, None, :].astype(np.float64) - t[None].astype(np.float64)) ** 2).sum(-1) if len(q) * n2 <= 4_000_000 else None
    for got_j, got_d, exp_j, exp_d in ((bj, bd, ej, ed), (sj, sd, fj, fd)):
        assert ((got_j >= 0) == (exp_j >= 0)).all()
        ok = exp_j >= 0
        np.testing.assert_allclose(got_d[ok], exp_d[ok], rtol=RTOL, atol=1e-6)       # distances: 1e-4 relative
        diff = ok & (got_j != exp_j)
        # an index may differ only at a near-tie: the chosen neighbour is as close as the oracle's within RTOL
        if diff.any():
            assert d is not None
            rows = np.nonzero(diff)[0]
            assert np.allclose(d[rows, got_j[rows]], exp_d[rows], rtol=RTOL, atol=1e-6)
            assert diff.mean() < 0.01
    return bj, bd, sj, sd


@pytest.mark.parametrize("n1,n2,dim", [(128, 128, 64), (128, 256, 128), (300, 500, 128), (1000, 777, 100),
                                        (77, 1300, 32), (5, 3, 128), (9, 1, 16), (129, 4097, 128)])
def test_l2_knn2_random(matcher, n1, n2, dim):
    rng = np.random.default_rng(n1 * 7 + n2)
    q = rng.standard_normal((n1, dim)).astype(np.float32)
    t = rng.standard_normal((n2, dim)).astype(np.float32)
    _check(matcher, q, t)


def test_l2_noisy_copies_small_distances(matcher):
    # train = query + small noise: the cancellation-prone case for ||q||^2 + ||t||^2 - 2 q.t
    rng = np.random.default_rng(5)
    q = (rng.standard_normal((2000, 128)) * 3).astype(np.float32)
    t = (q[rng.permutation(2000)] + rng.standard_normal((2000, 128)).astype(np.float32) * 0.01).astype(np.float32)
    bj, bd, _, _ = _check(matcher, q, t)
    assert bd.max() < 0.1


def test_l2_config2_shape_8k(matcher):
    rng = np.random.default_rng(11)
    q = rng.random((8192, 128), dtype=np.float32)           # SIFT-like non-negative descriptors
    t = rng.random((8192, 128), dtype=np.float32)
    _check(matcher, q, t)
