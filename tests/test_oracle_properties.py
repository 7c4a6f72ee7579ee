"""Property tests tying the oracle's formulations together (CPU, small sizes)."""
import numpy as np
import pytest

from oracle import oracle_np as onp
from oracle import orc

CASES = [
    # (n1, n2, desc_bits)
    (1, 1, 256), (2, 1, 256), (1, 2, 256), (7, 7, 1), (13, 5, 3), (5, 13, 8), (33, 33, 8),
    (64, 48, 17), (48, 64, 100), (40, 40, 128), (31, 57, 129), (20, 20, 300), (25, 9, 512), (64, 64, 6),
]


@pytest.mark.parametrize("n1,n2,bits", CASES)
def test_all_formulations_agree(n1, n2, bits):
    q = orc.gen_uniform(1000 + n1, n1, bits)
    t = orc.gen_uniform(2000 + n2, n2, bits)
    lit = orc.match_literal(q, t, kernighan=True)
    assert (lit == orc.match_sweep(q, t)).all()
    assert (lit == orc.match_rounds(q, t)).all()
    assert (lit == onp.match_literal_np(q, t)).all()
    py = onp.match_literal_py([onp.desc_to_int(r) for r in q], [onp.desc_to_int(r) for r in t])
    assert lit.tolist() == [list(x) for x in py]
    # structure: n1 rows, first min(n1,n2) real and sorted by (d, i, j), then the tail
    m = min(n1, n2)
    assert lit.shape == (n1, 3)
    keys = [tuple(r) for r in lit[:m, [2, 0, 1]].tolist()]
    assert keys == sorted(keys)
    assert len(set(lit[:m, 0])) == m and len(set(lit[:m, 1])) == m
    assert (lit[m:] == [0, 0, orc.TAIL_DISTANCE]).all()


def test_duplicates_force_many_rounds():
    # identical descriptors: every distance is 0, greedy pairs (k, k); one pair per round
    q = np.zeros((24, 32), dtype=np.uint8)
    t = np.zeros((24, 32), dtype=np.uint8)
    out, rounds = orc.match_rounds(q, t, return_rounds=True)
    assert out.tolist() == [[k, k, 0] for k in range(24)]
    assert rounds == 24
    assert (orc.match_literal(q, t) == out).all()


def test_empty_inputs():
    q = orc.gen_uniform(1, 4)
    e = np.zeros((0, 32), dtype=np.uint8)
    assert orc.match_literal(e, q).shape == (0, 3)
    assert orc.match_sweep(e, e).shape == (0, 3)
    for f in (orc.match_literal, orc.match_sweep, orc.match_rounds):
        with pytest.raises(orc.EmptyTrainError):
            f(q, e)


def test_knn2_and_ratio_crosscheck_against_matrix():
    q = orc.gen_uniform(5, 70, 8)           # 8-bit descriptors: heavy ties
    t = orc.gen_uniform(6, 50, 8)
    d = orc.distance_matrix(q, t).astype(np.int64)
    key = d * (1 << 20) + np.arange(50)[None, :]
    order = np.argsort(key, axis=1, kind="stable")
    bj, bd, sj, sd = orc.knn2(q, t)
    assert (bj == order[:, 0]).all() and (sj == order[:, 1]).all()
    assert (bd == d[np.arange(70), bj]).all() and (sd == d[np.arange(70), sj]).all()
    ckey = d * (1 << 20) + np.arange(70)[:, None]
    col_best = np.argmin(ckey, axis=0)
    got = orc.match_ratio_crosscheck(q, t, ratio=0.8, cross_check=True)
    exp = [(i, int(bj[i]), int(bd[i])) for i in range(70)
           if np.float32(bd[i]) < np.float32(0.8) * np.float32(sd[i]) and col_best[bj[i]] == i]
    assert [tuple(r) for r in got.tolist()] == exp
    # cross-check alone == round 1 of the greedy rounds (mutual nearest neighbours)
    mutual = orc.match_ratio_crosscheck(q, t, ratio=0.0, cross_check=True)
    greedy = {(a, b) for a, b, _ in orc.match_rounds(q, t)[:50].tolist()}
    assert {(a, b) for a, b, _ in mutual.tolist()} <= greedy


def test_l2_knn2_small():
    rng = np.random.default_rng(3)
    q = rng.standard_normal((20, 16)).astype(np.float32)
    t = rng.standard_normal((30, 16)).astype(np.float32)
    bj, bd, sj, sd = orc.l2_knn2(q, t)
    d = ((q[:, None, :].astype(np.float64) - t[None].astype(np.float64)) ** 2).sum(-1)
    o = np.argsort(d, axis=1, kind="stable")
    assert (bj == o[:, 0]).all() and (sj == o[:, 1]).all()
    np.testing.assert_allclose(bd, d[np.arange(20), bj], rtol=1e-6)


@pytest.mark.parametrize("K", [1, 2, 4])
def test_topk_subround_formulation_equals_reference(K):
    """The design study behind DESIGN.md section 9 (tools/sim_topk_rounds.py): resolving mutual pairs on per-row /
    per-column top-K lists between distance passes yields exactly the reference's greedy assignment."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "sim_topk_rounds", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sim_topk_rounds.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for seed, n1, n2, bits in [(1, 300, 300, 256), (2, 200, 350, 256), (3, 260, 120, 16), (4, 150, 150, 8)]:
        q = orc.gen_uniform(seed, n1, bits)
        t = orc.gen_uniform(seed + 100, n2, bits)
        got, passes = sim.match_topk_rounds(q, t, K)
        exp = orc.match_sweep(q, t)[:min(n1, n2)]
        assert got.shape == exp.shape and (got == exp).all(), (seed, K)
        assert len(passes) >= 1


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_candidate_edges_with_arbitrary_bounds_equal_reference(seed):
    """a distance pass keeps every edge with distance <= T, mutual-best sub-rounds on that list
    accept exactly what the reference's argmin scans (KeypointMatching.cs:38-66) emit next, and ANY bound is correct --
    the planner's Gaussian fit, the quartered bound after an overflow, or the smaller T' a sparse phase truncates its
    list to when the edges do not fit its shared memory (pgm_kernels.cuh, sparse_body).  Restated here in numpy with a
    random bound per pass (including passes that list nothing) against the sweep oracle."""
    rng = np.random.default_rng(seed)
    n1, n2, bits = [(300, 280, 256), (220, 360, 256), (260, 130, 16), (180, 180, 8), (350, 350, 64)][seed - 1]
    q = orc.gen_uniform(seed, n1, bits)
    t = orc.gen_uniform(seed + 50, n2, bits)
    if seed == 5:
        t[40:90] = q[7]                                   # a block of exact duplicates: heavy ties
    qb = np.unpackbits(q, axis=1).astype(np.int64)
    tb = np.unpackbits(t, axis=1).astype(np.int64)
    D = qb.sum(1)[:, None] + tb.sum(1)[None, :] - 2 * qb @ tb.T
    lr, lc = np.ones(n1, bool), np.ones(n2, bool)
    out = []
    while lr.any() and lc.any():
        ri, ci = np.flatnonzero(lr), np.flatnonzero(lc)
        sub = D[np.ix_(ri, ci)]
        # the classic accept of the pass: rows and columns that chose each other
        rb = (sub * (1 << 20) + ci[None, :]).argmin(1)
        cb = (sub * (1 << 20) + ri[:, None]).argmin(0)
        mutual = cb[rb] == np.arange(len(ri))
        ai, aj = ri[mutual], ci[rb[mutual]]
        out += [(int(D[i, j]), int(i), int(j)) for i, j in zip(ai, aj)]
        # candidate edges under a random bound, filtered by the accept
        T = int(rng.integers(-1, int(np.quantile(sub, 0.2)) + 2))
        ei, ej = np.nonzero(sub <= T)
        ei, ej = ri[ei], ci[ej]
        lr[ai] = False; lc[aj] = False
        keep = lr[ei] & lc[ej]
        ei, ej = ei[keep], ej[keep]
        while len(ei):                                    # sparse sub-rounds
            key = (D[ei, ej] << 40) | (ei.astype(np.int64) << 20) | ej
            rbest = np.full(n1, np.iinfo(np.int64).max); cbest = rbest[:n2].copy() if n2 <= n1 else np.full(n2, np.iinfo(np.int64).max)
            np.minimum.at(rbest, ei, key); np.minimum.at(cbest, ej, key)
            ok = (rbest[ei] == key) & (cbest[ej] == key)
            out += [(int(D[i, j]), int(i), int(j)) for i, j in zip(ei[ok], ej[ok])]
            lr[ei[ok]] = False; lc[ej[ok]] = False
            keep = lr[ei] & lc[ej]
            ei, ej = ei[keep], ej[keep]
    exp = orc.match_sweep(q, t)[:min(n1, n2)]
    got = np.array([(i, j, d) for d, i, j in sorted(out)], dtype=np.int32)
    assert got.shape == exp.shape and (got == exp).all()
