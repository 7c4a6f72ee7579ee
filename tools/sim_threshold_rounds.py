#!/usr/bin/env python
"""Design study (CPU only, numpy) for the "candidate edge" form of the greedy matcher.

Every distance pass over live rows x live columns keeps, besides each row's / column's minimum (the K = 1
mutual-nearest-neighbour round the engine has always run), EVERY edge whose distance is <= T.  Because the edge set
is complete up to T on both sides, the greedy matching restricted to those edges is a prefix of the reference's
greedy matching: sub-rounds of mutual-best on the sparse edge list accept exactly what the reference's argmin scans
(KeypointMatching.cs:38-66) would emit next, and need no distance evaluation.  T comes from a Gaussian fit of a
sample of distances so that a row sees about `c` candidates.

    python tools/sim_threshold_rounds.py 8192 U 6
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.sim_topk_rounds import distance_matrix  # noqa: E402


def inv_norm_tail(p: float) -> float:
    """z with P(N(0,1) < -z) = p (Abramowitz-Stegun 26.2.23)."""
    p = min(max(p, 1e-12), 0.5)
    t = math.sqrt(-2.0 * math.log(p))
    return t - (2.515517 + 0.802853 * t + 0.010328 * t * t) / (1 + 1.432788 * t + 0.189269 * t * t + 0.001308 * t ** 3)


def run(q, t, c, edge_budget=0):
    D = distance_matrix(q, t)
    n1, n2 = D.shape
    samp = D[:: max(1, n1 // 64), :: max(1, n2 // 64)][:64, :64].astype(np.float64)
    mu, sd = samp.mean(), samp.std()
    lr, lc = np.ones(n1, bool), np.ones(n2, bool)
    matched = []
    passes = []
    while lr.any() and lc.any():
        ri, ci = np.flatnonzero(lr), np.flatnonzero(lc)
        sub = D[np.ix_(ri, ci)]
        cc = max(c, edge_budget / max(len(ri), len(ci))) if edge_budget else c
        z = inv_norm_tail(min(cc / min(len(ri), len(ci)), 0.25))
        T = int(math.floor(mu - z * sd))
        # K = 1 round
        keyr = sub * (1 << 20) + ci[None, :]
        keyc = sub * (1 << 20) + ri[:, None]
        rb = keyr.argmin(1); cb = keyc.argmin(0)
        mutual = cb[rb] == np.arange(len(ri))
        acc_i, acc_j = ri[mutual], ci[rb[mutual]]
        n_k1 = len(acc_i)
        # candidate edges
        ei, ej = np.nonzero(sub <= T)
        ed = sub[ei, ej]; ei = ri[ei]; ej = ci[ej]
        n_edges = len(ei)
        lr[acc_i] = False; lc[acc_j] = False
        matched += list(zip(acc_i, acc_j))
        live = lr[ei] & lc[ej]
        ei, ej, ed = ei[live], ej[live], ed[live]
        n_live_edges = len(ei)
        sub_rounds = 0; n_sparse = 0
        per_sub = []
        while len(ei):
            key = (ed.astype(np.int64) << 40) | (ei.astype(np.int64) << 20) | ej
            rbest = np.full(n1, np.iinfo(np.int64).max); cbest = np.full(n2, np.iinfo(np.int64).max)
            np.minimum.at(rbest, ei, key); np.minimum.at(cbest, ej, key)
            ok = (rbest[ei] == key) & (cbest[ej] == key)
            ai, aj = ei[ok], ej[ok]
            lr[ai] = False; lc[aj] = False
            matched += list(zip(ai, aj))
            n_sparse += len(ai)
            live = lr[ei] & lc[ej]
            per_sub.append((len(ei), len(ai)))
            ei, ej, ed = ei[live], ej[live], ed[live]
            sub_rounds += 1
        passes.append(dict(nlr=len(ri), nlc=len(ci), T=T, k1=n_k1, edges=n_edges, live_edges=n_live_edges,
                           sub_rounds=sub_rounds, sparse_acc=n_sparse, per_sub=per_sub[:12]))
    return passes, matched, D


if __name__ == "__main__":
    from photogrammetry_b200 import synthetic
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dist = sys.argv[2] if len(sys.argv) > 2 else "U"
    c = float(sys.argv[3]) if len(sys.argv) > 3 else 6.0
    budget = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    q, t = synthetic.config2_pair(n, dist)
    passes, matched, D = run(q, t, c, budget)
    for p in passes:
        print(p)
    print("recompute factor", sum(p["nlr"] * p["nlc"] for p in passes) / (n * n))
    if n <= 4096:
        from oracle import orc
        ref = orc.match_sweep(q, t)
        got = sorted((int(D[i, j]), int(i), int(j)) for i, j in matched)
        exp = [(int(r[2]), int(r[0]), int(r[1])) for r in ref[: min(len(q), len(t))]]
        print("equals oracle sweep:", got == exp)
