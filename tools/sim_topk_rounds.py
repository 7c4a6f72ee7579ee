#!/usr/bin/env python
"""Design study for DESIGN.md section 9 (CPU only, numpy): greedy matching by mutual-nearest-neighbour rounds where
every distance pass keeps the K best columns of each live row and the K best rows of each live column, and mutual
pairs are then resolved on those lists ("sub-rounds") until no list yields a new pair -- only then are distances
recomputed for the survivors.  A row's first live list entry is its true best live column as long as the list is
not exhausted (everything outside the list is worse than every entry), so each accepted pair is locally dominant and
the result equals the reference's greedy assignment; K = 1 is the algorithm the CUDA engine runs today.

    python tools/sim_topk_rounds.py 8192 U        # passes, survivors per pass and recompute factor for K = 1,2,3,4,8
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def distance_matrix(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    qb = np.unpackbits(q, axis=1).astype(np.float32)
    tb = np.unpackbits(t, axis=1).astype(np.float32)
    return (qb.sum(1)[:, None] + tb.sum(1)[None, :] - 2 * qb @ tb.T).astype(np.int64)


def match_topk_rounds(q: np.ndarray, t: np.ndarray, K: int):
    """Returns (triples int32[min(n1,n2), 3] in (d, i, j) order, [(live rows, live cols, sub-rounds) per pass])."""
    D = distance_matrix(q, t)
    n1, n2 = D.shape
    keyr = D * (1 << 20) + np.arange(n2)[None, :]          # a row's key over columns: (d, j)
    keyc = D * (1 << 20) + np.arange(n1)[:, None]          # a column's key over rows: (d, i)
    lr, lc = np.ones(n1, bool), np.ones(n2, bool)
    out, passes = [], []
    while lr.any() and lc.any():
        ri, ci = np.flatnonzero(lr), np.flatnonzero(lc)
        kr, kc = min(K, len(ci)), min(K, len(ri))
        sub = keyr[np.ix_(ri, ci)]
        rl = np.sort(np.partition(sub, kr - 1, axis=1)[:, :kr], axis=1) & ((1 << 20) - 1)          # [rows, kr] column ids
        subc = keyc[np.ix_(ri, ci)]
        cl = (np.sort(np.partition(subc, kc - 1, axis=0)[:kc, :], axis=0) & ((1 << 20) - 1)).T     # [cols, kc] row ids
        n_sub = 0
        while True:
            ra = lc[rl]
            rfirst = np.where(ra.any(1), rl[np.arange(len(ri)), ra.argmax(1)], -1)
            ca = lr[cl]
            cfirst = np.where(ca.any(1), cl[np.arange(len(ci)), ca.argmax(1)], -1)
            rowprop = np.full(n1, -1); rowprop[ri] = np.where(lr[ri], rfirst, -1)
            colprop = np.full(n2, -2); colprop[ci] = np.where(lc[ci], cfirst, -2)
            i = np.flatnonzero(rowprop >= 0)
            j = rowprop[i]
            ok = colprop[j] == i
            if not ok.any():
                break
            out.append(np.stack([i[ok], j[ok], D[i[ok], j[ok]]], axis=1))
            lr[i[ok]] = False; lc[j[ok]] = False
            n_sub += 1
        passes.append((len(ri), len(ci), n_sub))
    tr = np.concatenate(out) if out else np.zeros((0, 3), np.int64)
    order = np.lexsort((tr[:, 1], tr[:, 0], tr[:, 2]))
    return tr[order].astype(np.int32), passes


if __name__ == "__main__":
    from photogrammetry_b200 import synthetic
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dist = sys.argv[2] if len(sys.argv) > 2 else "U"
    q, t = synthetic.config2_pair(n, dist)
    for K in (1, 2, 3, 4, 8):
        tr, passes = match_topk_rounds(q, t, K)
        print(f"{dist} {n} K={K}: {len(passes)} distance passes, live rows per pass {[p[0] for p in passes][:10]}, "
              f"sub-rounds {[p[2] for p in passes][:10]}, recompute factor {sum(a * b for a, b, _ in passes) / (n * n):.3f}")
