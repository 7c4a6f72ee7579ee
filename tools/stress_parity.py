"""Randomised parity stress (developer tool; the fixed cases live in tests/): random shapes, descriptor widths, duplicate
blocks and near-copies through the single-pair call (latency mode) and the batch engine (throughput mode), with the sparse
phase's capacity knob varied, every result compared bit for bit with the oracle's sweep restatement of
KeypointMatching.cs:14-69.

    python tools/stress_parity.py [cases=120] [seed=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import orc
from photogrammetry_b200 import synthetic
from photogrammetry_b200.keypoint_matching import Matcher

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
m = Matcher(0)
bad = 0
for k in range(cases):
    bits = int(rng.choice([4, 8, 32, 64, 128, 256, 256, 256, 384, 512]))
    big = rng.random() < 0.25
    n1 = int(rng.integers(1, 6000 if big else 1500)); n2 = int(rng.integers(1, 6000 if big else 1500))
    q = synthetic.uniform_descriptors(int(rng.integers(1 << 30)), n1, bits)
    kind = rng.integers(4)
    if kind == 0:
        t = synthetic.uniform_descriptors(int(rng.integers(1 << 30)), n2, bits)
    elif kind == 1:                                   # noisy copies of the queries (the matcher's typical input)
        src = q[rng.integers(0, n1, size=n2)]
        flip = rng.random(src.shape) < 0.02
        t = src ^ (flip * (1 << rng.integers(0, 8, size=src.shape))).astype(np.uint8)
        if bits % 8: t[:, bits // 8] &= (1 << (bits % 8)) - 1
        t[:, (bits + 7) // 8:] = 0
    elif kind == 2:                                   # blocks of exact duplicates: heavy ties
        t = synthetic.uniform_descriptors(int(rng.integers(1 << 30)), n2, bits)
        a = int(rng.integers(0, n2)); b = min(n2, a + int(rng.integers(1, 200)))
        t[a:b] = q[int(rng.integers(0, n1))]
        c = int(rng.integers(0, n1)); d = min(n1, c + int(rng.integers(1, 200)))
        q[c:d] = q[c]
    else:                                             # everything close to one descriptor
        base = synthetic.uniform_descriptors(7, 1, bits)
        t = np.repeat(base, n2, axis=0); t[:, 0] ^= rng.integers(0, 4, size=n2).astype(np.uint8)
        if bits < 8: t[:, 0] &= (1 << bits) - 1
    slots = str(int(rng.choice([1, 2, 3, 7, 32])))
    os.environ["PGM_SP_SLOTS_MAX"] = slots
    exp = orc.match_sweep(q, t) if n2 > 0 else None
    got = m.match_greedy(q, t, bits)
    ok = got.shape == exp.shape and bool((got == exp).all())
    if ok and rng.random() < 0.5:                     # the same pair through the batch engine
        imgs = np.concatenate([q, t]); offs = np.array([0, n1, n1 + n2], dtype=np.int64)
        pairs = np.array([(0, 1)] * 17, dtype=np.int32)
        tr, starts, counts = m.match_pairs_batch(imgs, offs, pairs, bits)
        ok = all(bool((tr[starts[p]:starts[p] + counts[p]] == exp).all()) for p in (0, 9, 16))
    if not ok:
        bad += 1
        print("MISMATCH", dict(case=k, n1=n1, n2=n2, bits=bits, kind=int(kind), slots=slots), flush=True)
print(f"{cases - bad} of {cases} cases bit-exact")
sys.exit(1 if bad else 0)
