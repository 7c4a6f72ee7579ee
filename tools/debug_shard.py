#!/usr/bin/env python
"""Single-GPU scale check of the train-sharded mode: emulated ranks vs the unsharded matcher at sizes the
pytest suite does not reach.  python tools/debug_shard.py 50000 200000"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from photogrammetry_b200 import sharding, synthetic
from photogrammetry_b200.keypoint_matching import Matcher

m = Matcher(0)
for n in [int(a) for a in sys.argv[1:]] or [50000]:
    q = synthetic.uniform_descriptors(1234, n, 256)
    t = synthetic.uniform_descriptors(5678, n, 256)
    t0 = time.perf_counter()
    base = m.match_greedy(q, t, 256)
    dt = time.perf_counter() - t0
    ok = sorted(base[:, 0].tolist()) == list(range(n)) and sorted(base[:, 1].tolist()) == list(range(n))
    ok &= bool((np.bitwise_count(q[base[:, 0]] ^ t[base[:, 1]]).sum(axis=1) == base[:, 2]).all())
    print(f"n={n} unsharded {dt*1e3:.1f} ms stats={m.stats()} properties_ok={ok}", flush=True)
    for shards in (2, 8):
        try:
            t0 = time.perf_counter()
            outs, rounds = sharding.match_train_sharded_emulated(m, q, t, shards)
            dt = time.perf_counter() - t0
            print(f"  shards={shards} rounds={rounds} {dt*1e3:.1f} ms identical={[bool((o == base).all()) for o in outs]}", flush=True)
        except Exception as e:
            print(f"  shards={shards} FAILED: {e!r}", flush=True)
