"""The small cases run under compute-sanitizer (memcheck / racecheck / synccheck / initcheck, one tool per gpurun call):

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py

  1. 1024 x 1024 single pair: latency mode -- init, round 0, the persistent tail kernel with its hand-rolled grid barrier,
     candidate-edge filter, sparse sub-rounds, finisher_prepare -> finisher hand-over through global memory, sliced ordering
  2. 300 x 280 all-duplicate descriptors: every distance is 0 (candidate lists overflow, one accept per round, n1 > n2 tail)
  3. 24 pairs of 512 x 512: throughput mode (standalone round / accept+filter / sparse / finisher / order kernels)
  4. float L2 512 x 512, D = 128: the tcgen05 CTA-pair kernel with its cluster mbarriers
  5. 3 emulated train shards of a 700 x 650 pair (export / propose / stable commit kernels)
Every result is checked against the oracle, so a sanitizer run is also a parity run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import orc
from photogrammetry_b200 import sharding, synthetic
from photogrammetry_b200.keypoint_matching import Matcher

m = Matcher(0)
q, t = synthetic.config2_pair(1024, "U")
assert (m.match_greedy(q, t, 256) == orc.match_sweep(q, t)).all()
print("case 1 ok", m.stats())
d = np.tile(synthetic.uniform_descriptors(3, 1, 256), (300, 1))
assert (m.match_greedy(d, d[:280], 256) == orc.match_sweep(d, d[:280])).all()
print("case 2 ok")
per, n_img = 512, 9
imgs = np.concatenate([synthetic.uniform_descriptors(100 + k, per, 256) for k in range(n_img)])
offs = np.arange(n_img + 1, dtype=np.int64) * per
pairs = sharding.all_pairs(n_img)[:24]
tr, starts, counts = m.match_pairs_batch(imgs, offs, pairs, 256)
for p, (a, b) in enumerate(pairs):
    assert (tr[starts[p]:starts[p] + counts[p]] == orc.match_sweep(imgs[offs[a]:offs[a + 1]], imgs[offs[b]:offs[b + 1]])).all()
print("case 3 ok", m.stats())
rng = np.random.default_rng(1)
fq, ft = rng.random((512, 128), dtype=np.float32), rng.random((512, 128), dtype=np.float32)
bj, bd, sj, sd = m.knn2_l2(fq, ft)
ej, ed, _, _ = orc.l2_knn2(fq, ft)
assert (np.abs(bd - ed) <= 1e-4 * np.maximum(ed, 1e-12)).all() and (bj != ej).mean() <= 0.01
print("case 4 ok")
q, t = synthetic.uniform_descriptors(5, 700, 256), synthetic.uniform_descriptors(6, 650, 256)
outs, rounds = sharding.match_train_sharded_emulated(m, q, t, 3)
exp = orc.match_sweep(q, t)
assert all((o == exp).all() for o in outs)
print("case 5 ok", rounds)
m.close()
print("ALL OK")
