import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from photogrammetry_b200 import sharding, synthetic
from photogrammetry_b200.keypoint_matching import Matcher
m = Matcher(0)
n = int(sys.argv[1]); G = int(sys.argv[2])
q = synthetic.uniform_descriptors(1234, n, 256); t = synthetic.uniform_descriptors(5678, n, 256)
d_q, d_t = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
shards = [sharding.TrainShardedMatcher(m, d_q, d_t[lo:hi], lo, n, 256, reduce_min=lambda x: None, all_gather=lambda x: None, world_size=G) for lo, hi in sharding.train_slices(n, G)]
r = 0
while True:
    for s in shards: s.step_round()
    red = torch.stack([s.exchange_view() for s in shards]).min(dim=0).values
    for s in shards:
        s.exchange_view().copy_(red); s.step_commit()
    ea = torch.stack([s.edges for s in shards])
    for s in shards: lr, dn = s.step_finish_round(ea)
    print("round", r, "edges per rank", [int(x) for x in ea[:, 0].cpu()], "live rows after", lr, "done", dn)
    r += 1
    if dn: break
