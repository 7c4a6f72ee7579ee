#!/usr/bin/env python
"""2-rank diagnostic of the train-sharded exchange: verifies every all-reduce against an all-gather + local min.
torchrun --nproc-per-node 2 tools/debug_shard_nccl.py N [--default-stream]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from photogrammetry_b200 import sharding, synthetic
from photogrammetry_b200.keypoint_matching import Matcher

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
use_default = "--default-stream" in sys.argv
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
q = synthetic.uniform_descriptors(1234, n, 256); t = synthetic.uniform_descriptors(5678, n, 256)
lo, hi = sharding.train_slices(n, world)[rank]
m = Matcher(local)
if not use_default:
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); m.set_stream(stream.cuda_stream)
else:
    m.set_stream(torch.cuda.current_stream().cuda_stream)
d_q = torch.from_numpy(q).to(dev); d_t = torch.from_numpy(t[lo:hi].copy()).to(dev)
sm = sharding.TrainShardedMatcher(m, d_q, d_t, lo, n, 256)

def checked_reduce(x, name, rnd):
    torch.cuda.synchronize()
    pre = x.clone(); gathered = [torch.empty_like(pre) for _ in range(world)]
    dist.all_gather(gathered, pre); torch.cuda.synchronize()
    exp = torch.stack(gathered).min(dim=0).values
    dist.all_reduce(x, op=dist.ReduceOp.MIN); torch.cuda.synchronize()
    bad = int((x != exp).sum())
    none = int((exp == 0x7F7F7F7F).sum())
    if rank == 0 or bad:
        print(f"[r{rank}] round {rnd} {name}: mismatches={bad} none={none} local_nonnone={int((pre != 0x7F7F7F7F).sum())}", flush=True)

prev = n + 1
for rnd in range(40):
    sm.step_round(); checked_reduce(sm.xkeys, "xkeys", rnd)
    sm.step_propose(); checked_reduce(sm.xacc, "xacc", rnd)
    lr, lc = sm.step_commit()
    print(f"[r{rank}] round {rnd}: live_rows={lr} live_cols_local={lc}", flush=True)
    if sm.done(lr): break
    if lr >= prev: print(f"[r{rank}] NO PROGRESS"); break
    prev = lr
dist.barrier(); dist.destroy_process_group()
