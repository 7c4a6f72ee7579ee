#!/usr/bin/env python
"""BASELINE configs[3]: one n x n pair with the train set sharded over the ranks (pgm_multi_*: NCCL inside the library).

    python tools/bench_sharded.py [--size 200000] [--dist U] [--check]                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/bench_sharded.py --check                                                              # G GPUs

--check compares every rank's triples with the unsharded single-GPU call on rank 0 (bit-identity) and also runs the
nearest / second-nearest search with the top-2 merge against the unsharded knn2."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from photogrammetry_b200 import sharding, synthetic
from photogrammetry_b200.keypoint_matching import Matcher


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=200_000)
    ap.add_argument("--dist", default="U", choices=["U", "C"])
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = Matcher(local)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); m.set_stream(stream.cuda_stream)
    popc, _ = m.measure_popc_peak(200)
    line = sharding.bench_train_sharded(m, stream, dev, rank, world, popc, n=args.n, dist_name=args.dist)
    if args.check:
        n = args.n
        q = synthetic.uniform_descriptors(1234, n, 256)
        t = synthetic.uniform_descriptors(5678, n, 256) if args.dist == "U" else synthetic.noisy_copy_descriptors(42, q, 256)
        lo, hi = sharding.train_slices(n, world)[rank]
        mg = sharding.MultiGpuMatcher(m, rank, world)
        d_q, d_t = torch.from_numpy(q).to(dev), torch.from_numpy(np.ascontiguousarray(t[lo:hi])).to(dev)
        got = mg.match_train_sharded(d_q, d_t, lo, n).clone()
        knn = [x.clone() for x in mg.knn2_train_sharded(d_q, d_t, lo)]
        xb, xn = mg.exchange()
        d_tf = torch.from_numpy(t).to(dev)
        ref = torch.empty((3, n), dtype=torch.int32, device=dev)
        m.match_greedy_dev(d_q.data_ptr(), n, d_tf.data_ptr(), n, 256, 32, ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr(), n)
        ref_knn = m.knn2_hamming_dev(d_q, d_tf, 256)
        torch.cuda.synchronize()
        same = torch.tensor([int(torch.equal(got, ref)), int(all(torch.equal(a, b) for a, b in zip(knn, ref_knn)))], device=dev)
        if world > 1:
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        line["every_rank_bit_identical_to_unsharded"] = bool(same[0].item())
        line["knn2_top2_merge_bit_identical_to_unsharded"] = bool(same[1].item())
        line["knn2_all_gather_bytes_per_rank"] = xb
        mg.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


main()
