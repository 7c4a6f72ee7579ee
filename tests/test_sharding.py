"""Multi-process host logic of the pair-sharded mode on CPU (gloo, world_size 2).

The CUDA matcher cannot run here, so each rank's `Matcher` is replaced by a stand-in that
answers from the oracle -- allowed in tests/ only; the product path never does this."""
import os
import socket

import numpy as np
import pytest

from oracle import orc
from photogrammetry_b200 import sharding, synthetic


def test_partition_is_contiguous_and_balanced():
    sizes = [100, 400, 50, 300, 300, 10, 250]
    pairs = sharding.all_pairs(len(sizes))
    assert len(pairs) == 21 and (pairs[:, 0] < pairs[:, 1]).all()
    for world in (1, 2, 3, 4, 8, 30):
        parts = sharding.partition_pairs(pairs, sizes, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(pairs)
        assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
        cost = sharding.pair_costs(pairs, sizes)
        loads = [cost[a:b].sum() for a, b in parts]
        if world <= 4:
            assert max(loads) <= cost.sum() / world + cost.max()
    assert sharding.partition_pairs(pairs[:0], sizes, 3) == [(0, 0)] * 3
    assert (sharding.consecutive_pairs(4) == [[0, 1], [1, 2], [2, 3]]).all()
    # uniform sizes -> equal counts (the 512-image config: 16352 pairs per GPU at 8 GPUs)
    parts = sharding.partition_pairs(sharding.all_pairs(512), [4096] * 512, 8)
    assert [b - a for a, b in parts] == [16352] * 8


def test_merge_top2_matches_unsharded_knn():
    q = orc.gen_uniform(1, 60, 8)
    t = orc.gen_uniform(2, 90, 8)          # 8-bit descriptors: many ties
    bj, bd, sj, sd = orc.knn2(q, t)
    none = np.uint32(0xFFFFFFFF)
    kb, ks = [], []
    for lo, hi in [(0, 31), (31, 32), (32, 90)]:
        a, b, c, d = orc.knn2(q, t[lo:hi])
        kb.append(np.where(a >= 0, (b.astype(np.uint32) << 20) | (a + lo).astype(np.uint32), none))
        ks.append(np.where(c >= 0, (d.astype(np.uint32) << 20) | (c + lo).astype(np.uint32), none))
    best, second = sharding.merge_top2(np.stack(kb), np.stack(ks))
    assert ((best & 0xFFFFF) == bj).all() and ((best >> 20) == bd).all()
    assert ((second & 0xFFFFF) == sj).all() and ((second >> 20) == sd).all()


class _OracleMatcher:
    """Stand-in for Matcher.match_pairs_batch in the CPU test (oracle-backed)."""

    def match_pairs_batch(self, all_desc, image_offsets, pair_list, desc_bits=256, reference_compat_tail=True):
        offs = np.asarray(image_offsets, dtype=np.int64)
        out, starts, counts, base = [], [], [], 0
        for a, b in np.asarray(pair_list).reshape(-1, 2):
            tr = orc.match_sweep(all_desc[offs[a]:offs[a + 1]], all_desc[offs[b]:offs[b + 1]])
            out.append(tr); starts.append(base); counts.append(len(tr)); base += len(tr)
        return (np.concatenate(out) if out else np.zeros((0, 3), np.int32), np.asarray(starts, np.int64),
                np.asarray(counts, np.int32))


def _worker(rank, world, port, sizes, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        imgs = [synthetic.uniform_descriptors(50 + k, n, 64) for k, n in enumerate(sizes)]
        all_desc = np.concatenate(imgs)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        pairs = sharding.all_pairs(len(sizes))
        local = sharding.match_pairs_sharded(_OracleMatcher(), all_desc, offs, pairs, rank, world, 64)
        res = sharding.gather_match_lists(local, rank, world, dst=0)
        if rank == 0:
            q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_pair_sharded_world2_gloo_equals_single_process():
    import torch.multiprocessing as mp
    sizes = [40, 25, 60, 10, 33]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, queue)) for r in range(2)]
    for p in procs:
        p.start()
    triples, starts, counts = queue.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    imgs = [synthetic.uniform_descriptors(50 + k, n, 64) for k, n in enumerate(sizes)]
    pairs = sharding.all_pairs(len(sizes))
    assert len(counts) == len(pairs) == 10
    for p, (a, b) in enumerate(pairs):
        exp = orc.match_sweep(imgs[a], imgs[b])
        assert counts[p] == len(exp) and (triples[starts[p]:starts[p] + counts[p]] == exp).all()


# ---- train-sharded nearest neighbours: the top-2 merge over an all-gather (gloo, world_size 2) --------
class _OracleKnnMatcher:
    """CPU stand-in for the four device-side Matcher methods TrainShardedKnn uses (oracle-backed, tests only)."""

    def knn2_hamming_dev(self, d_q, d_t, desc_bits=256):
        import torch
        n1 = d_q.shape[0]
        if n1 == 0 or d_t.shape[0] == 0:
            return tuple(torch.full((n1,), -1, dtype=torch.int32) for _ in range(4))
        return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in orc.knn2(d_q.numpy(), d_t.numpy()))

    def pack_top2_keys_dev(self, bj, bd, sj, sd, off):
        import torch
        none = torch.tensor(0x7F7F7F7F, dtype=torch.int32)
        return torch.stack([torch.where(bj >= 0, (bd << 20) | (bj + off), none),
                            torch.where(sj >= 0, (sd << 20) | (sj + off), none)])

    def merge_top2_dev(self, keys):
        import torch
        flat = keys.reshape(-1, keys.shape[-1]).numpy().astype(np.uint32)
        flat = np.where(flat == 0x7F7F7F7F, np.uint32(0xFFFFFFFF), flat)
        g = flat.shape[0] // 2
        best, second = sharding.merge_top2(flat.reshape(g, 2, -1)[:, 0], flat.reshape(g, 2, -1)[:, 1])
        def unpack(k):
            ok = k != 0xFFFFFFFF
            return (torch.from_numpy(np.where(ok, k & 0xFFFFF, -1).astype(np.int32)),
                    torch.from_numpy(np.where(ok, k >> 20, -1).astype(np.int32)))
        (bj, bd), (sj, sd) = unpack(best), unpack(second)
        return bj, bd, sj, sd

    def ratio_crosscheck_filter_dev(self, n2, bj, bd, sd, col, ratio=0.8, cross_check=True, max_dist=-1):
        import torch
        bj, bd, sd = bj.numpy(), bd.numpy(), sd.numpy()
        keep = bj >= 0
        if ratio > 0 and n2 >= 2:
            keep &= bd.astype(np.float32) < np.float32(ratio) * sd.astype(np.float32)
        if cross_check:
            keep &= col.numpy()[np.maximum(bj, 0)] == np.arange(len(bj))
        if max_dist >= 0:
            keep &= bd <= max_dist
        i = np.nonzero(keep)[0].astype(np.int32)
        return torch.from_numpy(np.stack([i, bj[i], bd[i]]).astype(np.int32))


def _knn_worker(rank, world, port, q_out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q = orc.gen_uniform(3, 150, 16)
        t = orc.gen_uniform(4, 201, 16)                       # 16-bit descriptors: many ties; odd size: uneven slices
        lo, hi = sharding.train_slices(len(t), world)[rank]
        sh = sharding.TrainShardedKnn(_OracleKnnMatcher(), torch.from_numpy(q), torch.from_numpy(t[lo:hi]), lo, len(t), 16)
        knn = [a.numpy() for a in sh.knn2()]
        kept = sh.match_ratio_crosscheck(0.9, True, -1).numpy()
        q_out.put((rank, knn, kept))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_train_sharded_knn_world2_gloo_equals_unsharded():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_knn_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got = [queue.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    q = orc.gen_uniform(3, 150, 16)
    t = orc.gen_uniform(4, 201, 16)
    exp_knn = orc.knn2(q, t)
    exp_kept = orc.match_ratio_crosscheck(q, t, 0.9, True, -1)
    for _, knn, kept in got:                                   # every rank holds the full, identical answer
        for a, b in zip(knn, exp_knn):
            assert (a == b).all()
        assert (kept.T == exp_kept).all()


@pytest.mark.parametrize("n1,n2,shards,bits", [(150, 201, 3, 16), (90, 64, 5, 8), (40, 4, 7, 256), (33, 1, 2, 256), (60, 77, 1, 12)])
def test_train_sharded_knn_emulated_ranks_cpu(n1, n2, shards, bits):
    """The host logic of TrainShardedKnn (slicing, key packing, padded column gather, merge) with every rank
    emulated in one process; slices smaller than the shard count and empty slices included."""
    import torch
    q = orc.gen_uniform(11, n1, bits)
    t = orc.gen_uniform(12, n2, bits)
    for ratio, cc, md in [(0.8, True, -1), (0.0, True, 3), (0.95, False, -1)]:
        knn, kept = sharding.knn_train_sharded_emulated(_OracleKnnMatcher(), torch.from_numpy(q), torch.from_numpy(t), shards,
                                                        bits, ratio, cc, md)
        for a, b in zip(knn, orc.knn2(q, t)):
            assert (a.numpy() == b).all()
        exp = orc.match_ratio_crosscheck(q, t, ratio, cc, md)
        assert kept.numpy().T.shape == exp.shape and (kept.numpy().T == exp).all()


# ---- train-sharded greedy matcher: the [R | P] exchange rule of one round (gloo, world_size 2) -------------------
# include/pgmatch.h: R[i] = row i's best key (d << 20 | global j) over the rank's columns, P[i] = the best key among the
# rank's columns that chose row i as their best row; after an element-wise MIN over the ranks row i is matched iff
# R[i] == P[i].  The device kernels (shard_export_rows / shard_propose_cols / shard_commit_mark) compute exactly this; the
# test restates the rank-local part in numpy, runs the exchange through gloo and checks the rule against the unsharded
# definition of a locally dominant edge (mutual best under the reference's (d, i, j) order).
NONE_KEY = 0x7F7F7F7F


def _local_rp(q, t_local, col_offset):
    d = np.bitwise_count(q[:, None, :] ^ t_local[None, :, :]).sum(axis=2).astype(np.int64)
    n1, n2l = d.shape
    R = np.full(n1, NONE_KEY, dtype=np.int64)
    P = np.full(n1, NONE_KEY, dtype=np.int64)
    if n2l:
        keys_rows = (d << 20) | (np.arange(n2l)[None, :] + col_offset)
        R = keys_rows.min(axis=1)
        keys_cols = (d << 20) | np.arange(n1)[:, None]                    # a column's key over the rows: (d, i)
        best_row = (keys_cols.min(axis=0) & 0xFFFFF).astype(np.int64)       # each local column's choice is final
        for j in range(n2l):
            i = best_row[j]
            P[i] = min(P[i], (d[i, j] << 20) | (j + col_offset))
    return R.astype(np.int32), P.astype(np.int32)


def _rp_worker(rank, world, port, q, t, out_q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.train_slices(len(t), world)[rank]
        R, P = _local_rp(q, t[lo:hi], lo)
        x = torch.from_numpy(np.concatenate([R, P]))
        dist.all_reduce(x, op=dist.ReduceOp.MIN)
        if rank == 0:
            out_q.put(x.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_rp_exchange_rule_world2_gloo():
    import torch.multiprocessing as mp
    q = synthetic.uniform_descriptors(3, 90, 16)            # 16-bit descriptors: plenty of tied distances
    t = synthetic.uniform_descriptors(4, 71, 16)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_rp_worker, args=(r, 2, port, q, t, queue)) for r in range(2)]
    for p in procs:
        p.start()
    x = queue.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n1 = len(q)
    R, P = x[:n1].astype(np.int64), x[n1:].astype(np.int64)
    assert (P >= R).all()
    d = np.bitwise_count(q[:, None, :] ^ t[None, :, :]).sum(axis=2).astype(np.int64)
    row_best = ((d << 20) | np.arange(len(t))[None, :]).min(axis=1)            # (d, j) order
    col_best = ((d << 20) | np.arange(n1)[:, None]).min(axis=0)                # (d, i) order
    mutual = np.array([(col_best[row_best[i] & 0xFFFFF] & 0xFFFFF) == i for i in range(n1)])
    assert (R == row_best).all()
    assert ((R == P) == mutual).all() and mutual.sum() >= 1
    # and those are the first accepts of the reference's greedy loop: every mutual pair is in the literal result
    exp = {(int(a), int(b)) for a, b, _ in orc.match_literal(q, t, kernighan=True)[:min(n1, len(t))]}
    assert all((i, int(R[i] & 0xFFFFF)) in exp for i in np.flatnonzero(mutual))
