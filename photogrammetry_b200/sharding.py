"""Multi-GPU sharding of the matching path -- host logic only.

Mode 1, many image pairs (BASELINE configs 3 and 5): image pairs are independent
units.  Every rank holds all descriptors (64 MB for 512 x 4096 x 32 B), takes a
contiguous block of the pair list balanced by estimated cost N_a * N_b, matches it
with its own ``Matcher`` and the per-pair results are gathered to rank 0.  There is
NO data-path collective; ``torch.distributed`` only moves the finished match lists.

Mode 2, one huge pair (BASELINE configs[3], 200k x 200k): :class:`TrainShardedMatcher`.  Every rank
holds all queries and a contiguous slice of the train set; per round the library computes the local
row/column argmins and two ``min`` all-reduces over N1 packed keys (NCCL over NVLink) merge them.
:class:`TrainShardedKnn` is the same sharding for the nearest / second-nearest search with ratio test and
cross-check: one all-gather of packed (best, second) keys and a device-side top-2 merge.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def all_pairs(n_images: int) -> np.ndarray:
    """All (a, b) with a < b, row-major: the all-pairs workload of BASELINE configs[4]."""
    a, b = np.triu_indices(n_images, k=1)
    return np.stack([a, b], axis=1).astype(np.int32)


def consecutive_pairs(n_images: int) -> np.ndarray:
    """(k, k+1) for a frame sequence: BASELINE configs[2]."""
    k = np.arange(max(n_images - 1, 0), dtype=np.int32)
    return np.stack([k, k + 1], axis=1)


def pair_costs(pairs: np.ndarray, image_sizes: Sequence[int]) -> np.ndarray:
    sizes = np.asarray(image_sizes, dtype=np.int64)
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    return sizes[pairs[:, 0]] * sizes[pairs[:, 1]]


def partition_pairs(pairs: np.ndarray, image_sizes: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Split the pair list into ``world_size`` contiguous blocks of near-equal cost.

    Returns ``[(start, stop)] * world_size`` (some may be empty).  Contiguous blocks keep the
    gathered output in pair-list order; the split points are where the cumulative cost
    crosses k/world_size of the total.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    n = len(pairs)
    if n == 0:
        return [(0, 0)] * world_size
    cum = np.cumsum(pair_costs(pairs, image_sizes).astype(np.float64))
    total = cum[-1]
    bounds = [0]
    for k in range(1, world_size):
        if total <= 0:
            cut = (n * k) // world_size
        else:
            cut = int(np.searchsorted(cum, total * k / world_size, side="left")) + 1
            cut = min(max(cut, bounds[-1]), n)
        bounds.append(cut)
    bounds.append(n)
    return [(bounds[k], bounds[k + 1]) for k in range(world_size)]


def shard_for_rank(pairs: np.ndarray, image_sizes: Sequence[int], rank: int, world_size: int) -> Tuple[int, int]:
    return partition_pairs(pairs, image_sizes, world_size)[rank]


def match_pairs_sharded(matcher, all_desc: np.ndarray, image_offsets: Sequence[int], pairs: np.ndarray,
                        rank: int, world_size: int, desc_bits: int = 256, reference_compat_tail: bool = True):
    """This rank's share of the pair list, matched with ``matcher.match_pairs_batch``.

    Returns ``(start, stop, triples, starts, counts)`` for pairs[start:stop]."""
    sizes = np.diff(np.asarray(image_offsets, dtype=np.int64))
    start, stop = shard_for_rank(pairs, sizes, rank, world_size)
    triples, starts, counts = matcher.match_pairs_batch(all_desc, image_offsets, pairs[start:stop], desc_bits,
                                                        reference_compat_tail)
    return start, stop, triples, starts, counts


def gather_match_lists(local, rank: int, world_size: int, dst: int = 0):
    """Gather the per-rank results of :func:`match_pairs_sharded` to ``dst`` and stitch them in
    pair-list order.  Returns ``(triples, starts, counts)`` on ``dst`` and ``None`` elsewhere."""
    start, stop, triples, starts, counts = local
    if world_size == 1:
        return triples, starts, counts
    import torch.distributed as dist
    payload = (int(start), int(stop), np.ascontiguousarray(triples), np.asarray(starts), np.asarray(counts))
    gathered = [None] * world_size if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    gathered.sort(key=lambda x: x[0])
    all_triples, all_starts, all_counts, base = [], [], [], 0
    for s, e, tr, st, ct in gathered:
        all_triples.append(tr)
        all_starts.append(np.asarray(st, dtype=np.int64) + base)
        all_counts.append(ct)
        base += len(tr)
    return (np.concatenate(all_triples) if all_triples else np.zeros((0, 3), np.int32),
            np.concatenate(all_starts) if all_starts else np.zeros(0, np.int64),
            np.concatenate(all_counts) if all_counts else np.zeros(0, np.int32))


def merge_top2(keys_best: np.ndarray, keys_second: np.ndarray):
    """Top-2 merge across train shards (north_star: "NVLink top-2 merge").

    ``keys_*``: uint32/uint64 ``[n_shards, n1]`` packed (distance << 20 | global train index) keys,
    ``0xFFFFFFFF`` = absent.  The global best is the min of the shard bests; the global second is the
    min over all remaining candidates.  Integer min on packed keys implements the (distance, index)
    tie-break, which is why the same operator serves an NCCL ``min`` all-reduce."""
    kb = np.asarray(keys_best)
    ks = np.asarray(keys_second)
    order = np.argsort(kb, axis=0, kind="stable")
    best = np.take_along_axis(kb, order[:1], axis=0)[0]
    runner = np.take_along_axis(kb, order[1:2], axis=0)[0] if kb.shape[0] > 1 else np.full_like(best, np.iinfo(kb.dtype).max)
    second_of_winner = np.take_along_axis(ks, order[:1], axis=0)[0]
    return best, np.minimum(runner, second_of_winner)


def train_slices(n2_total: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous column blocks [lo, hi) of the train set, one per rank, sizes differing by at most 1."""
    base, rem = divmod(n2_total, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class TrainShardedMatcher:
    """MatchKeypoints on ONE pair whose train set is sharded over the ranks, with the exchanges
    left to the caller's transport (``pgm_shard_*``; :class:`MultiGpuMatcher` is the form where the library owns NCCL).

    ``d_q``: torch uint8 ``[n1, stride]`` (all queries, on this rank's GPU);
    ``d_t_local``: torch uint8 ``[n2_local, stride]`` = train rows ``[col_offset, col_offset + n2_local)``.
    ``reduce_min(tensor)``: in-place MIN all-reduce of an int32 tensor across the ranks (default
    ``torch.distributed.all_reduce(op=MIN)``); ``all_gather(tensor) -> tensor[world, ...]`` (default
    ``torch.distributed.all_gather_into_tensor``).  Two exchanges per round: ``x = [R | P]`` (``2 * bound`` keys) and the
    ranks' candidate edges (include/pgmatch.h).  Returns int32 ``[3, n1]`` (qi, tj, dist) in the reference's order --
    identical on every rank and bit-identical to the unsharded ``Matcher.match_greedy``.
    """

    def __init__(self, matcher, d_q, d_t_local, col_offset: int, n2_total: int, desc_bits: int = 256,
                 reduce_min=None, all_gather=None, world_size: Optional[int] = None):
        import ctypes as C

        import torch
        self._torch = torch
        self._m = matcher
        self._lib = matcher._lib
        self.n1, self.stride = int(d_q.shape[0]), int(d_q.shape[1])
        self.n2_local, self.n2_total = int(d_t_local.shape[0]), int(n2_total)
        self._keep = (d_q, d_t_local)
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.world_size = int(world_size)
        self._sh = C.c_void_p()
        with matcher.torch_ordered(d_q.device):
            matcher._check(self._lib.pgm_shard_create(
                matcher._h, d_q.data_ptr(), self.n1, d_t_local.data_ptr() if self.n2_local else None, self.n2_local,
                int(col_offset), self.n2_total, int(desc_bits), self.stride, C.byref(self._sh)))
        dev = d_q.device
        cap = C.c_int32(0)
        matcher._check(self._lib.pgm_shard_edge_capacity(self.n1, self.n2_total, self.world_size, C.byref(cap)))
        self.edge_cap = cap.value
        self.x = torch.empty(2 * self.n1, dtype=torch.int32, device=dev)      # [R | P], the first 2 * bound entries in use
        self.edges = torch.zeros(1 + self.edge_cap, dtype=torch.int64, device=dev)   # count, then (d << 40 | i << 20 | global j)
        self.bound = self.n1
        self.out = torch.empty((3, self.n1), dtype=torch.int32, device=dev)
        self._reduce = reduce_min or self._dist_reduce
        self._gather = all_gather or self._dist_gather
        self.rounds = 0

    def _dist_reduce(self, t):
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)

    def _dist_gather(self, t):
        import torch
        import torch.distributed as dist
        if self.world_size == 1:
            return t.unsqueeze(0)
        out = torch.empty((self.world_size,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.contiguous())
        return out

    # the three local steps, exposed so a single process can drive several emulated ranks in lock-step
    # (every step runs on torch's current stream -- Matcher.torch_ordered -- unless the caller bound a stream, so
    # the collectives torch enqueues between the steps are ordered against the library's kernels)
    def exchange_view(self):
        return self.x[:2 * self.bound]

    def step_round(self):
        with self._m.torch_ordered(self.x.device):
            self._m._check(self._lib.pgm_shard_round(self._sh, self.x.data_ptr(), self.bound))

    def step_commit(self):
        with self._m.torch_ordered(self.x.device):
            self._m._check(self._lib.pgm_shard_commit(self._sh, self.x.data_ptr(), self.bound, self.edges.data_ptr(), self.edge_cap))

    def step_finish_round(self, edges_all):
        """``edges_all``: int64 ``[n_ranks, 1 + edge_cap]`` -> (live rows, done).  The next round's exchange shrinks to the
        live rows (rounded up to 1024)."""
        import ctypes as C
        lr, done = C.c_int32(0), C.c_int32(0)
        edges_all = edges_all.contiguous()
        with self._m.torch_ordered(self.x.device):
            self._m._check(self._lib.pgm_shard_finish_round(self._sh, edges_all.data_ptr(), int(edges_all.shape[0]),
                                                            self.edge_cap, C.byref(lr), C.byref(done)))
        self.bound = min(self.bound, max(1024, (lr.value + 1023) // 1024 * 1024))
        return lr.value, bool(done.value)

    def finish(self, reference_compat_tail: bool = True):
        import ctypes as C
        cnt, rounds = C.c_int32(0), C.c_int32(0)
        with self._m.torch_ordered(self.x.device):
            self._m._check(self._lib.pgm_shard_finish(
                self._sh, self.out[0].data_ptr(), self.out[1].data_ptr(), self.out[2].data_ptr(),
                1 if reference_compat_tail else 0, C.byref(cnt), C.byref(rounds)))
        self.rounds = rounds.value
        return self.out[:, :cnt.value]

    def match(self, reference_compat_tail: bool = True):
        """Library steps and the collectives alternate on torch's current stream (or on the stream the caller bound
        with ``matcher.set_stream``, which must then also be torch's current stream)."""
        prev = self.n1 + 1
        while True:
            self.step_round()
            self._reduce(self.exchange_view())
            self.step_commit()
            live_rows, done = self.step_finish_round(self._gather(self.edges))
            if done:
                break
            if live_rows >= prev:          # every round accepts at least the global minimum edge
                raise RuntimeError("train-sharded matcher made no progress (internal error)")
            prev = live_rows
        return self.finish(reference_compat_tail)

    def close(self):
        if self._sh and self._sh.value:
            self._lib.pgm_shard_destroy(self._sh)
            self._sh = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class MultiGpuMatcher:
    """One rank of a multi-GPU job with the communicator inside the library (``pgm_multi_*``, NCCL over NVLink): what
    a P/Invoke host gets.  ``unique_id`` (128 bytes from :func:`multi_unique_id` on rank 0) reaches the other ranks
    by any channel -- here ``torch.distributed.broadcast`` if no id is given."""

    def __init__(self, matcher, rank: int, world_size: int, unique_id: Optional[bytes] = None):
        import ctypes as C
        self._m, self._lib = matcher, matcher._lib
        self.rank, self.world_size = int(rank), int(world_size)
        if unique_id is None and self.world_size > 1:
            unique_id = broadcast_unique_id(self._lib, self.rank)
        self._mh = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        matcher._check(self._lib.pgm_multi_create(matcher._h, buf, self.rank, self.world_size, C.byref(self._mh)))

    def match_train_sharded(self, d_q, d_t_local, col_offset: int, n2_total: int, desc_bits: int = 256,
                            reference_compat_tail: bool = True, out=None):
        """int32 ``[3, count]`` device tensor (qi, tj, dist) in the reference's order, identical on every rank."""
        import ctypes as C

        import torch
        n1, stride, n2_local = int(d_q.shape[0]), int(d_q.shape[1]), int(d_t_local.shape[0])
        if out is None:
            out = torch.empty((3, max(n1, 1)), dtype=torch.int32, device=d_q.device)
        cnt, rounds = C.c_int32(0), C.c_int32(0)
        with self._m.torch_ordered(d_q.device):
            self._m._check(self._lib.pgm_multi_match_train_sharded_dev(
                self._mh, d_q.data_ptr(), n1, d_t_local.data_ptr() if n2_local else None, n2_local, int(col_offset),
                int(n2_total), int(desc_bits), stride, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), n1,
                C.byref(cnt), 1 if reference_compat_tail else 0, C.byref(rounds)))
        self.rounds = rounds.value
        return out[:, :cnt.value]

    def knn2_train_sharded(self, d_q, d_t_local, col_offset: int, desc_bits: int = 256):
        import torch
        n1, stride, n2_local = int(d_q.shape[0]), int(d_q.shape[1]), int(d_t_local.shape[0])
        out = torch.empty((4, max(n1, 1)), dtype=torch.int32, device=d_q.device)
        with self._m.torch_ordered(d_q.device):
            self._m._check(self._lib.pgm_multi_knn2_train_sharded_dev(
                self._mh, d_q.data_ptr(), n1, d_t_local.data_ptr() if n2_local else None, n2_local, int(col_offset),
                int(desc_bits), stride, *(out[k].data_ptr() for k in range(4))))
        return tuple(out[k, :n1] for k in range(4))

    def exchange(self):
        """(bytes this rank contributed to collectives, number of collectives) of the last call."""
        import ctypes as C
        b, n = C.c_int64(0), C.c_int32(0)
        self._m._check(self._lib.pgm_multi_get_exchange(self._mh, C.byref(b), C.byref(n)))
        return b.value, n.value

    def close(self):
        if getattr(self, "_mh", None) and self._mh.value:
            self._lib.pgm_multi_destroy(self._mh)
            self._mh = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def multi_unique_id(lib=None) -> bytes:
    import ctypes as C

    from . import _lib
    lib = lib or _lib.load()
    buf = (C.c_uint8 * 128)()
    rc = lib.pgm_multi_unique_id(buf)
    if rc != 0:
        raise _lib.PgmatchError(rc, "pgm_multi_unique_id failed (NCCL not loadable?)")
    return bytes(buf)


def broadcast_unique_id(lib, rank: int) -> bytes:
    """Rank 0 makes the communicator id, the others receive it over the default torch.distributed group."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(multi_unique_id(lib)), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


class TrainShardedKnn:
    """Nearest / second-nearest neighbour, ratio test and cross-check on ONE pair whose train set is
    sharded over the ranks -- north_star's "train-set shard with a top-2 merge" (SURVEY.md section 8e).

    Every rank holds all queries ``d_q`` and the train rows ``[col_offset, col_offset + n2_local)`` of
    ``train_slices(n2_total, world_size)``.  ``knn2()``: local search -> packed keys -> ONE all-gather of
    ``[2, n1]`` int32 per rank -> the two smallest of the 2G keys per query.  ``match_ratio_crosscheck()``
    adds the column side: each rank finds the best query of its own train rows (it sees every query, so
    the search needs no exchange) and the per-slice results are all-gathered into ``col_best_i[n2_total]``.
    Results are identical on every rank and bit-identical to the unsharded ``Matcher.knn2`` /
    ``Matcher.match_ratio_crosscheck``.

    ``all_gather(tensor, tag) -> tensor[world, *tensor.shape]`` (tag: "keys" or "cols"); default:
    ``torch.distributed.all_gather_into_tensor`` on the default group (NCCL over NVLink on GPUs).
    ``matcher`` provides ``knn2_hamming_dev``, ``pack_top2_keys_dev``, ``merge_top2_dev`` and
    ``ratio_crosscheck_filter_dev`` (keypoint_matching.Matcher)."""

    def __init__(self, matcher, d_q, d_t_local, col_offset: int, n2_total: int, desc_bits: int = 256,
                 all_gather=None, world_size: Optional[int] = None):
        self._m = matcher
        self.d_q, self.d_t = d_q, d_t_local
        self.col_offset, self.n2_total, self.desc_bits = int(col_offset), int(n2_total), int(desc_bits)
        self.n1, self.n2_local = int(d_q.shape[0]), int(d_t_local.shape[0])
        self._gather = all_gather or self._dist_gather
        if world_size is None:
            import torch.distributed as dist
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.world_size = int(world_size)
        self._slices = train_slices(self.n2_total, self.world_size)
        self._pad = max(max(hi - lo for lo, hi in self._slices), 1)

    def _dist_gather(self, t, tag):
        import torch
        import torch.distributed as dist
        if self.world_size == 1:
            return t.unsqueeze(0)
        t = t.contiguous()                          # output = the ranks' tensors concatenated along dim 0
        out = torch.empty((self.world_size * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out.view((self.world_size,) + tuple(t.shape))

    def local_keys(self):
        """This rank's exchange keys, int32 ``[2, n1]``."""
        bj, bd, sj, sd = self._m.knn2_hamming_dev(self.d_q, self.d_t, self.desc_bits)
        return self._m.pack_top2_keys_dev(bj, bd, sj, sd, self.col_offset)

    def local_col_best(self):
        """Best query (under (distance, i)) of every LOCAL train row, padded with -1 to the largest slice."""
        import torch
        ci = self._m.knn2_hamming_dev(self.d_t, self.d_q, self.desc_bits)[0]
        out = torch.full((self._pad,), -1, dtype=torch.int32, device=self.d_q.device)
        out[:self.n2_local] = ci
        return out

    def knn2(self):
        return self._m.merge_top2_dev(self._gather(self.local_keys(), "keys"))

    def match_ratio_crosscheck(self, ratio: float = 0.8, cross_check: bool = True, max_dist: int = -1):
        """int32 ``[3, count]`` = (i, j1, d1) of the kept queries in ascending i (global train indices)."""
        import torch
        bj, bd, _, sd = self.knn2()
        col = None
        if cross_check:
            allc = self._gather(self.local_col_best(), "cols")
            col = torch.cat([allc[r, :hi - lo] for r, (lo, hi) in enumerate(self._slices)]).contiguous()
        return self._m.ratio_crosscheck_filter_dev(self.n2_total, bj, bd, sd, col, ratio, cross_check, max_dist)


def knn_train_sharded_emulated(matcher, d_q, d_t, n_shards: int, desc_bits: int = 256, ratio: float = 0.8,
                               cross_check: bool = True, max_dist: int = -1):
    """All ranks of TrainShardedKnn emulated in one process on one device: every rank's contribution is
    computed first, the all-gathers become stacks, then every rank merges.  Returns
    ((best_j, best_d, second_j, second_d), kept_triples) after asserting that all ranks agree."""
    import torch
    n2 = int(d_t.shape[0])
    ranks = [TrainShardedKnn(matcher, d_q, d_t[lo:hi], lo, n2, desc_bits, world_size=n_shards)
             for lo, hi in train_slices(n2, n_shards)]
    bus = {"keys": torch.stack([r.local_keys() for r in ranks])}
    if cross_check:
        bus["cols"] = torch.stack([r.local_col_best() for r in ranks])
    results = []
    for r in ranks:
        r._gather = lambda t, tag: bus[tag]
        results.append((r.knn2(), r.match_ratio_crosscheck(ratio, cross_check, max_dist)))
    for knn, kept in results[1:]:
        assert all(torch.equal(a, b) for a, b in zip(knn, results[0][0])) and torch.equal(kept, results[0][1])
    return results[0]


def match_train_sharded_emulated(matcher, q, t, n_shards: int, desc_bits: int = 256):
    """All ``n_shards`` ranks of the train-sharded mode emulated in ONE process on ONE GPU, run in
    lock-step with the all-reduce replaced by an element-wise minimum and the all-gather by a stack (SURVEY.md section
    4.4 item 5: no inter-dependent concurrent kernels).  ``q`` / ``t``: numpy arrays or device tensors.  For tests and
    single-GPU validation."""
    import torch
    dev = torch.device("cuda", matcher.device)
    d_q = q if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q)).to(dev)
    d_t = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)).to(dev)
    n1, n2 = int(d_q.shape[0]), int(d_t.shape[0])
    shards = [TrainShardedMatcher(matcher, d_q, d_t[lo:hi], lo, n2, desc_bits, reduce_min=lambda x: None,
                                  all_gather=lambda x: None, world_size=n_shards)
              for lo, hi in train_slices(n2, n_shards)]
    prev = n1 + 1
    while True:
        for s in shards:
            s.step_round()
        red = torch.stack([s.exchange_view() for s in shards]).min(dim=0).values
        for s in shards:
            s.exchange_view().copy_(red)
            s.step_commit()
        edges_all = torch.stack([s.edges for s in shards])
        live, done = None, None
        for s in shards:
            lr, dn = s.step_finish_round(edges_all)
            assert live is None or (live, done) == (lr, dn), "ranks disagree on the live rows"
            live, done = lr, dn
        if done:
            break
        if live >= prev:
            raise RuntimeError("train-sharded matcher made no progress (internal error)")
        prev = live
    outs = [s.finish().T.contiguous().cpu().numpy() for s in shards]
    rounds = shards[0].rounds
    for s in shards:
        s.close()
    return outs, rounds


def bench_train_sharded(matcher, stream, dev, rank: int, world: int, popc_peak: float, n: int = 200_000,
                        dist_name: str = "U", reps: int = 2) -> dict:
    """BASELINE configs[3]: one n x n pair, train set sharded over the ranks with the library-owned NCCL communicator
    (every rank must call this).  Device time of the slowest rank; at world == 1 the same call runs unsharded."""
    import torch
    import torch.distributed as dist

    from . import synthetic
    q = synthetic.uniform_descriptors(1234, n, 256)
    t = synthetic.uniform_descriptors(5678, n, 256) if dist_name == "U" else synthetic.noisy_copy_descriptors(42, q, 256)
    lo, hi = train_slices(n, world)[rank]
    mg = MultiGpuMatcher(matcher, rank, world)
    with torch.cuda.stream(stream):
        d_q = torch.from_numpy(q).to(dev)
        d_t = torch.from_numpy(np.ascontiguousarray(t[lo:hi])).to(dev)
        out = torch.empty((3, n), dtype=torch.int32, device=dev)
        mg.match_train_sharded(d_q, d_t, lo, n, out=out)                      # warm-up (allocations, NCCL channels)
        times = []
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mg.match_train_sharded(d_q, d_t, lo, n, out=out)
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(min(times))
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        xb, xn = mg.exchange()
        res = out.cpu().numpy()
    ok = bool((np.sort(res[0]) == np.arange(n)).all() and (np.sort(res[1]) == np.arange(n)).all())
    key = res[2].astype(np.int64) * (1 << 40) + res[0].astype(np.int64) * (1 << 20) + res[1]
    ok &= bool((np.diff(key) > 0).all())
    chk = np.linspace(0, n - 1, 2000).astype(np.int64)
    ok &= bool((np.bitwise_count(q[res[0][chk]] ^ t[res[1][chk]]).sum(axis=1) == res[2][chk]).all())
    rounds = mg.rounds
    mg.close()
    return {"workload": f"configs[3]: one {n}x{n} pair ({dist_name}), train set sharded x{world}", "ms": ms,
            "evals_per_s": float(n) * n / (ms * 1e-3), "frac_of_popc_peak_per_gpu": float(n) * n * 8 / (ms * 1e-3) / popc_peak / world,
            "rounds": rounds, "collective": "per round ncclAllReduce(min, uint32) of [R | P] = 2 x bound keys + ncclAllGather of the ranks' candidate edges, inside the library"
            if world > 1 else "none (1 rank)", "collectives": xn, "exchange_bytes_per_rank": xb,
            "exchange_bytes_per_round": (xb / xn) if xn else 0.0,
            "properties_ok": ok, "properties": "permutation of rows and columns, strict (d, i, j) order, recomputed distances"}
