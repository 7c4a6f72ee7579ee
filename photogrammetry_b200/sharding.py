"""Multi-GPU sharding of the matching path -- host logic only.

Mode 1, many image pairs (BASELINE configs 3 and 5): image pairs are independent
units.  Every rank holds all descriptors (64 MB for 512 x 4096 x 32 B), takes a
contiguous block of the pair list balanced by estimated cost N_a * N_b, matches it
with its own ``Matcher`` and the per-pair results are gathered to rank 0.  There is
NO data-path collective; ``torch.distributed`` only moves the finished match lists.

One process per GPU (``torch.distributed``; NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

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
