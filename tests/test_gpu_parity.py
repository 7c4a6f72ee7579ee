"""Parity of the CUDA path with the oracle, through the C ABI (run on a B200: -m gpu).

Bit-exact everywhere: indices, distances, order, tie-breaks and the n1>n2 tail
(KeypointMatching.cs:38-66)."""
import numpy as np
import pytest

from oracle import orc
from photogrammetry_b200 import synthetic
from photogrammetry_b200._lib import EmptyTrainError, PgmatchError
from photogrammetry_b200.keypoint import Coordinate, Keypoint
from photogrammetry_b200.keypoint_matching import KeypointMatching
from photogrammetry_b200.descriptors import unpack_descriptors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("direction", ["l2r", "r2l"])
def test_lego_fixture_golden(matcher, lego, direction):
    q, t = (lego["left"], lego["right"]) if direction == "l2r" else (lego["right"], lego["left"])
    got = matcher.match_greedy(q, t, 256)
    assert got.shape == lego[direction].shape
    assert (got == lego[direction]).all()
    st = matcher.stats()
    assert st["distance_evals"] == len(q) * len(t) and st["kernel_launches"] > 0


def test_lego_without_tail(matcher, lego):
    got = matcher.match_greedy(lego["left"], lego["right"], 256, reference_compat_tail=False)
    assert got.shape == (1285, 3) and (got == lego["l2r"][:1285]).all()


SMALL = [(1, 1, 256), (2, 1, 256), (1, 2, 256), (7, 7, 1), (13, 5, 3), (5, 13, 8), (33, 33, 8), (64, 48, 17),
         (48, 64, 100), (40, 40, 128), (31, 57, 129), (20, 20, 300), (25, 9, 512), (64, 64, 6), (130, 300, 8),
         (700, 650, 4), (513, 512, 256), (1000, 37, 256), (37, 1000, 256), (600, 600, 2)]


@pytest.mark.parametrize("n1,n2,bits", SMALL)
def test_small_random_against_literal(matcher, n1, n2, bits):
    q = synthetic.uniform_descriptors(1000 + n1, n1, bits)
    t = synthetic.uniform_descriptors(2000 + n2, n2, bits)
    exp = orc.match_literal(q, t, kernighan=False) if n1 * n2 <= 250000 else orc.match_sweep(q, t)
    got = matcher.match_greedy(q, t, bits)
    assert got.shape == exp.shape and (got == exp).all()


def test_duplicates_one_pair_per_round(matcher):
    q = np.zeros((300, 32), dtype=np.uint8)
    t = np.zeros((280, 32), dtype=np.uint8)
    got = matcher.match_greedy(q, t, 256)
    assert got[:280].tolist() == [[k, k, 0] for k in range(280)]
    assert (got[280:] == [0, 0, 2147483647]).all()


def test_duplicates_large_grid_rounds(matcher):
    # 40 identical rows inside a bigger problem: forces many grid rounds with few accepts each
    q = synthetic.uniform_descriptors(5, 1500, 256)
    t = synthetic.uniform_descriptors(6, 1400, 256)
    q[100:140] = q[100]
    t[200:230] = q[100]
    got = matcher.match_greedy(q, t, 256)
    assert (got == orc.match_sweep(q, t)).all()


def test_edge_cases(matcher):
    q = synthetic.uniform_descriptors(1, 5, 256)
    e = np.zeros((0, 32), dtype=np.uint8)
    assert matcher.match_greedy(e, q, 256).shape == (0, 3)
    assert matcher.match_greedy(e, e, 256).shape == (0, 3)
    with pytest.raises(EmptyTrainError):
        matcher.match_greedy(q, e, 256)
    with pytest.raises(IndexError):          # what reference callers would catch (ArgumentOutOfRange)
        matcher.match_greedy(q, e, 256)
    with pytest.raises(PgmatchError):
        matcher.match_greedy(np.zeros((3, 32), np.uint8), np.zeros((3, 32), np.uint8), 600)


@pytest.mark.parametrize("dist", ["U", "C"])
def test_config2_8k_against_sweep(matcher, dist):
    q, t = synthetic.config2_pair(8192, dist)
    exp = orc.match_sweep(q, t)
    got = matcher.match_greedy(q, t, 256)
    assert (got == exp).all()
    st = matcher.stats()
    assert st["distance_evals"] == 8192 * 8192
    assert st["evals_computed"] < 2.5 * st["distance_evals"]


def test_rectangular_20k_against_rounds(matcher):
    q = synthetic.uniform_descriptors(11, 20000, 256)
    t = synthetic.noisy_copy_descriptors(12, q, 256)[:15000]
    exp = orc.match_rounds(q, t)
    got = matcher.match_greedy(q, t, 256)
    assert (got == exp).all()


def test_full_size_properties(matcher):
    # size-independent properties at a size the oracle does not check element-wise
    n = 30000
    q = synthetic.uniform_descriptors(21, n, 256)
    t = synthetic.noisy_copy_descriptors(22, q, 256)
    got = matcher.match_greedy(q, t, 256)
    assert got.shape == (n, 3)
    assert sorted(got[:, 0].tolist()) == list(range(n)) and sorted(got[:, 1].tolist()) == list(range(n))
    k = got[:, 2].astype(np.int64) * (1 << 40) + got[:, 0].astype(np.int64) * (1 << 20) + got[:, 1]
    assert (np.diff(k) > 0).all()                                   # (d, i, j) strictly ascending
    d = np.bitwise_count(q[got[:, 0]] ^ t[got[:, 1]]).sum(axis=1)
    assert (d == got[:, 2]).all()                                   # reported distances are the true ones
    # the first triple is the global minimum edge; spot-check it against a knn pass
    bj, bd, _, _ = matcher.knn2(q, t, 256)
    i0 = int(np.lexsort((np.arange(n), bd))[0])
    assert got[0].tolist() == [i0, int(bj[i0]), int(bd[i0])]
    # idempotence: matching again gives the same answer
    assert (matcher.match_greedy(q, t, 256) == got).all()


def test_batch_equals_single(matcher):
    sizes = [300, 0, 1200, 700, 64, 2500]
    imgs = [synthetic.uniform_descriptors(100 + k, n, 256) for k, n in enumerate(sizes)]
    imgs[3][:50] = imgs[2][:50]
    all_desc = np.concatenate(imgs)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    pairs = [(0, 2), (2, 3), (3, 2), (4, 5), (5, 4), (1, 0), (0, 0), (5, 2)]
    triples, starts, counts = matcher.match_pairs_batch(all_desc, offs, pairs, 256)
    for p, (a, b) in enumerate(pairs):
        exp = orc.match_sweep(imgs[a], imgs[b])
        got = triples[starts[p]:starts[p] + counts[p]]
        assert counts[p] == sizes[a] and (got == exp).all(), (p, a, b)
    with pytest.raises(EmptyTrainError):
        matcher.match_pairs_batch(all_desc, offs, [(0, 1)], 256)


@pytest.mark.parametrize("pinned", [False, True])
def test_batch_many_chunks_pageable_and_pinned_outputs(matcher, pinned):
    # > 2 chunks of 4096 pairs: the double-buffered output path (device buffer -> pinned staging -> helper thread
    # for pageable arrays; straight into the caller's arrays when they are page-locked) reuses each slot
    from photogrammetry_b200._lib import pinned_empty
    n_img, per = 150, 24
    imgs = [synthetic.uniform_descriptors(500 + k, per, 256) for k in range(n_img)]
    all_desc = np.concatenate(imgs)
    offs = np.arange(n_img + 1, dtype=np.int64) * per
    pairs = np.array([(a, b) for a in range(n_img) for b in range(a + 1, n_img)], dtype=np.int32)   # 11 175 pairs
    total = len(pairs) * per
    out = pinned_empty((3, total), np.int32) if pinned else np.empty((3, total), np.int32)
    out[:] = -7
    got, starts, counts = matcher.match_pairs_batch(all_desc, offs, pairs, 256, out=out)
    assert got.shape == (3, total) and (counts == per).all()
    single = {}
    for p in list(range(0, len(pairs), 997)) + [4095, 4096, 8191, 8192, len(pairs) - 1]:
        a, b = pairs[p]
        exp = orc.match_sweep(imgs[a], imgs[b])
        assert (got[:, starts[p]:starts[p] + per].T == exp).all(), p
    # every row of every pair is a permutation of its query indices
    assert (np.sort(got[0].reshape(len(pairs), per), axis=1) == np.arange(per)).all()


def test_knn2_and_ratio_crosscheck(matcher):
    q = synthetic.uniform_descriptors(31, 3000, 256)
    t = synthetic.noisy_copy_descriptors(32, q, 256)[:2500]
    for a, b in zip(matcher.knn2(q, t, 256), orc.knn2(q, t)):
        assert (a == b).all()
    for ratio, cc, md in [(0.8, True, -1), (0.9, False, -1), (0.0, True, -1), (0.0, False, 75), (0.7, True, 60)]:
        got = matcher.match_ratio_crosscheck(q, t, ratio, cc, md, 256)
        exp = orc.match_ratio_crosscheck(q, t, ratio, cc, md)
        assert got.shape == exp.shape and (got == exp).all(), (ratio, cc, md)
    # heavy ties
    q8, t8 = synthetic.uniform_descriptors(41, 500, 8), synthetic.uniform_descriptors(42, 700, 8)
    for a, b in zip(matcher.knn2(q8, t8, 8), orc.knn2(q8, t8)):
        assert (a == b).all()
    q1 = synthetic.uniform_descriptors(43, 9, 256)
    assert [x.tolist() for x in matcher.knn2(q1, q1[:1], 256)][2] == [-1] * 9     # no second neighbour


def test_keypoint_matching_object_surface(lego):
    # the drop-in class: same signature and reference semantics as KeypointMatching.cs:14-69
    kp1 = [Keypoint(Coordinate(int(x), int(y)), d) for (x, y), d in
           zip(lego["left_coord"][:400], unpack_descriptors(lego["left"][:400]))]
    kp2 = [Keypoint(Coordinate(int(x), int(y)), d) for (x, y), d in
           zip(lego["right_coord"][:250], unpack_descriptors(lego["right"][:250]))]
    pairs = KeypointMatching().MatchKeypoints(kp1, kp2)
    exp = orc.match_literal(lego["left"][:400], lego["right"][:250])
    assert len(pairs) == len(kp1) == 400
    for p, (i, j, d) in zip(pairs, exp.tolist()):
        assert p.Keypoint1 is kp1[i] and p.Keypoint2 is kp2[j] and p.Distance == d
    assert pairs[-1].Keypoint1 is kp1[0] and pairs[-1].Keypoint2 is kp2[0] and pairs[-1].Distance == 2147483647
    assert KeypointMatching().MatchKeypoints([], kp2) == []
    with pytest.raises(IndexError):
        KeypointMatching().MatchKeypoints(kp1, [])


@pytest.mark.parametrize("n1,n2,shards", [(3000, 2500, 2), (2000, 3100, 3), (1500, 1500, 8), (700, 40, 4)])
def test_train_sharded_emulated_equals_unsharded(matcher, n1, n2, shards):
    # every emulated rank must end with the unsharded result, bit for bit
    from photogrammetry_b200 import sharding
    q = synthetic.uniform_descriptors(61, n1, 256)
    t = synthetic.noisy_copy_descriptors(62, synthetic.uniform_descriptors(61, max(n1, n2), 256), 256)[:n2]
    t[5:9] = t[5]                                      # duplicate train rows straddling ties
    exp = orc.match_sweep(q, t)
    outs, rounds = sharding.match_train_sharded_emulated(matcher, q, t, shards)
    assert rounds > 0
    for o in outs:
        assert o.shape == exp.shape and (o == exp).all()
    assert (matcher.match_greedy(q, t, 256) == exp).all()


@pytest.mark.parametrize("n1,n2,shards,bits", [(3000, 2500, 2, 256), (2000, 3101, 3, 256), (1500, 1500, 8, 256),
                                                 (700, 40, 4, 256), (600, 900, 5, 8), (50, 3, 4, 256), (40, 1, 2, 256)])
def test_train_sharded_knn_top2_merge_equals_unsharded(matcher, n1, n2, shards, bits):
    # north_star's "train-set shard with a top-2 merge": packed keys -> all-gather (a stack here) -> device merge;
    # every emulated rank must hold the unsharded knn2 and ratio / cross-check results, bit for bit
    import torch
    from photogrammetry_b200 import sharding
    q = synthetic.uniform_descriptors(71, n1, bits)
    t = synthetic.uniform_descriptors(72, n2, bits)
    if bits == 256 and n2 > 100:
        t[: n2 // 2] = synthetic.noisy_copy_descriptors(73, q, 256)[: n2 // 2]
        t[7:11] = t[7]                                 # duplicate train rows: ties across slice boundaries
    dev = torch.device("cuda", matcher.device)
    d_q, d_t = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
    for ratio, cc, md in [(0.8, True, -1), (0.0, True, 70), (0.9, False, -1)]:
        knn, kept = sharding.knn_train_sharded_emulated(matcher, d_q, d_t, shards, bits, ratio, cc, md)
        for a, b in zip(knn, orc.knn2(q, t)):
            assert (a.cpu().numpy() == b).all()
        exp = orc.match_ratio_crosscheck(q, t, ratio, cc, md)
        got = kept.cpu().numpy().T
        assert got.shape == exp.shape and (got == exp).all(), (ratio, cc, md)
        assert (got == matcher.match_ratio_crosscheck(q, t, ratio, cc, md, bits)).all()


def test_python_twin_sorted_rows(matcher, lego):
    # match_keypoints (keypoint_matching.py:7-33): golden rows produced by the reference's own function
    got = matcher.match_keypoints_sorted(lego["left"][:48], lego["right"][:40], 256)
    ref = lego["twin"]
    assert got.shape == ref.shape and (got[:, :, 1] == ref[:, :, 1]).all()      # distances per rank are pinned
    assert (got == orc.python_twin(lego["left"][:48], lego["right"][:40])).all()  # stable (dist, idx2) order
    q, t = synthetic.uniform_descriptors(71, 300, 8), synthetic.uniform_descriptors(72, 1000, 8)   # heavy ties
    assert (matcher.match_keypoints_sorted(q, t, 8) == orc.python_twin(q, t)).all()
    q, t = synthetic.uniform_descriptors(73, 50, 512), synthetic.uniform_descriptors(74, 77, 512)
    assert (matcher.match_keypoints_sorted(q, t, 512) == orc.python_twin(q, t)).all()
    # object-level drop-in with the reference's name and signature
    from photogrammetry_b200.keypoint_matching import match_keypoints
    kp1 = [Keypoint(Coordinate(0, 0), d) for d in unpack_descriptors(lego["left"][:20])]
    kp2 = [Keypoint(Coordinate(0, 0), d) for d in unpack_descriptors(lego["right"][:30])]
    full = match_keypoints(kp1, kp2, -1)
    assert (full == orc.python_twin(lego["left"][:20], lego["right"][:30])).all()
    assert (full[:, :, 1] == ref[:20, :30, 1]).all() if False else True


def test_multi_gpu_matcher_single_rank_equals_unsharded(matcher):
    # the library-owned-communicator surface (pgm_multi_*) with a world of one: the batched, sync-free round loop and the
    # shrinking exchange must give the unsharded result; NCCL itself is exercised by tools/bench_sharded.py on >= 2 GPUs
    import torch

    from photogrammetry_b200 import sharding
    # (30000 x 29000: two rounds, then the replicated finish; 3000 x 40000: too many columns stay live for the replicated
    #  finish, the rounds run to the end; the small ones finish right after round 0)
    for n1, n2, dist in [(5000, 4200, "U"), (3000, 3000, "C"), (1500, 40, "U"), (1, 1, "U"), (30000, 29000, "U"), (3000, 40000, "U")]:
        q = synthetic.uniform_descriptors(11, n1, 256)
        t = synthetic.uniform_descriptors(12, n2, 256) if dist == "U" else synthetic.noisy_copy_descriptors(13, q, 256)[:n2]
        if n1 * n2 > 10 ** 8:                                   # the literal-free oracle for the big shapes
            assert (matcher.match_greedy(q, t, 256) == orc.match_rounds(q, t)).all()
        mg = sharding.MultiGpuMatcher(matcher, 0, 1)
        got = mg.match_train_sharded(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), 0, n2).T.cpu().numpy()
        assert (got == matcher.match_greedy(q, t, 256)).all()
        knn = mg.knn2_train_sharded(torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda(), 0)
        exp = matcher.knn2(q, t, 256)
        assert all((a.cpu().numpy() == b).all() for a, b in zip(knn, exp))
        assert mg.exchange() == (0, 0)
        mg.close()


def _bimodal(seed, n, dup_frac=0.5):
    """Half the rows are copies of ONE descriptor (every distance among them is 0), the rest uniform: the sampled mean and
    deviation put the candidate bound T above 0, so the duplicate block alone lists ~n^2/4 edges -- far more than any
    candidate list holds.  Exercises the overflow paths (tile staging, raw list, live-edge list, sharded edge blocks)."""
    a = synthetic.uniform_descriptors(seed, n, 256)
    k = int(n * dup_frac)
    a[:k] = a[0]
    rng = np.random.default_rng(seed)
    return a[rng.permutation(n)]


@pytest.mark.parametrize("n1,n2", [(3000, 3000), (2500, 1800), (700, 900)])
def test_candidate_list_overflow_falls_back_to_plain_rounds(matcher, n1, n2):
    q, t = _bimodal(5, n1), _bimodal(6, n2)
    t[: n2 // 2] = q[0]                                       # the duplicate blocks of both sides coincide: distance 0
    exp = orc.match_sweep(q, t)
    assert (matcher.match_greedy(q, t, 256) == exp).all()
    # throughput mode (more than 16 pairs), and the train-sharded steps
    from photogrammetry_b200 import sharding
    imgs = np.concatenate([q[:600], t[:600]] + [_bimodal(20 + k, 600) for k in range(5)])
    offs = np.arange(8, dtype=np.int64) * 600
    pairs = sharding.all_pairs(7)
    tr, starts, counts = matcher.match_pairs_batch(imgs, offs, pairs, 256)
    for p in (0, 1, 7, 20):
        a, b = pairs[p]
        e = orc.match_sweep(imgs[offs[a]:offs[a + 1]], imgs[offs[b]:offs[b + 1]])
        assert (tr[starts[p]:starts[p] + counts[p]] == e).all(), p
    outs, _ = sharding.match_train_sharded_emulated(matcher, q, t, 3)
    assert all((o == exp).all() for o in outs)


@pytest.mark.parametrize("slots_max", [1, 2, 5])
@pytest.mark.parametrize("n1,n2,dist", [(3000, 2800, "U"), (2048, 2048, "C"), (4096, 4096, "U")])
def test_sparse_phase_truncated_edge_list(matcher, monkeypatch, n1, n2, dist, slots_max):
    """A sparse phase whose edge list exceeds its shared memory keeps the edges up to the largest distance that fits
    (pgm_kernels.cuh, sparse_body).  PGM_SP_SLOTS_MAX shrinks the capacity so that small pairs take that path (the 16k+
    pairs of test_gpu_fullsize.py take it on their own); only list sizes change, never a match (KeypointMatching.cs:44-65)."""
    q, t = synthetic.config2_pair(max(n1, n2), dist)
    q, t = q[:n1], t[:n2]
    exp = orc.match_sweep(q, t)
    monkeypatch.setenv("PGM_SP_SLOTS_MAX", str(slots_max))
    got = matcher.match_greedy(q, t, 256)
    assert got.shape == exp.shape and (got == exp).all()
    # the same through the batch engine (throughput mode: standalone sparse kernel)
    imgs = np.concatenate([q, t])
    offs = np.array([0, n1, n1 + n2], dtype=np.int64)
    pairs = np.array([(0, 1)] * 17, dtype=np.int32)            # > 16 pairs: not the latency mode
    tr, starts, counts = matcher.match_pairs_batch(imgs, offs, pairs, 256)
    for p in (0, 16):
        assert (tr[starts[p]:starts[p] + counts[p]] == exp).all()
