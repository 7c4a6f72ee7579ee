"""BASELINE configs[2] (proxy D5): the K = 32 shifted-star sequence, detector -> BRIEF -> matcher on the
device, every consecutive pair compared with the oracle chain (oracle/detect_np.py + orc.match_sweep +
orc.match_ratio_crosscheck).  Bit-exact: keypoint lists, descriptor bytes, greedy triples in the reference's order
(KeypointMatching.cs:38-66), and the ratio / cross-check survivors."""
import numpy as np
import pytest

from oracle import detect_np as D
from oracle import orc
from photogrammetry_b200 import sequence
from photogrammetry_b200.descriptors import pack_descriptors

pytestmark = pytest.mark.gpu


def test_star_sequence_against_oracle_chain(matcher):
    import torch
    frames = sequence.shifted_frames(sequence.star_gray(), 32, 5).astype(np.float32)
    pairs = sequence.star_pairs()
    d_frames = torch.from_numpy(frames).cuda()
    res = sequence.match_sequence_dev(matcher, d_frames, 50.0, pairs)
    table = D.py_pairs_to_xy(pairs)
    descs = []
    for k in range(32):                           # the oracle's detector + BRIEF (Python generation) per frame
        xy, _ = D.detect_vectorised(frames[k], 50.0, True)
        assert int(res.counts[k]) == len(xy)
        got_xy = res.xy[res.offsets[k]:res.offsets[k + 1]].cpu().numpy()
        assert (got_xy == xy).all()
        exp_desc = pack_descriptors(D.brief_descriptors(frames[k], xy, table, lsb_first=True), 256)
        got_desc = res.desc[res.offsets[k]:res.offsets[k + 1]].cpu().numpy()
        assert (got_desc == exp_desc).all()
        descs.append(exp_desc)
    assert int(res.counts[0]) == 128 and int(res.counts[30]) == 100     # the counts the reference quotes (keypoint_detection.py:158)
    greedy = res.greedy.cpu().numpy()
    n_real = 0
    for k in range(31):
        q, t = descs[k], descs[k + 1]
        exp = orc.match_literal(q, t, kernighan=True) if len(q) * len(t) <= 40000 else orc.match_sweep(q, t)
        got = greedy[:, res.greedy_starts[k]:res.greedy_starts[k] + res.greedy_counts[k]].T
        assert got.shape == exp.shape and (got == exp).all(), k
        n_real += min(len(q), len(t))
        expf = orc.match_ratio_crosscheck(q, t, 0.8, True)
        gotf = res.filtered[k].cpu().numpy().T
        assert gotf.shape == expf.shape and (gotf == expf).all(), k
    assert n_real > 0


def test_batched_detector_equals_per_image_calls(matcher):
    import torch
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(5, 61, 83)).astype(np.float32)
    pairs = D.gaussian_pairs(3, 256, 12).reshape(-1, 4).astype(np.int32)
    for pg in (False, True):
        th = 40.0
        xy, sc, desc, counts = matcher.detect_describe_batch_dev(torch.from_numpy(frames).cuda(), th, pairs, capacity=4096,
                                                                 python_generation=pg)
        for k in range(5):
            exy, esc = matcher.fast_detect(frames[k], th, python_generation=pg)
            assert counts[k] == len(exy)
            assert (xy[k, :len(exy)].cpu().numpy() == exy).all() and (sc[k, :len(exy)].cpu().numpy() == esc).all()
            edesc = matcher.brief_describe(frames[k], exy, pairs, python_generation=pg)
            assert (desc[k, :len(exy)].cpu().numpy() == edesc).all()
        # truncation is reported, not silent
        _, _, _, c2 = matcher.detect_describe_batch_dev(torch.from_numpy(frames).cuda(), th, pairs, capacity=3, python_generation=pg)
        assert (c2 == counts).all()


def test_ratio_crosscheck_batch_equals_per_pair_calls(matcher):
    import torch
    rng = np.random.default_rng(9)
    sizes = [37, 1, 64, 200, 5, 129]
    for bits, stride in ((256, 32), (100, 16), (512, 64)):
        desc = [rng.integers(0, 256, size=(n, stride), dtype=np.uint8) for n in sizes]
        if bits == 100:
            for d in desc:
                d[:, 13:] = 0
                d[:, 12] &= 0x0F
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        pairs = np.array([(0, 1), (1, 0), (2, 3), (3, 2), (4, 5), (5, 5), (3, 3)], dtype=np.int32)
        d_all = torch.from_numpy(np.concatenate(desc)).cuda()
        for ratio, cc, md in ((0.8, True, -1), (0.0, True, -1), (0.9, False, 120), (0.0, False, -1)):
            out, starts, counts = matcher.match_ratio_crosscheck_batch_dev(d_all, offs, pairs, bits, ratio, cc, md)
            out = out.cpu().numpy()
            for p, (a, b) in enumerate(pairs):
                exp = matcher.match_ratio_crosscheck(desc[a], desc[b], ratio, cc, md, bits)
                got = out[:, starts[p]:starts[p] + counts[p]].T
                assert got.shape == exp.shape and (got == exp).all(), (bits, ratio, cc, md, p)
                assert (exp == orc.match_ratio_crosscheck(desc[a], desc[b], ratio, cc, md)).all()
