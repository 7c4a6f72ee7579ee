"""The BASELINE configs at (or near) full size inside `pytest -m gpu`: checked against the oracle where
it finishes in seconds, and against the unsharded GPU call plus size-independent properties at the full 200k x 200k."""
import numpy as np
import pytest

from oracle import orc
from photogrammetry_b200 import sharding, synthetic
from photogrammetry_b200._lib import pinned_empty

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _properties(tr, q, t):
    """Size-independent properties of a greedy result on a square pair: a permutation on both sides, strictly increasing
    (distance, i, j), recomputed distances, first triple = the global minimum of the sampled rows."""
    n = len(q)
    assert tr.shape == (n, 3)
    assert (np.sort(tr[:, 0]) == np.arange(n)).all() and (np.sort(tr[:, 1]) == np.arange(len(t))).all()
    key = tr[:, 2].astype(np.int64) * (1 << 40) + tr[:, 0].astype(np.int64) * (1 << 20) + tr[:, 1]
    assert (np.diff(key) > 0).all()
    idx = np.linspace(0, n - 1, 4000).astype(np.int64)
    d = np.bitwise_count(q[tr[idx, 0]] ^ t[tr[idx, 1]]).sum(axis=1)
    assert (d == tr[idx, 2]).all()


def test_config3_200k_eight_emulated_shards_equal_unsharded(matcher):
    import torch
    n = 200_000
    q = synthetic.uniform_descriptors(1234, n, 256)
    t = synthetic.uniform_descriptors(5678, n, 256)
    d_q, d_t = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    ref = torch.empty((3, n), dtype=torch.int32, device="cuda")
    with matcher.torch_ordered():
        matcher.match_greedy_dev(d_q.data_ptr(), n, d_t.data_ptr(), n, 256, 32, ref[0].data_ptr(), ref[1].data_ptr(),
                                 ref[2].data_ptr(), n)
    torch.cuda.synchronize()
    exp = ref.T.contiguous().cpu().numpy()
    _properties(exp, q, t)
    outs, rounds = sharding.match_train_sharded_emulated(matcher, d_q, d_t, 8)
    assert rounds > 5
    for o in outs:
        assert (o == exp).all()
    # the library-owned loop (batched rounds, shrinking exchange), world of one
    mg = sharding.MultiGpuMatcher(matcher, 0, 1)
    got = mg.match_train_sharded(d_q, d_t, 0, n).T.contiguous().cpu().numpy()
    mg.close()
    assert (got == exp).all()


def test_config3_60k_emulated_shards_against_the_rounds_oracle(matcher):
    n1, n2 = 60_000, 50_000
    q = synthetic.uniform_descriptors(21, n1, 256)
    t = synthetic.noisy_copy_descriptors(22, q, 256)[:n2]
    exp = orc.match_rounds(q, t)
    outs, _ = sharding.match_train_sharded_emulated(matcher, q, t, 8)
    for o in outs:
        assert o.shape == exp.shape and (o == exp).all()
    assert (matcher.match_greedy(q, t, 256) == exp).all()


def test_config4_allpairs_64_images_sampled_pairs_against_the_sweep_oracle(matcher):
    import torch
    n_img, per = 64, 4096
    imgs = np.concatenate([synthetic.uniform_descriptors(9000 + k, per, 256) for k in range(n_img)])
    offs = np.arange(n_img + 1, dtype=np.int64) * per
    pairs = sharding.all_pairs(n_img)                                  # 2016 pairs
    d_all = torch.from_numpy(imgs).cuda()
    out = torch.empty((3, len(pairs) * per), dtype=torch.int32, device="cuda")
    with matcher.torch_ordered():
        counts = matcher.match_pairs_batch_dev(d_all.data_ptr(), offs, pairs, 256, 32, out[0].data_ptr(), out[1].data_ptr(),
                                               out[2].data_ptr(), len(pairs) * per)
    torch.cuda.synchronize()
    st = matcher.stats()
    assert (counts == per).all()
    assert st["evals_computed"] < 1.25 * st["distance_evals"]          # the candidate edges keep the recompute factor low
    tr = out.T.contiguous().cpu().numpy().reshape(len(pairs), per, 3)
    assert (np.sort(tr[:, :, 0], axis=1) == np.arange(per)).all() and (np.sort(tr[:, :, 1], axis=1) == np.arange(per)).all()
    key = tr[:, :, 2].astype(np.int64) * (1 << 40) + tr[:, :, 0].astype(np.int64) * (1 << 20) + tr[:, :, 1]
    assert (np.diff(key, axis=1) > 0).all()
    for p in np.linspace(0, len(pairs) - 1, 9).astype(int):
        a, b = pairs[p]
        assert (tr[p] == orc.match_sweep(imgs[offs[a]:offs[a + 1]], imgs[offs[b]:offs[b + 1]])).all(), p


def test_config4_page_locked_outputs_over_three_chunks(matcher):
    import torch
    n_img, per = 110, 4096                                             # 5995 pairs = 3 chunks of the batch engine
    imgs = np.concatenate([synthetic.uniform_descriptors(7000 + k, per, 256) for k in range(n_img)])
    offs = np.arange(n_img + 1, dtype=np.int64) * per
    pairs = sharding.all_pairs(n_img)
    rows = len(pairs) * per
    p_imgs = pinned_empty(imgs.shape, np.uint8)
    p_imgs[:] = imgs
    p_out = pinned_empty((3, rows), np.int32)
    p_out[:] = -7
    soa, starts, counts = matcher.match_pairs_batch(p_imgs, offs, pairs, 256, out=p_out)
    assert soa.shape == (3, rows) and (counts == per).all() and (starts == np.arange(len(pairs)) * per).all()
    d_all = torch.from_numpy(imgs).cuda()
    d_out = torch.empty((3, rows), dtype=torch.int32, device="cuda")
    with matcher.torch_ordered():
        matcher.match_pairs_batch_dev(d_all.data_ptr(), offs, pairs, 256, 32, d_out[0].data_ptr(), d_out[1].data_ptr(),
                                      d_out[2].data_ptr(), rows)
    torch.cuda.synchronize()
    assert (d_out.cpu().numpy() == soa).all()
    for p in (0, 2928, 2929, 2930, 5857, 5858, 5994):                  # first / last pairs of every chunk
        a, b = pairs[p]
        exp = orc.match_sweep(imgs[offs[a]:offs[a + 1]], imgs[offs[b]:offs[b + 1]])
        assert (soa[:, p * per:(p + 1) * per].T == exp).all(), p


def test_padding_bits_above_desc_bits_do_not_corrupt_the_ordering(matcher):
    # 32-byte rows passed with desc_bits = 128 whose upper half is NOT zero: distances range up to 256; the
    # ordering histograms are sized by the row width, so the result is the 256-bit answer instead of corrupted memory
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, size=(700, 32), dtype=np.uint8)
    t = rng.integers(0, 256, size=(640, 32), dtype=np.uint8)
    exp = orc.match_sweep(q, t)
    assert (matcher.match_greedy(q, t, 128) == exp).all()
    rows = matcher.match_keypoints_sorted(q[:50], t, 128)
    d = np.bitwise_count(q[:50, None, :] ^ t[None, :, :]).sum(axis=2)
    assert (rows[:, :, 1] == np.sort(d, axis=1)).all()
