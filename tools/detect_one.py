#!/usr/bin/env python
"""Developer bench / profiling target for the keypoint producer: one 4000x3000 synthetic image (the size of the
reference's lego photos) of random overlapping rectangles -> FAST-12 -> NMS -> BRIEF -> match against a second
image shifted by 150 px (scripts/image_editing.py's construction).  Prints per-stage wall times of the C-ABI calls
(host buffers, copies included) and checks every stage against the oracle."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from photogrammetry_b200 import keypoint_detection as kd
from photogrammetry_b200.keypoint_matching import KeypointMatching, Matcher

check = "--check" in sys.argv
H, W = 3000, 4000
rng = np.random.default_rng(7)
img = np.full((H, W), 0.5, dtype=np.float32)
for _ in range(1500):
    x0, y0 = int(rng.integers(0, W - 40)), int(rng.integers(0, H - 40))
    w, h = int(rng.integers(20, 400)), int(rng.integers(20, 400))
    img[y0:y0 + h, x0:x0 + w] = np.float32(rng.random())
for _ in range(8000):                          # small blobs: every ring pixel differs from the centre
    x0, y0 = int(rng.integers(5, W - 8)), int(rng.integers(5, H - 8))
    img[y0:y0 + 3, x0:x0 + 3] = np.float32(rng.random())
img2 = np.zeros_like(img); img2[:, 150:] = img[:, :-150]
m = Matcher(0)
det = kd.KeypointDetection(kd.KeypointDetectionOptions(Threshold=0.1), seed=11, matcher=m)
out = {"image": f"{W}x{H} float32, 1500 random rectangles"}
for rep in range(3):
    t = {}
    t0 = time.perf_counter(); xy, sc = m.fast_detect(img, 0.1); t["fast_detect_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); kept = m.nms(xy, sc, 10); t["nms_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); desc = m.brief_describe(img, xy[kept], det._pair_table); t["brief_ms"] = (time.perf_counter() - t0) * 1e3
    xy2, sc2 = m.fast_detect(img2, 0.1); kept2 = m.nms(xy2, sc2, 10); desc2 = m.brief_describe(img2, xy2[kept2], det._pair_table)
    t0 = time.perf_counter(); tr = m.match_greedy(desc, desc2, 256); t["match_ms"] = (time.perf_counter() - t0) * 1e3
out.update(t)
out.update({"keypoints": int(len(xy)), "after_nms": int(len(kept)), "after_nms_2": int(len(kept2)),
            "mpix_per_s_fast": W * H / (t["fast_detect_ms"] * 1e-3) / 1e6,
            "matches_with_dx150": int(((xy2[kept2][tr[:, 1], 0] - xy[kept][tr[:, 0], 0] == 150) & (tr[:, 2] < 2**31 - 1)).sum()),
            "matched": int((tr[:, 2] < 2**31 - 1).sum())})
# the same chain with the image resident on the device and no intermediate leaving it
import torch
d_img = torch.from_numpy(img).cuda(); d_img2 = torch.from_numpy(img2).cuda()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gxy, gsc, gdesc = m.detect_describe_dev(d_img, 0.1, det._pair_table, 10, capacity=65536)
    gxy2, gsc2, gdesc2 = m.detect_describe_dev(d_img2, 0.1, det._pair_table, 10, capacity=65536)
    n1d, n2d = int(gdesc.shape[0]), int(gdesc2.shape[0])
    o = torch.empty((3, n1d), dtype=torch.int32, device="cuda")
    m.match_greedy_dev(gdesc.data_ptr(), n1d, gdesc2.data_ptr(), n2d, 256, 32, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), n1d)
    torch.cuda.synchronize(); dev_ms = (time.perf_counter() - t0) * 1e3
out["device_chain_two_images_plus_match_ms"] = dev_ms
out["device_chain_equals_host_calls"] = bool((gxy.cpu().numpy() == xy[kept]).all() and (gdesc.cpu().numpy() == desc).all()
                                             and (o.T.cpu().numpy() == tr).all())
if check:
    from oracle import detect_np as D, orc
    exy, esc = D.detect_vectorised(img, 0.1)
    ek = D.eliminate_redundant_vectorised(exy, esc, 10) if len(exy) < 200000 else None
    out["fast_equals_oracle"] = bool(len(exy) == len(xy) and (exy == xy).all() and (esc == sc).all())
    if ek is not None:
        out["nms_equals_oracle"] = bool(ek.tolist() == kept.tolist())
    from photogrammetry_b200.descriptors import pack_descriptors
    sub = kept[:200]
    out["brief_equals_oracle"] = bool((pack_descriptors(D.brief_descriptors(img, xy[sub], D.gaussian_pairs(11, 256, 50)), 256) == desc[:200]).all())
    out["match_equals_oracle"] = bool((orc.match_sweep(desc, desc2) == tr).all())
print(json.dumps(out))
