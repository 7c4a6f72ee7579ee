import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from photogrammetry_b200 import synthetic
from photogrammetry_b200.keypoint_matching import Matcher
from oracle import orc
m = Matcher(0); stream = torch.cuda.Stream(); m.set_stream(stream.cuda_stream)
for n_img, per in [(17, 1000), (9, 3000), (3, 600)]:
    imgs = [synthetic.uniform_descriptors(700 + k, per, 256) for k in range(n_img)]
    all_desc = np.concatenate(imgs); offs = np.arange(n_img + 1, dtype=np.int64) * per
    pairs = np.array([(k, k + 1) for k in range(n_img - 1)], dtype=np.int32)      # consecutive frames (configs[2] shape)
    with torch.cuda.stream(stream):
        d_all = torch.from_numpy(all_desc).cuda(); d_o = torch.empty((3, len(pairs) * per), dtype=torch.int32, device="cuda")
    def run(): m.match_pairs_batch_dev(d_all.data_ptr(), offs, pairs, 256, 32, d_o[0].data_ptr(), d_o[1].data_ptr(), d_o[2].data_ptr(), len(pairs) * per)
    for _ in range(3): run()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20): run()
    e1.record(stream); e1.synchronize()
    got = d_o.T.cpu().numpy().reshape(len(pairs), per, 3)
    ok = all((got[p] == orc.match_sweep(imgs[a], imgs[b])).all() for p, (a, b) in enumerate(pairs))
    print(n_img - 1, "pairs of", per, "ms", e0.elapsed_time(e1) / 20, "ok", ok, m.stats()["kernel_launches"])
