"""Developer probe: rounds / recompute factor of the batch entry point for a few batch sizes (env knobs: PGM_*)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from photogrammetry_b200 import sharding, synthetic
from photogrammetry_b200.keypoint_matching import Matcher

per = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = Matcher(0)
for n_img in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["6", "7", "10", "24"])]:
    imgs = np.concatenate([synthetic.uniform_descriptors(9000 + k, per, 256) for k in range(n_img)])
    offs = np.arange(n_img + 1, dtype=np.int64) * per
    pairs = sharding.all_pairs(n_img)
    d_all = torch.from_numpy(imgs).cuda()
    d_o = torch.empty((3, len(pairs) * per), dtype=torch.int32, device="cuda")
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.match_pairs_batch_dev(d_all.data_ptr(), offs, pairs, 256, 32, d_o[0].data_ptr(), d_o[1].data_ptr(), d_o[2].data_ptr(), len(pairs) * per)
        m.synchronize(); dt = time.perf_counter() - t0
    st = m.stats()
    print(json.dumps({"per": per, "pairs": len(pairs), "ms": dt * 1e3, "evals_per_s": len(pairs) * per * per / dt, "rounds": st["rounds"],
                      "recompute": st["evals_computed"] / st["distance_evals"], "launches": st["kernel_launches"], "syncs": st["host_syncs"]}))
