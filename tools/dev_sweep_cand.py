"""Developer sweep (not the contract bench): single-pair latency against the candidate-list knobs
PGM_CAND_TARGET (edges per row, pass 0) and PGM_CAND_ROW_MAX (per-row cap of later passes).  Every setting's
triples are compared with the default setting's (the knobs only size lists, never change a result)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from photogrammetry_b200 import synthetic
from photogrammetry_b200.keypoint_matching import Matcher

def main():
    cases = [(8192, "U"), (8192, "C"), (4096, "U"), (16384, "U")]
    targets = sys.argv[1].split(",") if len(sys.argv) > 1 else ["6"]
    caps = sys.argv[2].split(",") if len(sys.argv) > 2 else ["none", "24", "16", "12", "8", "6"]
    m = Matcher(0)
    stream = torch.cuda.Stream()
    m.set_stream(stream.cuda_stream)
    for n, dist in cases:
        q, t = synthetic.config2_pair(n, dist)
        with torch.cuda.stream(stream):
            dq = torch.from_numpy(q).cuda(); dt = torch.from_numpy(t).cuda()
            out = torch.empty((3, n), dtype=torch.int32, device="cuda")
        stream.synchronize()
        def call():
            m.match_greedy_dev(dq.data_ptr(), n, dt.data_ptr(), n, 256, 32, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), n)
        ref = None
        for tg in targets:
            for cap in caps:
                os.environ["PGM_CAND_TARGET"] = tg
                if cap == "none": os.environ.pop("PGM_CAND_ROW_MAX", None)
                else: os.environ["PGM_CAND_ROW_MAX"] = cap
                for _ in range(3): call()
                stream.synchronize()
                ts = []
                for _ in range(12):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream); call(); e1.record(stream); e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                got = out.cpu().numpy().copy()
                if ref is None: ref = got
                call(); stream.synchronize(); st1 = m.stats()              # (statistics are per call)
                print(json.dumps({"n": n, "dist": dist, "target": tg, "row_max": cap, "ms_med": round(float(np.median(ts)), 4),
                                  "ms_min": round(float(np.min(ts)), 4), "same_as_default": bool((got == ref).all()),
                                  "passes": st1["rounds"],
                                  "recompute": round(st1["evals_computed"] / (n * n), 4)}), flush=True)
main()
