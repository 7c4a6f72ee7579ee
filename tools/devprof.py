"""Developer probe: whole-call time and the standalone round-0 launch of one greedy pair (env knobs: PGM_*)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from photogrammetry_b200 import synthetic
from photogrammetry_b200.keypoint_matching import Matcher

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dist = sys.argv[2] if len(sys.argv) > 2 else "U"
m = Matcher(0)
stream = torch.cuda.Stream(); m.set_stream(stream.cuda_stream)
q, t = synthetic.config2_pair(n, dist)
with torch.cuda.stream(stream):
    dq, dt = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    out = torch.empty((3, n), dtype=torch.int32, device="cuda")
stream.synchronize()
def step():
    m.match_greedy_dev(dq.data_ptr(), n, dt.data_ptr(), n, 256, 32, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), n)
for _ in range(5): step()
stream.synchronize()
ts = []
FLUSH = os.environ.get("FLUSH") == "1"
if FLUSH:
    with torch.cuda.stream(stream):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in range(20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        if FLUSH: flush.fill_(k & 0xFF)
        e0.record(stream); step(); e1.record(stream)
    e1.synchronize(); ts.append(e0.elapsed_time(e1))
st = m.stats()
m.set_profiling(True)
r0 = []
for _ in range(5):
    step(); a, b = m.round_profile(); r0.append(float(a[0]) if len(a) else float("nan"))
m.set_profiling(False)
env = {k: v for k, v in os.environ.items() if k.startswith("PGM_")}
print(json.dumps({"env": env, "n": n, "dist": dist, "call_ms_med": float(np.median(ts)), "call_ms_min": float(np.min(ts)),
                  "round0_ms": float(np.median(r0[1:])), "rounds": st["rounds"], "evals_computed": st["evals_computed"]}))
