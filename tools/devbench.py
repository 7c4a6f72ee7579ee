"""Developer micro-bench (not the contract bench): times the C-ABI entry points on one GPU."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from photogrammetry_b200 import synthetic
from photogrammetry_b200.keypoint_matching import Matcher

def timeit(fn, stream, warm=3, reps=10):
    for _ in range(warm): fn()
    stream.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

def main():
    sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["8192"])]
    m = Matcher(0)
    stream = torch.cuda.Stream()
    m.set_stream(stream.cuda_stream)
    popc, lop3 = m.measure_popc_peak(300)
    print(json.dumps({"popc32_per_s": popc, "lop3_per_s": lop3, "evals256_roofline": popc / 8}))
    dists = sys.argv[2].split(",") if len(sys.argv) > 2 else ["U", "C"]
    ops = sys.argv[3].split(",") if len(sys.argv) > 3 else ["knn", "greedy", "host"]
    for n in sizes:
        for dist in dists:
            q, t = synthetic.config2_pair(n, dist)
            with torch.cuda.stream(stream):
                dq = torch.from_numpy(q).cuda(non_blocking=False); dt = torch.from_numpy(t).cuda()
                out = torch.empty((7, n), dtype=torch.int32, device="cuda")
            stream.synchronize()
            lib, h = m._lib, m._h
            def knn():
                lib.pgm_knn2_hamming_dev(h, dq.data_ptr(), n, dt.data_ptr(), n, 256, 32, out[3].data_ptr(), out[4].data_ptr(), out[5].data_ptr(), out[6].data_ptr())
            def greedy():
                m.match_greedy_dev(dq.data_ptr(), n, dt.data_ptr(), n, 256, 32, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), n)
            def host():
                m.match_greedy(q, t, 256)
            if "knn" in ops:
              med, mn = timeit(knn, stream)
              print(json.dumps({"n": n, "dist": dist, "op": "knn2_dev", "ms_med": med, "ms_min": mn, "evals_per_s": n * n / (mn * 1e-3), "frac_popc": n * n * 8 / (mn * 1e-3) / popc}))
            if "greedy" in ops:
              med, mn = timeit(greedy, stream, warm=3, reps=4)
              st = m.stats()
              print(json.dumps({"n": n, "dist": dist, "op": "greedy_dev", "ms_med": med, "ms_min": mn, "evals_per_s": n * n / (med * 1e-3), "stats": st}))
            if "host" not in ops: continue
            t0 = time.perf_counter(); 
            for _ in range(5): host()
            dt_ms = (time.perf_counter() - t0) / 5 * 1e3
            print(json.dumps({"n": n, "dist": dist, "op": "greedy_host_e2e", "ms": dt_ms, "evals_per_s": n * n / (dt_ms * 1e-3)}))
main()
