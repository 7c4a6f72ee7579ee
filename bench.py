#!/usr/bin/env python
"""bench.py -- headline benchmark of the descriptor-matching hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one synthetic pair, 8192 query x 8192 train
256-bit descriptors (distribution U: i.i.d. uniform, seeds 1234/5678), matched
with the reference's semantics (full Hamming matrix + greedy one-to-one
assignment, KeypointMatching.cs:14-69).  One step = one MatchKeypoints call on
that pair.  At N > 1 every rank matches its own pair of the same shape (image
pairs are independent units: no data-path collective, weak scaling).

metric  = distance evaluations per second: N1*N2 per pair (each (i,j) counted
          once, however many rounds recompute it) / device time.
value   = inputs and outputs resident in HBM (pgm_match_hamming_greedy_dev).
e2e     = the same through the host-buffer C-ABI call the reference-facing wrapper
          makes (pgm_match_hamming_greedy): H2D of both descriptor sets and D2H
          of the triples inside the timed region, wall clock.
roofline= integer pipe: one 256-bit distance = 8 POPC.32; the denominator is a
          POPC micro-benchmark run live on the same GPU (MEASURED_PEAKS.json has
          no integer-pipe figure).
cpu_baseline / --impl reference = the oracle's literal restatement of the C#
          algorithm (kind "port": the C# itself cannot run here, no dotnet).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "descriptor distance-evals/s (greedy MatchKeypoints, 256-bit Hamming)"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=8192, help="descriptors per image (both sides)")
    ap.add_argument("--dist", default="U", choices=["U", "C"])
    ap.add_argument("--cpu-sample", type=int, default=2048, help="rows/cols of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the all-pairs / knn extras")
    return ap.parse_args()


# --------------------------------------------------------------------------
# clocks (recipe: /opt/skills/guides/B200_PROFILING.md)
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the observed range (the region is short and bursty)
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# CPU baseline (the oracle: test infrastructure, used here only as the timed baseline)
# --------------------------------------------------------------------------
def cpu_literal_rate(q: np.ndarray, t: np.ndarray, sample: int, steps: int = 1):
    from oracle import orc
    qs, ts = q[:sample], t[:sample]
    best = None
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        orc.match_literal(qs, ts, kernighan=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(qs) * len(ts) / best, best


def run_reference(args, rank: int):
    """The reference's own CPU algorithm (oracle port, 1 thread like the C#) on a bounded sample."""
    if rank != 0:
        return
    from photogrammetry_b200 import synthetic
    q, t = synthetic.config2_pair(args.n, args.dist)
    s = min(args.cpu_sample, args.n)
    from oracle import orc
    for _ in range(min(args.warmup, 1)):
        orc.match_literal(q[:s], t[:s], kernighan=True)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        orc.match_literal(q[:s], t[:s], kernighan=True)
        times.append(time.perf_counter() - t0)
    T = float(np.sum(times))
    value = s * s * args.steps / T
    sample = (f"first {s}x{s} descriptors of the {args.n}x{args.n} pair; literal O(N^3) restatement of "
              f"KeypointMatching.cs:14-82 (Kernighan CountOnes, full matrix, N1 argmin scans), 1 thread like the C#; "
              f"evals/s falls further at the full size because the scan phase is cubic")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (XOR + popcount)", "data": "synthetic",
        "config": {"workload": f"configs[1]: single synthetic pair {args.n}x{args.n}, 256-bit, distribution {args.dist}",
                   "cpu_sample": s},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from photogrammetry_b200 import synthetic
    from photogrammetry_b200.keypoint_matching import Matcher

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n, bits, stride = args.n, 256, 32

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # every rank owns one pair of the configs[1] shape (rank 0: the configs[1] seeds themselves)
    if rank == 0:
        q, t = synthetic.config2_pair(n, args.dist, bits)
    else:
        q = synthetic.uniform_descriptors(1234 + 1000 * rank, n, bits)
        t = (synthetic.uniform_descriptors(5678 + 1000 * rank, n, bits) if args.dist == "U"
             else synthetic.noisy_copy_descriptors(42 + rank, q, bits))

    m = Matcher(local_rank)
    stream = torch.cuda.Stream(device=dev)
    m.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        d_q = torch.from_numpy(q).to(dev)
        d_t = torch.from_numpy(t).to(dev)
        d_out = torch.empty((3, n), dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    stream.synchronize()

    def step_dev():
        m.match_greedy_dev(d_q.data_ptr(), n, d_t.data_ptr(), n, bits, stride,
                           d_out[0].data_ptr(), d_out[1].data_ptr(), d_out[2].data_ptr(), n)

    # ---- value: device-resident, CUDA events on the launching stream, L2 flushed between steps
    for _ in range(max(args.warmup, 3)):
        step_dev()
    stream.synchronize()
    launches_per_step = m.stats()["kernel_launches"]
    rounds = m.stats()["rounds"]
    evals_computed = m.stats()["evals_computed"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush.fill_(k & 0xFF)                   # L2 flush, outside the timed interval
            ev[k][0].record(stream)
            step_dev()
            ev[k][1].record(stream)
    stream.synchronize()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    T_ms = max_over_ranks(float(np.sum(step_ms)))
    value = world * float(n) * n * args.steps / (T_ms * 1e-3)

    # ---- e2e: host buffers through the C ABI, wall clock, H2D + D2H inside.  Inputs and outputs live in
    # page-locked host memory (pgm_host_alloc), which the library copies from / to directly.
    from photogrammetry_b200._lib import pinned_empty
    pq, pt = pinned_empty(q.shape, np.uint8), pinned_empty(t.shape, np.uint8)
    pq[:] = q
    pt[:] = t
    pout = pinned_empty((3, n), np.int32)
    for _ in range(3):
        m.match_greedy(pq, pt, bits, out=pout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got = m.match_greedy(pq, pt, bits, out=pout)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    st_host = m.stats()
    # the same call with ordinary (pageable) numpy arrays, staged through the library's pinned buffers
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got_pageable = m.match_greedy(q, t, bits)
    e2e_pageable_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * float(n) * n * args.steps / e2e_s
    matched_per_s = world * float(min(n, n)) * args.steps / (T_ms * 1e-3)
    assert got.shape == (3, n) and got_pageable.shape == (n, 3) and (got.T == got_pageable).all()

    # ---- roofline of the dominant kernel (hamming_round_kernel), separate profiling pass
    popc_peak, lop3_peak = m.measure_popc_peak(300)
    m.set_profiling(True)
    prof_ms, prof_ev = [], []
    for _ in range(5):
        step_dev()
        a, b = m.round_profile()
        prof_ms.append(a); prof_ev.append(b)
    m.set_profiling(False)
    pm, pe = np.concatenate(prof_ms[1:]), np.concatenate(prof_ev[1:])
    # In latency mode the dominant kernel is launched once per step (the full N1 x N2 round); the
    # remaining short rounds run the same device code inside the persistent tail kernel.
    achieved = float(pe.sum() * 8 / (pm.sum() * 1e-3))                     # POPC.32-equivalents / s
    round_share = float(pm.sum() / len(prof_ms[1:]) / np.mean(step_ms))
    traffic = None
    tp = os.path.join(ROOT, "profiles", "round_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "int-pipe",
        "bound_note": "POPC.32 issue rate (SURVEY 8d); neither HBM nor the tensor pipe bounds this path",
        "kernel": "hamming_round_kernel<8> (round 0: all N1 x N2 distances + row/column argmin)",
        "achieved": achieved / 1e9, "peak": popc_peak / 1e9, "unit": "Gpopc32/s", "frac": achieved / popc_peak,
        "peak_source": "measured live: register-only POPC micro-benchmark on this GPU (pgm_measure_popc_peak); "
                       "nominal 148 SM x 16/clk x 1.965 GHz = 4654",
        "algorithmic": "8 POPC.32 per 256-bit distance (SURVEY 8d) x N1*N2 distances of the launch",
        "note": "the kernel uses a carry-save popcount (5 POPC + 14 LOP3 per distance), which is how frac can "
                "exceed 1.0 of the plain 8-POPC roofline; against the carry-save POPC bound (5 per distance) "
                "the same launch sits at frac_vs_carry_save_bound",
        "frac_vs_carry_save_bound": achieved * 5.0 / 8.0 / popc_peak,
        "lop3_peak_gops": lop3_peak / 1e9,
        "launch_ms": float(pm.mean()), "launches_per_step": int(len(pm) / len(prof_ms[1:])),
        "kernel_share_of_step": round_share,
        "hbm_gbs_for_context": float((2 * n * stride + 12 * n) / (np.mean(step_ms) * 1e-3) / 1e9),
        "traffic": traffic,
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": T_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 (XOR + popcount)", "data": "synthetic",
        "config": {"workload": f"configs[1]: single synthetic pair {n}x{n}, 256-bit, distribution {args.dist}, "
                               f"one pair per GPU per step", "n1": n, "n2": n, "desc_bits": bits,
                   "l2": "flushed between timed steps (256 MiB fill, outside the event interval); the 0.5 MB of "
                         "descriptors is L2-resident within a step by nature",
                   "parallelism": f"pair-sharded x{world}, no collective"},
        "matched_pairs_per_s": matched_per_s,
        "rounds_per_step": rounds, "evals_computed_per_step": evals_computed,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st_host["h2d_bytes"],
                "d2h_bytes_per_step": st_host["d2h_bytes"], "ms_per_step": 1e3 * e2e_s / args.steps,
                "host_memory": "page-locked inputs and outputs (pgm_host_alloc); copies inside the timed region",
                "pageable_value": world * float(n) * n * args.steps / e2e_pageable_s},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline,
        "clocks": clocks,
    }

    if rank == 0 and not args.no_extras:
        line["extras"] = extras(m, stream, dev, popc_peak)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = min(args.cpu_sample, n)
        rate, secs = cpu_literal_rate(q, t, s)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {s}x{s} descriptors of the pair, oracle literal restatement of KeypointMatching.cs:14-82 "
                      f"(Kernighan CountOnes, full matrix, cubic argmin scans), 1 thread like the C#, {secs:.1f} s"}
        try:
            from oracle import orc
            t0 = time.perf_counter()
            orc.match_sweep(q, t)
            dt = time.perf_counter() - t0
            line["cpu_fast"] = {"value": n * n / dt, "unit": UNIT, "cores": orc.num_threads(),
                                "what": "oracle counting-sort sweep (popcnt, OpenMP), full pair, same output"}
        except Exception as e:  # pragma: no cover
            line["cpu_fast"] = {"error": str(e)}
    m.close()
    if rank == 0:
        print(json.dumps(line), flush=True)


def extras(m, stream, dev, popc_peak):
    """Secondary figures on the same GPU: the nearest/second-nearest pass and a scaled-down all-pairs batch."""
    import torch

    from photogrammetry_b200 import synthetic
    out = {}
    n = 8192
    q, t = synthetic.config2_pair(n, "U")
    with torch.cuda.stream(stream):
        d_q, d_t = torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev)
        o = torch.empty((4, n), dtype=torch.int32, device=dev)
    stream.synchronize()
    lib, h = m._lib, m._h

    def knn():
        lib.pgm_knn2_hamming_dev(h, d_q.data_ptr(), n, d_t.data_ptr(), n, 256, 32, *(o[k].data_ptr() for k in range(4)))
    for _ in range(3):
        knn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        knn()
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out["knn2_8k"] = {"ms": ms, "evals_per_s": n * n / (ms * 1e-3), "frac_of_popc_peak": n * n * 8 / (ms * 1e-3) / popc_peak}

    # float descriptors (north_star extension): squared-L2 top-2 on tcgen05, whole call (operand split, GEMM with
    # the norms folded in, exact refinement), device resident; roofline = tensor pipe, measured bf16 peak
    peak_tf, peak_src = 1590.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    try:
        peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        peak_src = "MEASURED_PEAKS.json bf16_tflops (burst)"
    except Exception:
        pass
    for nf in (8192, 32768):
        dim = 128
        with torch.cuda.stream(stream):
            fq = torch.rand((nf, dim), device=dev)
            ft = torch.rand((nf, dim), device=dev)
            fj = torch.empty((2, nf), dtype=torch.int32, device=dev)
            fd = torch.empty((2, nf), device=dev)
        stream.synchronize()

        def l2():
            m._check(lib.pgm_knn2_l2_dev(h, fq.data_ptr(), nf, ft.data_ptr(), nf, dim, fj[0].data_ptr(), fd[0].data_ptr(),
                                         fj[1].data_ptr(), fd[1].data_ptr(), None))
        for _ in range(3):
            l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            l2()
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * 3 * dim * nf * nf / (ms * 1e-3) / 1e12           # three bf16 split terms of K = D each
        out[f"l2_knn2_{nf // 1024}k_d{dim}"] = {
            "ms": ms, "evals_per_s": nf * float(nf) / (ms * 1e-3), "launches": 3,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                         "peak_source": peak_src,
                         "algorithmic": "executed bf16 flops: 2 x 3 x D per distance (hi.hi + lo.hi + hi.lo); "
                                        "2 x D per distance in the fp32 sense is a third of this"}}
        del fq, ft, fj, fd

    # all-pairs (configs[4] scaled down): 48 images x 4096 descriptors, all i<j pairs, device resident
    n_img, per = 48, 4096
    imgs = np.concatenate([synthetic.uniform_descriptors(9000 + k, per, 256) for k in range(n_img)])
    offs = np.arange(n_img + 1, dtype=np.int64) * per
    pairs = np.array([(a, b) for a in range(n_img) for b in range(a + 1, n_img)], dtype=np.int32)
    with torch.cuda.stream(stream):
        d_all = torch.from_numpy(imgs).to(dev)
        d_o = torch.empty((3, len(pairs) * per), dtype=torch.int32, device=dev)
    stream.synchronize()

    def allpairs():
        m.match_pairs_batch_dev(d_all.data_ptr(), offs, pairs, 256, 32, d_o[0].data_ptr(), d_o[1].data_ptr(),
                                d_o[2].data_ptr(), len(pairs) * per)
    allpairs()
    stream.synchronize()
    t0 = time.perf_counter()
    allpairs()
    stream.synchronize()
    dt = time.perf_counter() - t0
    ev = len(pairs) * float(per) * per
    out["allpairs_48x4096"] = {"pairs": int(len(pairs)), "seconds": dt, "evals_per_s": ev / dt,
                               "matched_pairs_per_s": len(pairs) * per / dt, "stats": m.stats()}
    return out


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
