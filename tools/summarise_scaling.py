"""Summarise gpurun_out/scale_N.jsonl (tools/run_scaling.sh) into one table / profiles/r02_scaling.json."""
import glob, json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"what": "bench.py --gpus N (configs[4] headline, configs[3] extra) and tools/bench_sharded.py --check per N, one box per N",
       "runs": {}}
for path in sorted(glob.glob(os.path.join(root, "gpurun_out", "scale_*.jsonl"))):
    for ln in open(path):
        d = json.loads(ln)
        if "metric" in d:
            n = d["n_gpus"]
            x = d.get("extras", {}).get("configs3_200k_train_sharded", {})
            out["runs"].setdefault(str(n), {})["bench"] = {
                "value": d["value"], "ms_per_step": d["ms_per_step"], "steps": d["steps"], "e2e_value": d["e2e"]["value"],
                "e2e_ms_per_step": d["e2e"]["ms_per_step"], "whole_call_frac": d["roofline"]["whole_call_frac"],
                "round0_frac": d["roofline"]["frac"], "recompute_factor": d["recompute_factor"], "clocks": d["clocks"],
                "configs3_extra": {k: x.get(k) for k in ("ms", "evals_per_s", "rounds", "collectives", "exchange_bytes_per_rank", "properties_ok")},
                "configs1_extra_ms": d.get("extras", {}).get("configs1_pair_8192_U", {}).get("ms"),
                "configs2_extra_ms": d.get("extras", {}).get("configs2_star_sequence_K32", {}).get("ms_per_sequence")}
        elif "workload" in d:
            n = int(d["workload"].rsplit("x", 1)[1])
            out["runs"].setdefault(str(n), {})["sharded_check"] = {k: d.get(k) for k in (
                "ms", "evals_per_s", "rounds", "collectives", "exchange_bytes_per_rank", "properties_ok",
                "every_rank_bit_identical_to_unsharded", "knn2_top2_merge_bit_identical_to_unsharded")}
base = out["runs"].get("1", {}).get("bench")
for n, r in out["runs"].items():
    if base and "bench" in r:
        r["bench"]["speedup_vs_1gpu"] = r["bench"]["value"] / base["value"]
        r["bench"]["efficiency"] = r["bench"]["value"] / base["value"] / int(n)
json.dump(out, open(os.path.join(root, "profiles", "r02_scaling.json"), "w"), indent=1)
for n in sorted(out["runs"], key=int):
    r = out["runs"][n]
    b, s = r.get("bench", {}), r.get("sharded_check", {})
    print(f"N={n}: configs[4] {b.get('value', 0):.4g} evals/s (e2e {b.get('e2e_value', 0):.4g}), x{b.get('speedup_vs_1gpu', 0):.2f}; "
          f"configs[3] bench-extra {b.get('configs3_extra', {}).get('ms')} ms, check-run {s.get('ms')} ms, bit-identical {s.get('every_rank_bit_identical_to_unsharded')}")
