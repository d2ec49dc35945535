#!/usr/bin/env python3
"""tools/show_bench.py FILE -- one screen of a bench.py JSON line."""
import json
import sys

j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = j["roofline"]
print("HEAD %s: value %.1f e2e %.1f ms/step %.2f (e2e %.2f) | %s kms %.2f share %.2f all-kernels %.2f | launches %d" % (
    j["config"]["workload"][:28], j["value"], j["e2e"]["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], r["kernel"][:28],
    r["kernel_ms"], r["kernel_share_of_step"], r["all_kernels_ms_per_step"], j["gpu_launches"]))
print("  e2e phases", {k: round(v, 2) for k, v in j["e2e"]["phases_ms_last_step"].items()})
print("  parity", j.get("parity"), "| cpu", j.get("cpu_baseline", {}).get("value"), "| issue", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in j.get("issue", {}).items() if k in ("pair_evals_per_strand_nt", "frac_of_lane_peak")})
b = j.get("binary")
if b:
    print("  binary: %.2f G strand-nt/s wall %.2fs | %s" % (b.get("value", 0), b.get("wall_s", 0), b.get("driver_summary")))
    print("  binary prefix:", b.get("prefix"))
for e in j.get("per_config", []):
    print("%-13s value %7.2f e2e %7.2f ms %7.1f kernels %7.2f filter %s cands %d surv %d score %s parity %s/%s useful %.3f" % (
        e["config"], e["value"], e["e2e"]["value"], e["ms_per_step"], e["kernels_ms_per_step"],
        ("%.2f" % e["filter_ms_per_step"]) if e["filter_ms_per_step"] is not None else "-", e["candidates_per_step_rank0"],
        e["survivors_of_level0_rank0"], e.get("score_prescreen"), e["parity"].get("full_equal_across_paths"),
        e["parity"].get("prefix_equal_oracle"), e.get("useful_work", {}).get("frac_of_lane_peak", 0)))
