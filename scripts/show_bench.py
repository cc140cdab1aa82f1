#!/usr/bin/env python3
"""prints the interesting parts of a bench.py JSON line: show_bench.py <file>"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d.get("roofline", {})
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"))
print("kernels ms/step", {k: round(v, 1) for k, v in r.get("ms_by_kernel_per_step", {}).items()}, "dominant", r.get("kernel"), "frac", round(r.get("frac", 0), 3),
      "launch ms", round(r.get("avg_launch_ms", 0), 2), "traffic", r.get("traffic"), "step frac of HBM ceiling", round(r.get("step_frac_of_hbm_ceiling", 0), 3))
if "cpu_baseline" in d: print("cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"], "cores; 1 thread", round(d["cpu_baseline"].get("single_thread", {}).get("value", 0)))
if "mpc_step_latency" in d: print("mpc", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["mpc_step_latency"].items() if k != "config"})
if "config4" in d: c = d["config4"]; print("config4", round(c["value"]), "e2e", round(c["e2e"]["value"]), "ms", round(c["ms_per_step"], 1), "nccl MB/step", round(c["e2e"]["nccl_bytes_per_step"] / 1e6, 1))
if "config5" in d: print("config5", [(c["controller"], c["knots"], round(c["ocp_iterations_per_s"])) for c in d["config5"]["cases"]])
if "divergent_batch" in d: v = d["divergent_batch"]; print("divergent", round(v["value"]), v["iterations_per_ocp"], "batch its", v["batch_iterations"], "straggler", round(v["straggler_factor"], 2), v["batch_iterations_by_active_share"], "ms/batch-it", round(v["ms_per_batch_iteration"], 2)); print("   stream", round(v["stream"]["value"]), "uniform", round(v["uniform_yardstick"]["value"]), v["per_iteration_cost_vs_uniform"])
print("clocks", d.get("clocks"))
