import os
#!/usr/bin/env python3
"""profiles/ table + traffic JSON from an `ncu --set full` report holding one launch of each kernel family.
usage: ncu_table.py <report.ncu-rep> <out.md> <out.json> <workload> <batch> <T>"""
import csv, io, json, subprocess, sys
rep, out_md, out_json, workload, batch, T = sys.argv[1:7]
batch, T = int(batch), int(T)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
def col(r, k, scale=1.0):
    try: return float(r[hdr.index(k)].replace(",", "")) * scale
    except Exception: return float("nan")
def unit(k): return units[hdr.index(k)]
fam = {"node_calc_kernel": "calc_diff", "node_cost_kernel": "calc_diff", "node_diff_kernel": "calc_diff", "backward_kernel": "backward", "rollout_kernel": "rollout", "decide_kernel": "decide"}
seen = {}; lines = []; traffic = {}; flines = []; slines = []
stall_keys = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = name.split("<")[0].replace("void ", "").replace("empc::", "")
    if short == "rollout_kernel" and ", 8>" in name: continue
    if short in seen or short not in fam: continue
    seen[short] = 1
    def gb(k):
        v = col(r, k); u = unit(k)
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    dur = col(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1.0)
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    lines.append(f"| `{short}` | {dur:.2f} | {rd/1e9:.2f} | {wr/1e9:.2f} | {(rd+wr)/dur/1e6:.0f} | {int(col(r,'launch__registers_per_thread'))} | "
                 f"{col(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{col(r,'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {col(r,'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active'):.1f} | "
                 f"{col(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {col(r,'smsp__inst_executed.sum')/1e6:.0f} |")
    # FP64 work of the launch: thread-level DADD/DMUL/DFMA counts (rate per elapsed SMSP cycle x elapsed cycles) and the
    # tensor-path FLOPs of the DMMAs; algorithmic in the sense of "what this formulation executes", counted by the hardware
    cyc = col(r, "smsp__cycles_elapsed.sum") / max(col(r, "smsp__cycles_elapsed.sum") / col(r, "smsp__cycles_elapsed.avg"), 1.0) if "smsp__cycles_elapsed.avg" in hdr else col(r, "sm__cycles_elapsed.max")
    ops = {k: col(r, f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") * cyc for k in ("dadd", "dmul", "dfma")}
    tens = col(r, "sm__ops_path_tensor_src_fp64.sum")
    flops = ops["dadd"] + ops["dmul"] + 2 * ops["dfma"] + (tens if tens == tens else 0.0)
    nodes = batch * (T + 1)
    flines.append(f"| `{short}` | {ops['dadd']/1e9:.2f} | {ops['dmul']/1e9:.2f} | {ops['dfma']/1e9:.2f} | {tens/1e9:.1f} | {flops/1e9:.1f} | {flops/nodes/1e3:.1f} | {flops/dur/1e9:.2f} |")
    st = sorted(((col(r, k), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in stall_keys), reverse=True)
    st = [(v, n) for v, n in st if v == v and n != "selected"][:5]
    slines.append(f"| `{short}` | {sum(v for v, _ in st) + 1.0:.1f} | " + ", ".join(f"{n} {v:.2f}" for v, n in st) + " |")
    t = traffic.setdefault(fam[short], {"dram_bytes_per_launch": 0.0, "kernels": {}})
    t["dram_bytes_per_launch"] += rd + wr
    t["kernels"][short] = {"ms": dur, "dram_read_bytes": rd, "dram_write_bytes": wr}
open(out_md, "w").write("| kernel | ms | DRAM read GB | DRAM write GB | DRAM GB/s | regs | warps active % | issue active % | FP64 pipe % | DMMA pipe % | DRAM % of peak | M warp-instr |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n" + "\n".join(lines) + "\n")
open(out_md, "a").write("\nFP64 work per launch (hardware counters; DADD/DMUL/DFMA are thread-level instruction counts, DMMA column in FLOPs; "
                       "FLOPs = DADD + DMUL + 2 DFMA + DMMA FLOPs; per node = / (batch x (T+1)); rollout and decide touch each node once per trial):\n\n"
                       "| kernel | G DADD | G DMUL | G DFMA | G DMMA FLOP | GFLOP per launch | kFLOP per node | TFLOP/s |\n|---|---|---|---|---|---|---|---|\n" + "\n".join(flines) + "\n")
open(out_md, "a").write("\nWarp stall reasons (cycles a resident warp waits per instruction it issues, ncu `smsp__average_warps_issue_stalled_*_per_issue_active`; "
                       "top five besides the issue cycle itself):\n\n| kernel | ≈ cycles per issued instruction (top five + 1) | stalls |\n|---|---|---|\n" + "\n".join(slines) + "\n")
# stamped with the hash of the CUDA sources next to this script (the library under ncu was built from them): bench.py refuses
# the numbers when the sources have changed since (bench.py source_hash)
import hashlib
_h = hashlib.sha256()
_csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "eagle-mpc_b200", "csrc")
for _f in sorted(os.listdir(_csrc)):
    if _f.endswith((".cu", ".cuh")):
        _h.update(open(os.path.join(_csrc, _f), "rb").read())
json.dump({"workload": workload, "batch": batch, "T": T, "source": rep.split("/")[-1], "kernels": traffic, "source_hash": _h.hexdigest()[:16]},
          open(out_json, "w"), indent=1)
print(open(out_md).read())
