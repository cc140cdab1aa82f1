#!/usr/bin/env python3
"""Shared-memory / LSU wavefronts of one kernel attributed to CUDA source lines (ncu source page + nvdisasm line info).
usage: ncu_wavefronts.py <report.ncu-rep> <lib.sass from nvdisasm --print-line-info -c> <function-substring> <units (e.g. nodes)> [top]"""
import csv, collections, io, re, subprocess, sys
rep, sass, pat = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
prof = [r for r in rows[2:] if len(r) > col['Instructions Executed']]
lines = []; cur = None; fn = None
for l in open(sass):
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m: lines.append((cur, m.group(2)))
assert len(prof) == len(lines), (len(prof), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])  # shared wf, excessive, global sectors, local sectors, instrs
for p, (key, ins) in zip(prof, lines):
    a = agg[key]
    a[0] += int(p[col['L1 Wavefronts Shared']] or 0); a[1] += int(p[col['L1 Wavefronts Shared Excessive']] or 0)
    a[2] += int(p[col['L2 Theoretical Sectors Global']] or 0); a[3] += int(p[col['L2 Theoretical Sectors Local']] or 0)
    a[4] += int(p[col['Instructions Executed']] or 0)
tot = [sum(a[i] for a in agg.values()) for i in range(5)]
print(f"per unit: shared wavefronts {tot[0]/units:.1f} (excessive {tot[1]/units:.1f}), global sectors {tot[2]/units:.1f}, local sectors {tot[3]/units:.1f}, warp instructions {tot[4]/units:.1f}")
for key, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][3] / 4))[:top]:
    print(f"{str(key):30s} shared wf/unit {a[0]/units:7.1f} (excess {a[1]/units:6.1f})  global sect {a[2]/units:6.1f} local sect {a[3]/units:6.1f} instr {a[4]/units:6.1f}")
