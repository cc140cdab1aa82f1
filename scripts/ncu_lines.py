#!/usr/bin/env python3
"""Attribute ncu per-instruction samples / executed counts to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <lib.sass from nvdisasm --print-line-info -c> <function-substring> [top]"""
import csv, collections, io, re, subprocess, sys
rep, sass, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
# the source page of a multi-kernel report is a concatenation of sections, each starting with a "Kernel Name" row
secs = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
kpat = sys.argv[5] if len(sys.argv) > 5 else None
pick = secs[0]
if kpat:
    pick = [i for i in secs if kpat in allrows[i][1]][0]
end = min([i for i in secs if i > pick] + [len(allrows)])
rows = allrows[pick:end]
hdr = rows[1]
iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
prof = [r for r in rows[2:] if len(r) > iE]
# static listing with line info
lines = []
cur = None; fn = None; inl = None
for l in open(sass):
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m: lines.append((cur, m.group(2)))
print("profile instrs", len(prof), "static instrs", len(lines))
n = min(len(prof), len(lines))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
mism = 0
for k in range(n):
    op_p = re.sub(r'^\s*(@!?U?P\d+\s+)?', '', prof[k][iS]).split()[0]
    op_s = re.sub(r'^(@!?U?P\d+\s+)?', '', lines[k][1]).split()[0]
    if op_p != op_s: mism += 1
    a = agg[lines[k][0]]
    a[0] += int(prof[k][iE]); a[1] += int(prof[k][iN])
    for h, i in stall_cols.items():
        v = int(prof[k][i] or 0)
        if v: a[2][h] += v
print("opcode mismatches", mism)
tot_s = sum(a[1] for a in agg.values()); tot_e = sum(a[0] for a in agg.values())
print(f"total samples {tot_s}, executed {tot_e}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ", ".join(f"{h[6:]}={v}" for h, v in a[2].most_common(3))
    print(f"{str(key):28s} samples {a[1]:8d} ({100*a[1]/tot_s:5.1f}%) exec {a[0]:12d} ({100*a[0]/tot_e:5.1f}%)  {st}")
