#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total and share."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("empc::", "")
    v = float(r[iv].replace(",", "")); u = r[iu]
    ms = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    c = tot.setdefault(name, [0, 0.0]); c[0] += 1; c[1] += ms
s = sum(v[1] for v in tot.values())
print("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ms:.2f} | {ms / n:.3f} | {100 * ms / s:.1f}% |")
print(f"| total | {sum(v[0] for v in tot.values())} | {s:.2f} | | |")
