#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: stall samples and executed instructions by SASS opcode / address space."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
iS = hdr.index('Warp Stall Sampling (All Samples)'); iI = hdr.index('Instructions Executed'); iSrc = hdr.index('Source'); iAS = hdr.index('Address Space')
num = lambda s: int(s) if s.strip().isdigit() else 0
tot = sum(num(r[iS]) for r in data); toti = sum(num(r[iI]) for r in data)
print("kernel", rows[0][1][:70]); print("total samples", tot, "total warp instr", toti, "sass lines", len(data))
byop = collections.Counter(); byopi = collections.Counter()
for r in data:
    toks = r[iSrc].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0]
    sp = r[iAS]
    key = op + ("/" + sp if sp not in ('-', '') else '')
    byop[key] += num(r[iS]); byopi[key] += num(r[iI])
print("--- by opcode: samples%, instr%")
for k, v in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f"{k:22s} {100*v/max(tot,1):6.1f}% {100*byopi[k]/max(toti,1):6.1f}%")
