#!/bin/bash
# SASS listing + opcode histogram per kernel of the flagship instantiation (Dim<3,6>) from the built library
set -e
out=profiles/sass; mkdir -p $out; tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all $OLDPWD/eagle-mpc_b200/lib/libempc_b200.so > /dev/null && nvdisasm -c solver.sm_100a.cubin > all.sass )
python - "$tmp/all.sass" "$out" <<'PY'
import re, sys, collections
src, out = sys.argv[1:3]
cur = None; bufs = collections.OrderedDict()
for l in open(src):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: cur = m.group(1); bufs[cur] = []; continue
    if cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): bufs[cur].append(l.rstrip())
# free-path instantiations (mangled template flags Lb0E...) + the Box-solver Riccati sweep and its box QP
want = {"node_calc_kernel": "node_calc", "node_cost_kernel": "node_cost", "node_diff_kernel": "node_diff",
        "backward_kernelINS_3DimILi3ELi6EEELb0ELb0E": "backward", "backward_kernelINS_3DimILi3ELi6EEELb1ELb1E": "backward_box",
        "bw_box_qpILi9E": "bw_box_qp",
        "rollout_kernelINS_3DimILi3ELi6EEELi4ELb0E": "rollout_w4", "decide_kernelINS_3DimILi3ELi6EEELb0E": "decide"}
summary = []
for name, lines in bufs.items():
    if "3DimILi3ELi6" not in name and "bw_box_qpILi9E" not in name: continue
    for key, short in want.items():
        if key in name:
            ops = collections.Counter(re.sub(r'^(@!?U?P\d+\s+)?', '', re.sub(r'^\s+/\*[0-9a-f]+\*/\s+', '', x)).split()[0].split('.')[0].rstrip(';') for x in lines)
            open(f"{out}/{short}.sass", "w").write(f"// {name}\n" + "\n".join(lines) + "\n")
            top = ", ".join(f"{k} {v}" for k, v in ops.most_common(12))
            summary.append(f"| `{short}` | {len(lines)} | {ops.get('DMMA',0)} | {ops.get('LDGSTS',0)} | {ops.get('DFMA',0)} | {ops.get('LDL',0)+ops.get('STL',0)} | {top} |")
open(f"{out}/README.md", "w").write("# SASS of the flying_arm_3 (Dim<3,6>) kernels, `nvdisasm -c` of the built library\n\n"
    "| kernel | instructions | DMMA | LDGSTS (cp.async) | DFMA | local ld/st | most frequent opcodes |\n|---|---|---|---|---|---|---|\n" + "\n".join(summary) + "\n")
print("\n".join(summary))
PY
rm -rf $tmp
