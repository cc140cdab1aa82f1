#!/bin/bash
# DRAM bytes and duration of one launch of each kernel family (third batch-iteration): scripts/ncu_dram.sh [tag]
tag=${1:-dram}
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-mpc --no-config4 --no-config5 --no-divergent"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'node_calc|node_cost|node_diff|backward|rollout_kernel|decide' -s 16 -c 8 --csv --log-file gpurun_out/dram_$tag.csv python bench.py $ARGS > /dev/null 2>&1
python - gpurun_out/dram_$tag.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
acc = collections.OrderedDict()
for r in rows[1:]:
    k = r[ik].split('<')[0].replace('void empc::', '') + (' W8' if ', 8>' in r[ik] else '')
    acc.setdefault((r[0], k), {})[r[im]] = (float(r[iv].replace(',', '')), r[iu])
for (i, k), m in acc.items():
    print(k, {n.split('__')[1][:16]: f'{v:.3f} {u}' for n, (v, u) in m.items()})
PY
