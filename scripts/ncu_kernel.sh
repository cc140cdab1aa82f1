#!/bin/bash
# ncu --set full capture of one kernel family: scripts/ncu_kernel.sh <regex> <tag> [skip] [count]
pat=$1; tag=$2; skip=${3:-2}; cnt=${4:-1}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -f -o gpurun_out/prof_$tag \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-mpc --no-config4 --no-config5 --no-divergent > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$tag.ncu-rep
