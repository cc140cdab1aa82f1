#!/bin/bash
# one `ncu --set full` capture holding one launch of every kernel family of a solve (third batch-iteration), plus the
# launch list of the first batch-iterations with per-launch durations: scripts/ncu_all.sh <tag>
tag=${1:-all}
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-mpc --no-config4 --no-config5 --no-divergent"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'node_calc|node_cost|node_diff|backward|rollout_kernel|decide' -s 16 -c 8 -f -o gpurun_out/prof_$tag \
  python bench.py $ARGS > gpurun_out/ncu_$tag.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py $ARGS > gpurun_out/bench_under_ncu_$tag.log 2>&1
echo "ncu launch list rc=$?"; ls -la gpurun_out/prof_$tag.ncu-rep gpurun_out/launches_$tag.csv
