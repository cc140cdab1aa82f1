#!/bin/bash
# per-kernel device times (ncu, serialised) of the first batch-iterations: scripts/ncu_times.sh <tag> [count]
tag=${1:-t}; cnt=${2:-40}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c $cnt --csv --log-file gpurun_out/times_$tag.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-mpc > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/times_$tag.csv
