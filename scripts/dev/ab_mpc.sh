#!/bin/bash
# A/B of the single-instance MPC-step latency (config 3) between CUDA libraries: scripts/dev/ab_mpc.sh lib1 lib2 ...
for rep in 1 2; do
  for lib in "$@"; do
    echo -n "$lib: "; EMPC_LIB=$PWD/$lib python scripts/dev/mpc_latency_only.py 2>&1 | tail -1 | cut -c1-160
  done
done
