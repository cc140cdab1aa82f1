#!/bin/bash
# times the bench's per-kernel split for the product library and every variant library in build_var/
mkdir -p gpurun_out
for lib in eagle-mpc_b200/lib/libempc_b200.so build_var/libempc_*.so; do
  echo "== $lib"
  EMPC_LIB=$PWD/$lib timeout 300 bash scripts/bench_brief.sh
done 2>&1 | tee gpurun_out/variants_${1:-v}.txt
