#!/bin/bash
# quick GPU pass while iterating on a kernel: parity tests, then value + per-kernel split
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 bash scripts/bench_brief.sh 2>&1 | tee gpurun_out/brief_$tag.txt
