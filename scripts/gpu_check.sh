#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (own arm + reference arm), ncu launch list.  Outputs land in gpurun_out/.
# usage: scripts/gpu_check.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
nproc >> gpurun_out/smi_$tag.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 0 --batch 4096 --no-cpu-baseline --no-mpc > gpurun_out/bench_under_ncu_$tag.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/pytest_gpu_$tag.log; cat gpurun_out/smoke_$tag.log | tail -2; cat gpurun_out/bench_$tag.json; cat gpurun_out/bench_ref_$tag.json
