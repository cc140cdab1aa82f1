#!/bin/bash
# end-of-round measurement pass on one GPU box; everything that comes back fits the 64 MiB of gpurun_out/ (the ncu report is
# turned into its tables on the box and removed)
tag=${1:-r2d}
mkdir -p gpurun_out
bash scripts/gpu_check.sh $tag > gpurun_out/gpu_check_$tag.log 2>&1
bash scripts/ncu_all.sh $tag > gpurun_out/ncu_all_$tag.log 2>&1
python scripts/ncu_table.py gpurun_out/prof_$tag.ncu-rep gpurun_out/kernels_ncu_$tag.md gpurun_out/traffic_$tag.json hexacopter370_flying_arm_3_displacement 4096 400 > gpurun_out/ncu_table_$tag.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/ncu_summary_$tag.txt 2>&1
rm -f gpurun_out/prof_$tag.ncu-rep
bash scripts/ncu_dram.sh $tag > gpurun_out/dram_$tag.txt 2>&1
python scripts/bench_overlays.py > gpurun_out/overlays_$tag.log 2>&1
bash scripts/sanitize_overlays.sh > gpurun_out/sanitizer_overlays_$tag.txt 2>&1
du -sh gpurun_out; ls gpurun_out | head -40
tail -2 gpurun_out/gpu_check_$tag.log | cut -c1-300
