#!/usr/bin/env python3
"""Print the metrics we track from an .ncu-rep (raw page) + the top stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:80])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print(f"  {k:75s} {r[i]:>18s} {units[i]}")
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warp') and h.endswith('_per_issue_active.ratio') and r[i]]
    if not st:
        st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr) if 'warp_issue_stalled' in h and h.endswith('.ratio') and r[i]]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"  stall {h:73s} {v:10.2f}")
