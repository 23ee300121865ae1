#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the GPU-less box): per-kernel key metrics + per-phase instruction shares."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct',
 'l1tex__t_sector_hit_rate.pct','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_bytes.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed_pipe_tensor.sum','sm__inst_executed_pipe_uniform.sum']
short = lambda w: w.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio','').replace('.avg.pct_of_peak_sustained_active','%').replace('.avg.pct_of_peak_sustained_elapsed','%')
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f'{short(w)} = {r[i]} {units[i]}')
