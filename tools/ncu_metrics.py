#!/usr/bin/env python
"""Print a fixed set of metrics from an .ncu-rep (raw page) for every captured launch."""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers',
 'smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed.avg.per_cycle_active','smsp__inst_executed.sum',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_xu.sum',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
 'l1tex__data_bank_conflicts_pipe_lsu.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum','smsp__inst_executed_op_global_ld.sum']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index('Kernel Name')
print('kernels:', [d[ki][:60] for d in data])
for w in WANT:
    if w in hdr:
        i = hdr.index(w); print(f'{w} [{units[i]}]:', [d[i] for d in data])
if len(sys.argv) > 2:
    for h in hdr:
        if sys.argv[2] in h: i = hdr.index(h); print(h, units[i], [d[i] for d in data])
