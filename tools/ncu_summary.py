#!/usr/bin/env python
"""Prints the handful of ncu raw-page metrics the roofline discussion needs (from `ncu -i x.ncu-rep --page raw --csv`)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print("%-88s %-14s %s" % (h, units[i], "  ".join(r[i] for r in data)))
