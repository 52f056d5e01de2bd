#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): key raw metrics + the most stalled instructions."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for name, unit, val in zip(h, u, v):
    if name in want:
        print(f"{name} [{unit}] = {val}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
toti = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"-- {tot} stall samples, {toti} warp instructions; top instructions by samples (samples, executed, long_sb, short_sb, sass)")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    print(r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["stall_long_sb"]], r[ix["stall_short_sb"]], r[1].strip()[:80])
