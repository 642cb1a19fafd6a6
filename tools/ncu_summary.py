#!/usr/bin/env python
"""tools/ncu_summary.py -- condense `ncu -i X.ncu-rep --page raw --csv` into the handful of lines profiles/ keeps.

    python tools/ncu_summary.py gpurun_out/ncu_step_r2a_raw.csv [--launch K] > profiles/ncu_step_r2a_summary.txt
"""
import csv
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    # header row = metric names, next row = units, then one row per launch
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    launches = rows[hdr + 2:]
    which = [int(sys.argv[sys.argv.index("--launch") + 1])] if "--launch" in sys.argv else range(len(launches))
    kcol = names.index("Kernel Name")
    for li in which:
        r = launches[li]
        print(f"Kernel (captured launch {li}): {r[kcol]}")
        for m in KEEP:
            if m in names:
                c = names.index(m)
                print(f"{m:95s} {r[c]:>16s} {units[c]}")
        print()


if __name__ == "__main__":
    main()
