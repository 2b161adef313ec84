#!/usr/bin/env python3
"""Key metrics of an .ncu-rep (read here on the CPU box):  python tools/ncu_summary.py file.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ", "launch__occupancy_limit",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum ", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "smsp__average_warps_issue_stalled", "sm__inst_executed_pipe_uniform",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tmem",
        "sm__pipe_tmem", "sm__pipe_tc", "sm__inst_executed_pipe_tc"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, r):
            if any(k in h + " " for k in KEYS) and v not in ("", "0", "0.000000"):
                print("  %-95s %-12s %s" % (h, u, v))


if __name__ == "__main__":
    main()
