#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --set full` report:  ncu -i X.ncu-rep --page raw --csv > raw.csv;
python tools/ncu_summary.py raw.csv > profiles/<name>_summary.csv  (row 2 = units, as ncu prints them)."""
import csv
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utccp.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__cluster_size",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    keep = [c for c in COLS if c in hdr]
    ix = [hdr.index(c) for c in keep]
    w = csv.writer(sys.stdout)
    w.writerow(keep)
    w.writerow([units[i] for i in ix])
    for r in rows[2:]:
        out = [r[i] for i in ix]
        out[0] = out[0].split("(")[0].replace("d4pg::", "")
        w.writerow(out)


if __name__ == "__main__":
    main(sys.argv[1])
