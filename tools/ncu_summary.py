#!/usr/bin/env python
"""Condenses `ncu -i X.ncu-rep --page raw --csv` output into the metrics DESIGN.md / bench.py quote (one row per captured
launch), so that the evidence under profiles/ stays small and readable.
    ncu -i gpurun_out/<tag>/lj_full.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/<name>.csv"""
import csv
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    stalls = [i for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in cols] + ["stall reasons (warps per issue-active cycle, > 0.1)"])
    for r in rows[2:]:
        st = sorted(((float(r[i]), hdr[i][len(STALL):-len("_per_issue_active.ratio")]) for i in stalls if r[i]), reverse=True)
        w.writerow([r[i][:110] for i in cols] + ["; ".join(f"{n} {v:.2f}" for v, n in st if v > 0.1)])


if __name__ == "__main__":
    main(sys.argv[1])
