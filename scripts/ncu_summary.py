#!/usr/bin/env python3
"""Summarise an Nsight Compute report (.ncu-rep) into the handful of metrics DESIGN.md / bench.py cite.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kcol = hdr.index("Kernel Name")
    print(f"# ncu summary of `{path}`\n")
    for r in data:
        print(f"## launch {r[hdr.index('ID')]}: `{r[kcol][:110]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"| {w} | {r[i]} | {units[i]} |")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(r[hdr.index("dram__bytes_write.sum")]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
            t = float(r[hdr.index("gpu__time_duration.sum")]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[units[hdr.index("gpu__time_duration.sum")]]
            print(f"| **dram traffic (read+write)** | {(rd + wr) / 1e9:.4f} | GB |")
            print(f"| **dram GB/s under ncu** | {(rd + wr) / t / 1e9:.1f} | GB/s |")
        except Exception as e:  # pragma: no cover
            print(f"| traffic | n/a ({e}) | |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
