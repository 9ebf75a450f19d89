#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into the handful of metrics profiles/ keeps."""
import csv
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_global_loads",
        "sass__inst_executed_global_stores", "launch__occupancy_limit_registers")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(rows[1], vals)))
        print("kernel", d.get("Kernel Name", ("", "?"))[1])
        for k in hdr:
            if k in KEYS or "issue_stalled" in k and "per_issue_active" in k or k.startswith("smsp__pcsamp"):
                u, v = d[k]
                try:
                    if float(v) == 0:
                        continue
                except ValueError:
                    pass
                print(f"  {k:90s} {u:16s} {v}")


if __name__ == "__main__":
    main(sys.argv[1])
