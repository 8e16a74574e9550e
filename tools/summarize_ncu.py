#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export into the handful of metrics DESIGN.md
and bench.py quote (one row per profiled launch)."""
import csv
import sys

KEEP = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_active.avg", "sm_cycles_active"),
    ("sm__cycles_elapsed.max", "sm_cycles_elapsed"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    out.writerow(["kernel"] + ["%s[%s]" % (short, units[idx[m]]) for m, short in KEEP if m in idx])
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        out.writerow([name] + [r[idx[m]] for m, _ in KEEP if m in idx])


if __name__ == "__main__":
    main(sys.argv[1])
