#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into the per-kernel table kept under profiles/.

  python tools/ncu_summary.py gpurun_out/<tag>_full_raw.csv > profiles/<tag>_full_summary.md
"""
import csv
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

COLS = [
    ("us", "gpu__time_duration.sum", "time"),
    ("grid", "launch__grid_size", None),
    ("blk", "launch__block_size", None),
    ("regs", "launch__registers_per_thread", None),
    ("dram_rd_MB", "dram__bytes_read.sum", "MB"),
    ("dram_wr_MB", "dram__bytes_write.sum", "MB"),
    ("dram_%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("l2_%", "lts__throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("l2_hit_%", "lts__t_sector_hit_rate.pct", None),
    ("sm_%", "sm__throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("tensor_%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", None),
    ("occ_%", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main(path):
    r = csv.reader(open(path))
    hdr = next(r)
    units = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    ik = idx["Kernel Name"]
    print("| # | kernel | " + " | ".join(c[0] for c in COLS) + " | GB/s |")
    print("|---|---|" + "---|" * (len(COLS) + 1))
    for n, row in enumerate(r):
        name = re.sub(r"\(.*", "", row[ik]).replace("void ", "").replace("gk::", "")
        out, t_us, by = [], None, 0.0
        for label, key, kind in COLS:
            i = idx.get(key)
            v = num(row[i]) if i is not None else None
            if v is None:
                out.append("-")
                continue
            u = units[i]
            if kind == "time":
                v *= UNIT.get(u, 1.0)
                t_us = v
            elif kind == "MB":
                v *= UNIT.get(u, 1.0) / 1e6
                by += v
            out.append(f"{v:.1f}" if v < 1000 else f"{v:.0f}")
        gbs = by * 1e6 / (t_us * 1e-6) / 1e9 if t_us else 0.0
        print(f"| {n} | {name} | " + " | ".join(out) + f" | {gbs:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1])
