#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into the few metrics the roofline needs.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--md]

Prints one block per captured launch: duration, DRAM read/write bytes, L2 (lts) bytes from the SMs,
tensor-pipe active %, L2/DRAM/SM throughput %, registers, achieved warps.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("lts__t_sectors_srcunit_tex.sum", "l2_sectors_from_sm"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d["Kernel Name"]
        print(f"## {name[:110]}")
        for k, label in KEYS:
            if k in d:
                print(f"  {label:26s} {d[k]} {u[k]}")
        try:
            sect = float(d["lts__t_sectors_srcunit_tex.sum"]) * 32
            print(f"  {'l2_bytes_from_sm':26s} {sect / 1e9:.3f} GB")
        except (KeyError, ValueError):
            pass
        print()


if __name__ == "__main__":
    main()
