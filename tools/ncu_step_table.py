#!/usr/bin/env python
"""Per-launch table from `ncu -i step.ncu-rep --page raw --csv` (exported on the GPU box; the .ncu-rep of a whole
step is too big to bring back): duration, tensor-pipe %, DRAM bytes, L2->SM bytes, warp instructions, SM/mem/L2 %.

    python tools/ncu_step_table.py gpurun_out/r1_step_v7_raw.csv"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("lts__t_sectors_srcunit_tex.sum", "l2->sm"),
        ("sm__inst_executed.sum", "winst"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("launch__grid_size", "grid")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"{'#':>3} {'kernel':<44}" + "".join(f"{n:>9}" for _, n in WANT))
    for i, r in enumerate(data):
        name = r[idx["Kernel Name"]].replace("void ", "").replace("drn::", "")[:44]
        vals = []
        for w, n in WANT:
            c = idx.get(w)
            try:
                f, u = float(r[c].replace(",", "")), units[c]
            except (TypeError, ValueError):
                vals.append("-")
                continue
            if n == "us":
                f = f / 1e3 if u in ("ns", "nsecond") else f * 1e3 if u in ("ms", "msecond") else f
                vals.append(f"{f:.1f}")
            elif n in ("dram_rd", "dram_wr"):
                vals.append(f"{f * {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1) / 1e6:.1f}M")
            elif n == "l2->sm":
                vals.append(f"{f * 32 / 1e6:.0f}M")
            elif n == "winst":
                vals.append(f"{f / 1e6:.2f}M")
            else:
                vals.append(f"{f:.1f}")
        print(f"{i:>3} {name:<44}" + "".join(f"{v:>9}" for v in vals))


if __name__ == "__main__":
    main()
