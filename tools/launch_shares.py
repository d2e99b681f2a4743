#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel-family time shares.

    python tools/launch_shares.py gpurun_out/launches.csv [--steps N]

ncu serialises launches and runs them cold-cache, so only the SHARES are comparable with the
CUDA-event numbers of bench.py.  --steps divides the totals by the number of steps captured."""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--steps", type=int, default=0, help="steps captured (0: infer from the count of first-conv launches)")
    a = ap.parse_args()
    rows = [l for l in open(a.csv) if not l.startswith("==")]
    rd = csv.DictReader(rows)
    fam = collections.OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |drn::(tc::)?", "", name)
        unit = r["Metric Unit"]
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        grid = r.get("Grid Size", "")
        key = name
        e = fam.setdefault(key, [0, 0.0, 0.0])
        e[0] += 1
        e[1] += us
        e[2] = max(e[2], us)
    steps = a.steps or max(1, next((c for k, (c, _, _) in fam.items() if "conv3x3_c3" in k), 1))
    tot = sum(e[1] for e in fam.values())
    print(f"steps captured: {steps}; summed kernel time per step: {tot / steps:.1f} us; launches per step: {sum(e[0] for e in fam.values()) / steps:.0f}")
    print(f"{'kernel':<60} {'n/step':>7} {'us/step':>9} {'share':>7} {'max us':>8}")
    for k, (c, us, mx) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:<60} {c / steps:7.1f} {us / steps:9.1f} {100 * us / tot:6.1f}% {mx:8.1f}")


if __name__ == "__main__":
    main()
