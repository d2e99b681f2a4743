#!/usr/bin/env python
"""DRAM traffic of the roofline kernel (the fc6 GEMM) from a per-launch ncu raw CSV of one bench step, written to
profiles/r2_fc6_traffic.json -- the file bench.py reads `roofline.traffic` from (never a literal in bench.py).

    ncu --set full --clock-control none --profile-from-start off --csv --page raw ... python bench.py --profile-step   (GPU box)
    python tools/ncu_traffic.py gpurun_out/r2_step_raw.csv r50_bf16 [commit]                                               (here)
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    path, workload = sys.argv[1], sys.argv[2]
    commit = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    dur = ix["gpu__time_duration.sum"]
    gemms = [r for r in data if "gemm_tc_kernel" in r[ix["Kernel Name"]]]
    top = max(gemms, key=lambda r: float(r[dur].replace(",", "")))  # fc6 is the longest launch of the step by far
    rd = to_bytes(top[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(top[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    out_path = os.path.join(ROOT, "profiles", "r2_fc6_traffic.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    table[workload] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                       "kernel": top[ix["Kernel Name"]][:80], "duration_under_ncu": top[dur] + " " + units[dur],
                       "source": f"ncu --set full capture of one bench step ({os.path.basename(path)}), dram__bytes_read.sum + dram__bytes_write.sum "
                                 f"of the longest gemm_tc_kernel launch (fc6){', tree ' + commit if commit else ''}"}
    json.dump(table, open(out_path, "w"), indent=1)
    print(json.dumps(table[workload]))


if __name__ == "__main__":
    main()
