#!/bin/bash
# ncu --set full per-launch capture of one step of the final tree (per-launch table + fc6 DRAM traffic)
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 400 ncu --set full --clock-control none --profile-from-start off --csv --page raw --log-file $O/r2_step_raw_final3.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_step_final3.log 2>&1
wc -l $O/r2_step_raw_final3.csv; tail -2 $O/r2_ncu_step_final3.log | cut -c1-200
