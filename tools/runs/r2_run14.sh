#!/bin/bash
# final-ish evidence run: per-launch ncu table of one step (for the launch shares + fc6 DRAM traffic), multi-scale policy, tests
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $O/r2_gpu_tests_14.log
timeout 600 python tools/multiscale_bench.py --steps 120 --workload r50 > $O/r2_multiscale_r50_v2.json 2> $O/r2_multiscale_r50_v2.err
timeout 900 ncu --set full --clock-control none --profile-from-start off --csv --page raw --log-file $O/r2_step_raw.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_launches.log 2>&1
DRN_B200_CUDA_GRAPH=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_eager.json 2> $O/r2_bench_eager.err
tail -3 $O/r2_gpu_tests_14.log; tail -c 500 $O/r2_multiscale_r50_v2.json; wc -l $O/r2_step_raw.csv $O/r2_launches.csv
