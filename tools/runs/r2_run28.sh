#!/bin/bash
# re-entry check of the current tree: full GPU suite + default bench line (all legs)
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2_gpu_tests_28.log
( time timeout 900 python bench.py ) > $O/r2_bench_28.json 2> $O/r2_bench_28.err
tail -3 $O/r2_gpu_tests_28.log; tail -4 $O/r2_bench_28.err; tail -c 3000 $O/r2_bench_28.json
