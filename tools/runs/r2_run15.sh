#!/bin/bash
# state check of the tree after the PCL commit: GPU tests, warm per-layer table, the default bench line (all legs)
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_gpu_tests_15.log
timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers_15.txt 2> $O/r2_layers_15.err
( time timeout 900 python bench.py ) > $O/r2_bench_default_15.json 2> $O/r2_bench_default_15.err
tail -3 $O/r2_gpu_tests_15.log; tail -5 $O/r2_bench_default_15.err; tail -40 $O/r2_layers_15.txt
