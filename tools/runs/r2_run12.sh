#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --mode train --no-cpu-baseline --library-baseline none > $O/r2_bench_train_1gpu.json 2> $O/r2_bench_train_1gpu.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_12.json 2> $O/r2_bench_12.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_gpu_tests_12.log
tail -3 $O/r2_gpu_tests_12.log; tail -c 200 $O/r2_bench_train_1gpu.err
