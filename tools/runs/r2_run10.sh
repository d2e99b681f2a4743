#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "tc_" 2>&1 | tail -8 > $O/r2_kernel_tests_10.log
timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers10_r50.txt 2>&1
timeout 300 python tools/layer_bench.py --workload v16_bf16 > $O/r2_layers10_v16.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_10.json 2> $O/r2_bench_10.err
tail -3 $O/r2_kernel_tests_10.log; tail -c 300 $O/r2_bench_10.err; head -4 $O/r2_layers10_r50.txt
