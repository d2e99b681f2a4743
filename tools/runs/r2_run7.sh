#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "tc_ or fp32_tc" 2>&1 | tail -8 > $O/r2_kernel_tests_7.log
for halo in 0 1; do
  DRN_TC_HALO=$halo timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers7_r50_halo${halo}.txt 2>&1
  DRN_TC_HALO=$halo timeout 300 python tools/layer_bench.py --workload v16_bf16 > $O/r2_layers7_v16_halo${halo}.txt 2>&1
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_7.json 2> $O/r2_bench_7.err
tail -3 $O/r2_kernel_tests_7.log; tail -c 300 $O/r2_bench_7.err
