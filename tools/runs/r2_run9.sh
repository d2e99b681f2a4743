#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -8 > $O/r2_kernel_tests_9.log
for halo in 0 1; do
  DRN_TC_HALO=$halo timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers9_r50_halo${halo}.txt 2>&1
done
DRN_TC_HALO=1 timeout 300 python tools/layer_bench.py --workload v16_bf16 > $O/r2_layers9_v16_halo1.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_9.json 2> $O/r2_bench_9.err
DRN_TC_HALO=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_9_halo1.json 2> $O/r2_bench_9_halo1.err
tail -3 $O/r2_kernel_tests_9.log; tail -c 300 $O/r2_bench_9.err
