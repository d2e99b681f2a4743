#!/bin/bash
# GPU call 3 of round 2: k-block rotation (L2 hot-spot hypothesis) x halo mode
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "tc_" 2>&1 | tail -15 > $O/r2_tc_tests_3.log
for w in r50_bf16 v16_bf16; do
  DRN_TC_ROT=1 DRN_TC_HALO=0 timeout 300 python tools/layer_bench.py --workload $w > $O/r2_layers_${w}_rot1_halo0.txt 2>&1
  DRN_TC_ROT=1 DRN_TC_HALO=1 timeout 300 python tools/layer_bench.py --workload $w > $O/r2_layers_${w}_rot1_halo1.txt 2>&1
done
timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers_r50_bf16_default.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_3.json 2> $O/r2_bench_3.err
timeout 1100 python -m pytest tests/test_gpu_bf16_parity.py -x -q 2>&1 | tail -40 > $O/r2_parity_3.log
tail -3 $O/r2_tc_tests_3.log; tail -3 $O/r2_parity_3.log; tail -c 300 $O/r2_bench_3.err
