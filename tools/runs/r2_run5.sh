#!/bin/bash
# GPU call 5 of round 2: warp-uniform TMA/MMA issue (elect.sync) -- correctness, per-layer timings, bench
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -15 > $O/r2_kernel_tests_5.log
for rot in 0 1; do for halo in 0 1; do
  DRN_TC_ROT=$rot DRN_TC_HALO=$halo timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers5_r50_rot${rot}_halo${halo}.txt 2>&1
done; done
DRN_TC_ROT=0 DRN_TC_HALO=0 timeout 300 python tools/layer_bench.py --workload v16_bf16 > $O/r2_layers5_v16_rot0_halo0.txt 2>&1
DRN_TC_ROT=0 DRN_TC_HALO=1 timeout 300 python tools/layer_bench.py --workload v16_bf16 > $O/r2_layers5_v16_rot0_halo1.txt 2>&1
DRN_TC_ROT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_5.json 2> $O/r2_bench_5.err
timeout 1100 python -m pytest tests/test_gpu_bf16_parity.py tests/test_gpu_model_parity.py -x -q 2>&1 | tail -30 > $O/r2_parity_5.log
tail -3 $O/r2_kernel_tests_5.log; tail -3 $O/r2_parity_5.log; tail -c 300 $O/r2_bench_5.err
