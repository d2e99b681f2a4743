#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "tc_ or fp32_tc" 2>&1 | tail -8 > $O/r2_kernel_tests_8.log
for halo in 0 1; do
  DRN_TC_HALO=$halo timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers8_r50_halo${halo}.txt 2>&1
done
DRN_TC_HALO=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -o $O/r2_conv_layers_v3_base python tools/r2_ncu_conv.py > $O/r2_ncu_conv_v3_base.log 2>&1
tail -3 $O/r2_kernel_tests_8.log
