#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
DRN_TC_HALO=0 DRN_TC_ROT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -o $O/r2_conv_layers_v2_base python tools/r2_ncu_conv.py > $O/r2_ncu_conv_v2_base.log 2>&1
DRN_TC_HALO=1 DRN_TC_ROT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -o $O/r2_conv_layers_v2_halo python tools/r2_ncu_conv.py > $O/r2_ncu_conv_v2_halo.log 2>&1
ls -la $O/*v2*.ncu-rep
