#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -o $O/r2_conv_layers_v4 python tools/r2_ncu_conv.py > $O/r2_ncu_conv_v4.log 2>&1
ls -la $O/r2_conv_layers_v4.ncu-rep
