#!/bin/bash
# per-layer warm table with CTA pairs forced on / off (the 9000-block cross-over rule predates the r2 issue-loop fixes)
cd "$(dirname "$0")/../.."
O=gpurun_out
for cg in 1 2; do
DRN_TC_CTA_GROUP=$cg timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers_22_cg$cg.txt 2> $O/r2_layers_22_cg$cg.err
done
DRN_TC_PDL=0 timeout 300 python tools/parts_bench.py --only first_conv > $O/r2_parts_22_nopdl.txt 2>&1
paste <(cut -c1-60 $O/r2_layers_22_cg1.txt) <(cut -c44-60 $O/r2_layers_22_cg2.txt)
grep -o '"conv_stack_ms": [0-9.]*' $O/r2_parts_22_nopdl.txt
