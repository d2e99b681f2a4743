#!/bin/bash
# ncu of the re-tiled first conv (one launch, full set + source counters)
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv3x3_c3 -c 1 -o $O/r2_c3_tile_v2 -f python tools/parts_bench.py --only first_conv --reps 2 > $O/r2_ncu_c3.log 2>&1
tail -3 $O/r2_ncu_c3.log; ls -la $O/r2_c3_tile_v2.ncu-rep
