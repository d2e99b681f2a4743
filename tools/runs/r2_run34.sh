#!/bin/bash
# per-tile pipeline trace of CTA 0 for the layers whose tile time is far above their MMA time
cd "$(dirname "$0")/../.."
O=gpurun_out
DRN_TC_DEBUG=2048 timeout 300 python tools/tile_trace.py --only "64,64,3;64,256,1;256,1024,1;512,2048,1;256,256,3;1024,256,1" > $O/r2_tile_trace.txt 2> $O/r2_tile_trace.err
tail -3 $O/r2_tile_trace.err; cat $O/r2_tile_trace.txt
