#!/bin/bash
# prologue trimmed (parallel barrier init, no 64-bit division outside stream-K): trace + parity + layer table + bench
cd "$(dirname "$0")/../.."
O=gpurun_out
DRN_TC_DEBUG=2048 timeout 300 python tools/tile_trace.py --only "64,64,3;64,256,1;1024,256,1" --tiles 3 > $O/r2_tile_trace_36.txt 2> $O/r2_tile_trace_36.err
tail -3 $O/r2_tile_trace_36.err; cat $O/r2_tile_trace_36.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tc_" 2>&1 | tail -3 > $O/r2_gpu_tests_36.log
tail -2 $O/r2_gpu_tests_36.log
timeout 200 python tools/layer_bench.py > $O/r2_layers_36.txt 2> $O/r2_layers_36.err
cut -c1-75 $O/r2_layers_36.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_36.json 2> $O/r2_bench_36.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_36.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
