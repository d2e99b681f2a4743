#!/bin/bash
# two staging slots for the 8-epilogue-warp 256-wide variant (now default) vs 16 epilogue warps (DRN_TC_EPI16=1)
cd "$(dirname "$0")/../.."
O=gpurun_out
DRN_TC_EPI16=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tc_" 2>&1 | tail -3 > $O/r2_gpu_tests_35_epi16.log
tail -2 $O/r2_gpu_tests_35_epi16.log
timeout 200 python tools/layer_bench.py > $O/r2_layers_35_nbuf2.txt 2> $O/r2_layers_35_nbuf2.err
DRN_TC_EPI16=1 timeout 200 python tools/layer_bench.py > $O/r2_layers_35_epi16.txt 2> $O/r2_layers_35_epi16.err
echo "                                            nbuf2   epi16"
paste <(cut -c1-52 $O/r2_layers_35_nbuf2.txt) <(cut -c44-52 $O/r2_layers_35_epi16.txt)
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_35.json 2> $O/r2_bench_35.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_35.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
DRN_TC_EPI16=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_35_epi16.json 2> $O/r2_bench_35_epi16.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_35_epi16.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
timeout 300 python tools/tta_bench.py > $O/r2_tta_bench_35.json 2> $O/r2_tta_bench_35.err; tail -2 $O/r2_tta_bench_35.json | cut -c1-300
