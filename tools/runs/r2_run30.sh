#!/bin/bash
# what bounds the wide-N short-K layers: TMA stores queued behind the operand loads?  ablation (no stores), st.global epilogue,
# more staging slots with 3 pipeline stages
cd "$(dirname "$0")/../.."
O=gpurun_out
DRN_TC_EPI_STG=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tc_" 2>&1 | tail -4 > $O/r2_gpu_tests_30_stg.log
tail -3 $O/r2_gpu_tests_30_stg.log
timeout 200 python tools/layer_bench.py > $O/r2_layers_30_base.txt 2> $O/r2_layers_30_base.err
DRN_TC_DEBUG=1 timeout 200 python tools/layer_bench.py > $O/r2_layers_30_nost.txt 2> $O/r2_layers_30_nost.err
DRN_TC_EPI_STG=1 timeout 200 python tools/layer_bench.py > $O/r2_layers_30_stg.txt 2> $O/r2_layers_30_stg.err
DRN_TC_ALT=1 timeout 200 python tools/layer_bench.py > $O/r2_layers_30_alt1.txt 2> $O/r2_layers_30_alt1.err
DRN_TC_ALT=2 timeout 200 python tools/layer_bench.py > $O/r2_layers_30_alt2.txt 2> $O/r2_layers_30_alt2.err
DRN_TC_ALT=3 timeout 200 python tools/layer_bench.py > $O/r2_layers_30_alt3.txt 2> $O/r2_layers_30_alt3.err
echo "                                            base   noST    stg   alt1   alt2   alt3"
paste <(cut -c1-52 $O/r2_layers_30_base.txt) <(cut -c44-52 $O/r2_layers_30_nost.txt) <(cut -c44-52 $O/r2_layers_30_stg.txt) <(cut -c44-52 $O/r2_layers_30_alt1.txt) <(cut -c44-52 $O/r2_layers_30_alt2.txt) <(cut -c44-52 $O/r2_layers_30_alt3.txt)
DRN_TC_EPI_STG=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_30_stg.json 2> $O/r2_bench_30_stg.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_30_stg.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
