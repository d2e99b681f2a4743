#!/bin/bash
# evidence of the final tree: full GPU suite, default bench line (all legs), ncu per-launch capture of one step, launch list,
# the other BASELINE workloads, eval / train modes, warm per-layer and per-op tables
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2_gpu_tests_final2.log
( time timeout 900 python bench.py ) > $O/r2_bench_final2.json 2> $O/r2_bench_final2.err
timeout 600 ncu --set full --clock-control none --profile-from-start off --csv --page raw --log-file $O/r2_step_raw_final2.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_step_final2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches_final2.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_launches_final2.log 2>&1
for w in r18_fp32_tc r18_bf16 v16_bf16 r101_coco_bf16; do
  timeout 200 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline --library-baseline none > $O/r2_bench_final2_$w.json 2> $O/r2_bench_final2_$w.err
done
timeout 200 python bench.py --steps 20 --warmup 5 --mode eval --no-cpu-baseline --library-baseline none > $O/r2_bench_final2_eval.json 2> $O/r2_bench_final2_eval.err
timeout 200 python bench.py --steps 20 --warmup 5 --mode train --no-cpu-baseline --library-baseline none > $O/r2_bench_final2_train.json 2> $O/r2_bench_final2_train.err
timeout 200 python tools/layer_bench.py > $O/r2_layers_final2.txt 2> $O/r2_layers_final2.err
timeout 200 python tools/parts_bench.py > $O/r2_parts_final2.txt 2> $O/r2_parts_final2.err
tail -3 $O/r2_gpu_tests_final2.log; tail -4 $O/r2_bench_final2.err; wc -l $O/r2_step_raw_final2.csv $O/r2_launches_final2.csv
for f in $O/r2_bench_final2*.json; do python -c "
import json,sys
l=[x for x in open('$f') if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; done
