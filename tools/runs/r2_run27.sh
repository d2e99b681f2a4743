#!/bin/bash
# evidence of the current tree: full GPU suite, default bench line (all legs), ncu per-launch capture of one step, launch list,
# the other BASELINE workloads, eval / train modes, TTA
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2_gpu_tests_27.log
( time timeout 900 python bench.py ) > $O/r2_bench_final.json 2> $O/r2_bench_final.err
timeout 900 ncu --set full --clock-control none --profile-from-start off --csv --page raw --log-file $O/r2_step_raw_final.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_step_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches_final.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_launches_final.log 2>&1
for w in r18_fp32_tc r18_bf16 v16_bf16 r101_coco_bf16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline --library-baseline none > $O/r2_bench_final_$w.json 2> $O/r2_bench_final_$w.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --mode eval --no-cpu-baseline --library-baseline none > $O/r2_bench_final_eval.json 2> $O/r2_bench_final_eval.err
timeout 300 python bench.py --steps 20 --warmup 5 --mode train --no-cpu-baseline --library-baseline none > $O/r2_bench_final_train.json 2> $O/r2_bench_final_train.err
timeout 300 python tools/tta_bench.py > $O/r2_tta_bench_final.json 2> $O/r2_tta_bench_final.err
tail -3 $O/r2_gpu_tests_27.log; tail -4 $O/r2_bench_final.err; wc -l $O/r2_step_raw_final.csv $O/r2_launches_final.csv
for f in $O/r2_bench_final*.json; do python -c "
import json,sys
l=[x for x in open('$f') if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print('$f', d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; done
