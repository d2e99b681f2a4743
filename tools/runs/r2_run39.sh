#!/bin/bash
# launch 1 of the stage pair on a second stream beside the MIL kernels: full GPU suite, warm timing, default bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2_gpu_tests_39.log
tail -3 $O/r2_gpu_tests_39.log
( time timeout 600 python bench.py ) > $O/r2_bench_39.json 2> $O/r2_bench_39.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_39.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['losses'], d['parts'], d['roofline']['frac'], d['cpu_baseline']['value'], {k:v['value'] for k,v in d['library_baseline'].items()})"
tail -4 $O/r2_bench_39.err
timeout 200 python bench.py --steps 20 --warmup 5 --mode train --no-cpu-baseline --library-baseline none > $O/r2_bench_39_train.json 2> $O/r2_bench_39_train.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_39_train.json') if l.startswith('{')][-1]); print('train', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches_39.csv python bench.py --steps 3 --warmup 3 --profile-step --no-cpu-baseline --library-baseline none > $O/r2_ncu_launches_39.log 2>&1
wc -l $O/r2_launches_39.csv
