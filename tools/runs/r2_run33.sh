#!/bin/bash
# fp32_tc backward (split tensor-core GEMMs for dX and dW): gradient goldens, SGD trajectory, train-mode bench of configs[1]
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -15 > $O/r2_gpu_tests_33.log
tail -15 $O/r2_gpu_tests_33.log
timeout 300 python bench.py --steps 10 --warmup 3 --mode train --workload r18_fp32_tc --no-cpu-baseline --library-baseline none > $O/r2_bench_33_train_r18_fp32_tc.json 2> $O/r2_bench_33_train_r18_fp32_tc.err
tail -3 $O/r2_bench_33_train_r18_fp32_tc.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_33_train_r18_fp32_tc.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
