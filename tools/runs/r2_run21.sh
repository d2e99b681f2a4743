#!/bin/bash
# heads tail kernels (register rows, integer counter reductions): parity, warm per-op timings, bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model_parity.py -m gpu -x -q 2>&1 | tail -4 > $O/r2_gpu_tests_21.log
tail -3 $O/r2_gpu_tests_21.log
timeout 300 python tools/parts_bench.py > $O/r2_parts_21.txt 2> $O/r2_parts_21.err
grep -v "^{" $O/r2_parts_21.txt; tail -2 $O/r2_parts_21.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_21.json 2> $O/r2_bench_21.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_21.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
