#!/bin/bash
# row-major ROIPool gather: parity tests + the bench line's parts
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_pcl.py -m gpu -x -q -k "roipool or pcl" 2>&1 | tail -8 > $O/r2_gpu_tests_16.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_16.json 2> $O/r2_bench_16.err
tail -3 $O/r2_gpu_tests_16.log; tail -3 $O/r2_bench_16.err; python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_16.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['parts'])"
