#!/bin/bash
# weight tiles fetched before griddepcontrol.wait: parity, conv-stack timing on/off, timeline
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model_parity.py tests/test_gpu_bf16_parity.py -m gpu -x -q 2>&1 | tail -4 > $O/r2_gpu_tests_24.log
tail -3 $O/r2_gpu_tests_24.log
for v in 0 1; do
DRN_TC_PREFETCH_B=$v timeout 300 python tools/parts_bench.py --only first_conv > $O/r2_parts_24_pre$v.txt 2>&1
echo "prefetch_b=$v"; grep -o '"conv_stack_ms": [0-9.]*' $O/r2_parts_24_pre$v.txt
done
DRN_TC_DEBUG=256 timeout 300 python tools/timeline_probe.py > $O/r2_timeline_24.txt 2> $O/r2_timeline_24.err
tail -3 $O/r2_timeline_24.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_24.json 2> $O/r2_bench_24.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_24.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
