#!/bin/bash
# L2-persisting layer weights on/off, full step (same box, alternating)
cd "$(dirname "$0")/../.."
O=gpurun_out
for rep in 1 2; do for v in 0 1; do
DRN_TC_L2_PERSIST=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_25_p${v}_$rep.json 2> $O/r2_bench_25_p${v}_$rep.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_25_p${v}_$rep.json') if l.startswith('{')][-1]); print('persist=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['parts']['conv_stack_ms'], d['roofline']['kernel_ms'], d['clocks'])"
tail -2 $O/r2_bench_25_p${v}_$rep.err
done; done
