#!/bin/bash
# fc6 tile order (N-fastest vs M-fastest), full step, same box, alternating
cd "$(dirname "$0")/../.."
O=gpurun_out
for rep in 1 2; do for v in 0 1; do
DRN_TC_MMAJOR=$v timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --library-baseline none > $O/r2_bench_26_m${v}_$rep.json 2> $O/r2_bench_26_m${v}_$rep.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_26_m${v}_$rep.json') if l.startswith('{')][-1]); print('m_major=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['losses'])"
tail -2 $O/r2_bench_26_m${v}_$rep.err
done; done
