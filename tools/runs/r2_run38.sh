#!/bin/bash
# all refinement stages in two launches (drn_oicr_stages_fwd): bit-identity, model parity, backward, warm timing, bench A/B
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model_parity.py tests/test_gpu_backward.py -m gpu -x -q -k "fused_tail or oicr or golden or graph or sgd or iter_size or weighted" 2>&1 | tail -5 > $O/r2_gpu_tests_38.log
tail -4 $O/r2_gpu_tests_38.log
timeout 200 python tools/parts_bench.py --only wsddn_mil_pgt,oicr_stage_fused,oicr_stages > $O/r2_parts_38.txt 2> $O/r2_parts_38.err
grep -v "^{" $O/r2_parts_38.txt; tail -2 $O/r2_parts_38.err
DRN_B200_STAGE_PARALLEL=0 timeout 200 python tools/parts_bench.py --only wsddn_mil_pgt,oicr_stage_fused,oicr_stages > $O/r2_parts_38_seq.txt 2> $O/r2_parts_38_seq.err
grep -v "^{" $O/r2_parts_38_seq.txt; tail -2 $O/r2_parts_38_seq.err
for v in 1 0 1 0; do
DRN_B200_STAGE_PARALLEL=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_38_$v.json 2> $O/r2_bench_38_$v.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_38_$v.json') if l.startswith('{')][-1]); print('stage_parallel=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['losses'])"
done
