#!/bin/bash
# first conv (balanced tiles, batched patch loads), packed epilogue math (FADD2 / FFMA2 / cvt.relu): parity + A/B timings;
# CTA-pair threshold re-sweep on the current kernel
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bf16_parity.py tests/test_gpu_model_parity.py -m gpu -x -q 2>&1 | tail -4 > $O/r2_gpu_tests_29.log
tail -3 $O/r2_gpu_tests_29.log
for th in 0 4 6 8; do
  DRN_C3_TH=$th timeout 200 python tools/parts_bench.py --only first_conv,maxpool2x2 > $O/r2_parts_29_th$th.txt 2> $O/r2_parts_29_th$th.err
  echo "TH=$th"; grep -v "^{" $O/r2_parts_29_th$th.txt | head -12; tail -2 $O/r2_parts_29_th$th.err
done
timeout 200 python tools/layer_bench.py > $O/r2_layers_29_new.txt 2> $O/r2_layers_29_new.err
DRN_TC_DEBUG=512 timeout 200 python tools/layer_bench.py > $O/r2_layers_29_scalar.txt 2> $O/r2_layers_29_scalar.err
DRN_TC_CTA_GROUP=2 timeout 200 python tools/layer_bench.py > $O/r2_layers_29_cg2.txt 2> $O/r2_layers_29_cg2.err
DRN_TC_CTA_GROUP=1 timeout 200 python tools/layer_bench.py > $O/r2_layers_29_cg1.txt 2> $O/r2_layers_29_cg1.err
paste <(cut -c1-75 $O/r2_layers_29_new.txt) <(cut -c44-52 $O/r2_layers_29_scalar.txt) <(cut -c44-52 $O/r2_layers_29_cg2.txt) <(cut -c44-52 $O/r2_layers_29_cg1.txt)
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_29.json 2> $O/r2_bench_29.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_29.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
