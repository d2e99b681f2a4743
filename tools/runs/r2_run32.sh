#!/bin/bash
# first conv v3 (8 channels per thread, conflict-free smem layouts), resident-weight HALO layers issued as one block per tile:
# parity + A/B timings
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bf16_parity.py tests/test_gpu_model_parity.py -m gpu -x -q 2>&1 | tail -4 > $O/r2_gpu_tests_32.log
tail -3 $O/r2_gpu_tests_32.log
for th in 0 4 8; do
  DRN_C3_TH=$th timeout 200 python tools/parts_bench.py --only first_conv > $O/r2_parts_32_th$th.txt 2> $O/r2_parts_32_th$th.err
  echo "TH=$th"; grep -v "^{" $O/r2_parts_32_th$th.txt | head -3; tail -2 $O/r2_parts_32_th$th.err
done
timeout 200 python tools/layer_bench.py > $O/r2_layers_32_new.txt 2> $O/r2_layers_32_new.err
DRN_TC_DEBUG=1024 timeout 200 python tools/layer_bench.py > $O/r2_layers_32_old.txt 2> $O/r2_layers_32_old.err
echo "                                              new    per-tap loop"
paste <(cut -c1-75 $O/r2_layers_32_new.txt) <(cut -c44-52 $O/r2_layers_32_old.txt) | head -8
tail -1 $O/r2_layers_32_new.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --library-baseline none > $O/r2_bench_32.json 2> $O/r2_bench_32.err
python -c "
import json; d=json.loads([l for l in open('$O/r2_bench_32.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parts'])"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv3x3_c3 -c 1 -o $O/r2_c3_tile_v3 -f python tools/parts_bench.py --only first_conv --reps 2 > $O/r2_ncu_c3_v3.log 2>&1
