#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_tta.py tests/test_gpu_backward.py -m gpu -x -q 2>&1 | tail -8 > $O/r2_tests_13.log
timeout 600 python tools/multiscale_bench.py --steps 120 --workload r50 > $O/r2_multiscale_r50.json 2> $O/r2_multiscale_r50.err
for w in r18_fp32_tc r18_bf16 v16_bf16 r101_coco_bf16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline --library-baseline none > $O/r2_bench_$w.json 2> $O/r2_bench_$w.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --mode eval --no-cpu-baseline --library-baseline none > $O/r2_bench_eval.json 2> $O/r2_bench_eval.err
timeout 300 python tools/tta_bench.py > $O/r2_tta_bench.json 2> $O/r2_tta_bench.err
tail -3 $O/r2_tests_13.log; tail -c 600 $O/r2_multiscale_r50.json; tail -c 200 $O/r2_multiscale_r50.err
