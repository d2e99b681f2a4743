#!/bin/bash
# ROIPool gather with 8 ROIs per CTA: parity tests + warm timings per variant
cd "$(dirname "$0")/../.."
O=gpurun_out
for v in 2 4; do
DRN_ROIPOOL_GATHER=$v timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "roipool" 2>&1 | tail -2 > $O/r2_gpu_tests_20_v$v.log
tail -1 $O/r2_gpu_tests_20_v$v.log
done
for v in 2 4 0; do
  DRN_ROIPOOL_GATHER=$v timeout 300 python tools/parts_bench.py --only roipool > $O/r2_parts_20_v$v.txt 2> $O/r2_parts_20_v$v.err
  echo "variant $v"; grep -v "^{" $O/r2_parts_20_v$v.txt; tail -2 $O/r2_parts_20_v$v.err
done
