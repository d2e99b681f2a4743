#!/bin/bash
# predicated-load ROIPool gather + tiled first conv: parity tests, then warm timings per variant
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "roipool or first_conv or c3" 2>&1 | tail -4 > $O/r2_gpu_tests_18.log
tail -3 $O/r2_gpu_tests_18.log
for v in 0 2; do
  DRN_ROIPOOL_GATHER=$v DRN_C3_TILED=$((v/2)) timeout 300 python tools/parts_bench.py --only roipool,first_conv > $O/r2_parts_18_v$v.txt 2> $O/r2_parts_18_v$v.err
  echo "variant $v"; grep -v "^{" $O/r2_parts_18_v$v.txt; tail -2 $O/r2_parts_18_v$v.err
done
