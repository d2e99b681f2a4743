#!/bin/bash
# cp.async ROIPool gather (variant 4): parity tests under that variant, warm timings
cd "$(dirname "$0")/../.."
O=gpurun_out
DRN_ROIPOOL_GATHER=4 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "roipool" 2>&1 | tail -4 > $O/r2_gpu_tests_19.log
tail -3 $O/r2_gpu_tests_19.log
for v in 4 0; do
  DRN_ROIPOOL_GATHER=$v timeout 300 python tools/parts_bench.py --only roipool > $O/r2_parts_19_v$v.txt 2> $O/r2_parts_19_v$v.err
  echo "variant $v"; grep -v "^{" $O/r2_parts_19_v$v.txt; tail -2 $O/r2_parts_19_v$v.err
done
