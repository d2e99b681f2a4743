#!/bin/bash
# ROIPool gather variants (DRN_ROIPOOL_GATHER = 0 bin-major / 2 / 3 row-major) + warm per-op table of the non-GEMM kernels
cd "$(dirname "$0")/../.."
O=gpurun_out
for v in 0 2 3; do
  DRN_ROIPOOL_GATHER=$v timeout 300 python tools/parts_bench.py --only roipool > $O/r2_parts_17_v$v.txt 2> $O/r2_parts_17_v$v.err
  echo "variant $v"; grep -v "^{" $O/r2_parts_17_v$v.txt; tail -2 $O/r2_parts_17_v$v.err
done
timeout 300 python tools/parts_bench.py > $O/r2_parts_17.txt 2> $O/r2_parts_17.err
grep -v "^{" $O/r2_parts_17.txt; tail -2 $O/r2_parts_17.err
