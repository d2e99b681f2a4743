#!/bin/bash
# GPU call 2 of round 2: parity tests of the new files, smoke, halo-mode probes and per-layer timings
cd "$(dirname "$0")/../.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16_parity.py -x -q 2>&1 | tail -40 > $O/r2_parity_2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke_2.log 2>&1
timeout 300 python tools/probe_shifted_desc.py > $O/r2_probe_shifted.log 2>&1
DRN_TC_HALO_BO=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tc_conv3x3" 2>&1 | tail -15 > $O/r2_halo_test_bo0.log
DRN_TC_HALO_BO=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tc_conv3x3" 2>&1 | tail -15 > $O/r2_halo_test_bo1.log
BO=0; grep -q "passed" $O/r2_halo_test_bo1.log && ! grep -q "failed" $O/r2_halo_test_bo1.log && ! grep -q " passed" $O/r2_halo_test_bo0.log && BO=1
grep -q "failed" $O/r2_halo_test_bo0.log && grep -q "passed" $O/r2_halo_test_bo1.log && ! grep -q "failed" $O/r2_halo_test_bo1.log && BO=1
echo "BO=$BO" > $O/r2_halo_bo.txt
for w in r50_bf16 v16_bf16; do
  DRN_TC_HALO=0 timeout 300 python tools/layer_bench.py --workload $w > $O/r2_layers_${w}_halo0.txt 2>&1
  DRN_TC_HALO_BO=$BO DRN_TC_HALO=1 timeout 300 python tools/layer_bench.py --workload $w > $O/r2_layers_${w}_halo1.txt 2>&1
done
for dbg in 32 64 1 2; do
  DRN_TC_HALO=0 DRN_TC_DEBUG=$dbg timeout 300 python tools/layer_bench.py --workload r50_bf16 > $O/r2_layers_r50_dbg${dbg}.txt 2>&1
done
tail -3 $O/r2_parity_2.log; tail -2 $O/r2_smoke_2.log; cat $O/r2_probe_shifted.log; tail -3 $O/r2_halo_test_bo0.log; tail -3 $O/r2_halo_test_bo1.log
