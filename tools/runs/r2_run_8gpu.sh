#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --mode train > $O/r2_bench_train_${N}gpu_sharded.json 2> $O/r2_bench_train_${N}gpu_sharded.err
DRN_B200_SHARDED=0 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --mode train > $O/r2_bench_train_${N}gpu_allreduce.json 2> $O/r2_bench_train_${N}gpu_allreduce.err
CHECK_TIMED=3 timeout 300 $TR tools/train_sharded_check.py > $O/r2_sharded_check_${N}gpu.log 2>&1
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 > $O/r2_bench_fwd_${N}gpu.json 2> $O/r2_bench_fwd_${N}gpu.err
tail -c 400 $O/r2_bench_train_${N}gpu_sharded.json; echo; tail -c 300 $O/r2_bench_train_${N}gpu_sharded.err; tail -2 $O/r2_sharded_check_${N}gpu.log | cut -c1-300
