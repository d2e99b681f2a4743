#!/bin/bash
# 8-GPU lines of the final tree: forward+loss for BASELINE configs[2], [3], [4]; training step with the fused fc6 reduce-scatter
cd "$(dirname "$0")/../.."
O=gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 200 $TR bench.py --gpus $N --steps 30 --warmup 5 > $O/r2_final_fwd_${N}gpu.json 2> $O/r2_final_fwd_${N}gpu.err
timeout 200 $TR bench.py --gpus $N --steps 30 --warmup 5 --workload v16_bf16 > $O/r2_final_fwd_v16_${N}gpu.json 2> $O/r2_final_fwd_v16_${N}gpu.err
timeout 200 $TR bench.py --gpus $N --steps 30 --warmup 5 --workload r101_coco_bf16 > $O/r2_final_fwd_r101_${N}gpu.json 2> $O/r2_final_fwd_r101_${N}gpu.err
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --mode train > $O/r2_final_train_${N}gpu.json 2> $O/r2_final_train_${N}gpu.err
for f in fwd fwd_v16 fwd_r101 train; do python -c "
import json
l=[x for x in open('$O/r2_final_${f}_${N}gpu.json') if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print('$f', d.get('n_gpus'), d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; tail -2 $O/r2_final_${f}_${N}gpu.err | cut -c1-200; done
