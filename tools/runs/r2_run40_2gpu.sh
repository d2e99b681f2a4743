#!/bin/bash
# final tree at 2 GPUs: forward+loss and the sharded training step through torchrun
cd "$(dirname "$0")/../.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_final_fwd_2gpu.json 2> $O/r2_final_fwd_2gpu.err
timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 5 --mode train > $O/r2_final_train_2gpu.json 2> $O/r2_final_train_2gpu.err
for f in fwd train; do python -c "
import json
l=[x for x in open('$O/r2_final_${f}_2gpu.json') if x.startswith('{')]
d=json.loads(l[-1]) if l else {}
print('$f', d.get('n_gpus'), d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; tail -2 $O/r2_final_${f}_2gpu.err | cut -c1-200; done
