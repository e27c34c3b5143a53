#!/bin/bash
N=$1
mkdir -p gpurun_out
bash tools/gpu_scale.sh $N "auto" 16
HPMN_COMM_SMS=48 HPMN_EXCHANGE=auto timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 16 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reserve48', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/timeline.py --out gpurun_out/timeline_n$N.txt 2>/dev/null | tail -16
