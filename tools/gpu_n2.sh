#!/bin/bash
# N=2: correctness of the exchange (dp_check), A/B of the SM reserve, rank-0 timeline
mkdir -p gpurun_out
bash tools/gpu_scale.sh 2 "auto" 16
HPMN_COMM_SMS=0 HPMN_EXCHANGE=auto timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 16 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reserve0', d['value'], d['ms_per_step'], d['e2e']['value'])"
HPMN_COMM_SMS=48 HPMN_EXCHANGE=auto timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 16 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reserve48', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/timeline.py --out gpurun_out/timeline_n2.txt 2>/dev/null | tail -28
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu or torchrun or comm_stream" 2>&1 | tail -3
