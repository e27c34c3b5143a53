#!/bin/bash
# 2-GPU call: the 2-GPU parity test and both bench arms under torchrun
TAG=${1:-n2}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu or dist" 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_n2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "rc=$?"
tail -c 600 $OUT/${TAG}_bench_n2.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.0f ms %.4f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
