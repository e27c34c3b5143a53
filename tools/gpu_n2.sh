#!/bin/bash
# 2-GPU call: GPU tests (incl. the 2-GPU parity test), 1-GPU probe, bench under torchrun with / without the comm overlap
TAG=${1:-n2}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest_n2.txt
timeout 120 python -m tests.probe_xlong 256 5 2>&1 | grep -E "B=256"
for V in "X=0" "HPMN_NO_COMM_OVERLAP=1"; do
  env $V timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $OUT/${TAG}_bench_n2_$V.json 2> $OUT/${TAG}_bench_n2.err; echo "$V rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n2_$V.json").read().strip().splitlines()[-1])
    print("$V N=2 value %.0f ms %.4f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("parse failed", e); print(open("$OUT/${TAG}_bench_n2.err").read()[-1500:])
PY
done
