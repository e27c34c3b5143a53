#!/bin/bash
# tools/gpu_scale.sh N "mode1 mode2 ..." [steps] : dp_check + bench.py at N GPUs for each exchange mode (run under gpurun --gpus N)
N=$1; MODES=$2; STEPS=${3:-12}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 -m tools.dp_check 2>&1 | grep -E "^mode|Error|error" | head -12
for m in $MODES; do
  HPMN_EXCHANGE=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/r2_n${N}_$m.json 2> gpurun_out/r2_n${N}_$m.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_n${N}_$m.json").read().strip().splitlines()[-1])
    print("N=$N $m value %.0f ms/step %.4f e2e %.0f exchange %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["exchange"]))
except Exception as e:
    print("N=$N $m failed", e); print(open("gpurun_out/r2_n${N}_$m.err").read()[-1200:])
PY
done
# rank-0 kernel timeline of the step incl. the exchange (tools/timeline.py under torchrun)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/timeline.py --out gpurun_out/timeline_n$N.txt 2>/dev/null | tail -16
