#!/bin/bash
# fused attention+head kernel: parity, then A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/mid_pytest.txt
cat gpurun_out/mid_pytest.txt
python bench.py --steps 30 --warmup 5 > gpurun_out/mid_bench_fused.json 2> gpurun_out/mid_bench_fused.err
HPMN_NO_FUSE_MID=1 python bench.py --steps 30 --warmup 5 > gpurun_out/mid_bench_separate.json 2> gpurun_out/mid_bench_separate.err
python bench.py --steps 30 --warmup 5 > gpurun_out/mid_bench_fused2.json 2>> gpurun_out/mid_bench_fused.err
for f in fused separate fused2; do python - <<PY
import json
d=json.loads(open("gpurun_out/mid_bench_$f.json").read().strip().splitlines()[-1])
print("$f", d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("families") or d.get("kernels") or "")
PY
done
