#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --out gpurun_out/timeline_new.txt > /dev/null 2>&1
cat gpurun_out/timeline_new.txt
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 > gpurun_out/tl_$name.json 2> gpurun_out/tl_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/tl_$name.json").read().strip().splitlines()[-1])
    f=d.get("families") or d.get("kernels") or {}
    print("$name", round(d["ms_per_step"],4), round(d["value"]), round(d["e2e"]["value"]), "wgrad", round(f.get("gru_wgrad",{}).get("ms_per_step",0),4))
except Exception as e:
    print("$name failed", e)
PY
}
run default A=1
run midpdl HPMN_MID_PDL=1
run stages2 HPMN_WGRAD_STAGES=2
run stages2_c50 HPMN_WGRAD_STAGES=2 HPMN_WGRAD_CARVEOUT=50
run stages3_c75 HPMN_WGRAD_CARVEOUT=75
run zero148 HPMN_ZERO_CTAS=296
run default2 A=1
