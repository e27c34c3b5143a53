#!/bin/bash
# One gpurun call: GPU tests, bench (+ A/B switches), ncu launch list.  Everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_session.sh <tag> [quick]
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
RC=$?
tail -5 $OUT/${TAG}_pytest.log
echo "pytest rc=$RC"
if [ $RC -ne 0 ]; then
  echo "== pytest with HPMN_WGRAD_FENCE=1"
  HPMN_WGRAD_FENCE=1 timeout 400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_fence.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest_fence.log
  echo "== pytest with HPMN_NO_TC=1"
  HPMN_NO_TC=1 timeout 400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_notc.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_pytest_notc.log
fi
echo "== bench"
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench.json"))
    print("value %.0f  ms %.4f  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k, v in d["kernels"].items():
        print("  %-12s %.4f ms %s" % (k, v["ms_per_step"], ("%.0f GB/s" % v["GBps"]) if "GBps" in v else ""))
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/${TAG}_bench.err").read()[-2000:])
PY
echo "== bench HPMN_WGRAD_FENCE=1"
HPMN_WGRAD_FENCE=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_fence.json 2>/dev/null
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_fence.json'));print('fence=1: ms %.4f wgrad %.4f'%(d['ms_per_step'],d['kernels']['gru_wgrad']['ms_per_step']))"
if [ "$2" != "quick" ]; then
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; echo "rc=$?"
fi
echo done
