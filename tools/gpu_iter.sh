#!/bin/bash
# short iteration call: GPU parity tests, then the XLong step probe (per-family times) with the wavefront cycle counters
TAG=${1:-it}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
for NL in ${2:-5}; do
  echo "== NL=$NL"; timeout 120 python -m tests.probe_xlong 256 $NL 2>&1 | tee $OUT/${TAG}_probe_nl$NL.txt | grep -E "step|rec_|wgrad"
  HPMN_WAVE_DEBUG=1 timeout 120 python -m tests.probe_xlong 256 $NL 2>&1 | grep -E "wave_" | sort | tee $OUT/${TAG}_dbg_nl$NL.txt
done
