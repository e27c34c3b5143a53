#!/bin/bash
# diagnostic gpurun call: tests, step probes under switches, ncu --set full captures
TAG=${1:-p}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== probe"; timeout 120 python -m tests.probe_xlong 256 2>&1 | tee $OUT/${TAG}_probe.txt
echo "== probe NO_OVERLAP"; HPMN_NO_OVERLAP=1 timeout 120 python -m tests.probe_xlong 256 2>&1 | head -1
echo "== probe WAVE_DEBUG"; HPMN_WAVE_DEBUG=1 timeout 120 python -m tests.probe_xlong 256 2>&1 | grep -E "wave_|step" | head -20
if [ -n "$2" ]; then
echo "== ncu full: $2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-6} -c ${4:-3} -o $OUT/${TAG}_full -f python -m tests.probe_xlong 256 > $OUT/${TAG}_ncu.log 2>&1; echo "rc=$?"; ls -la $OUT/${TAG}_full.ncu-rep
fi
