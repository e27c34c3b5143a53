#!/bin/bash
OUT=gpurun_out/layers.txt; mkdir -p gpurun_out; : > $OUT
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $OUT
for NL in 1 2 3 4 5; do echo "== NL=$NL" >> $OUT; timeout 120 python -m tests.probe_xlong 256 $NL 2>&1 | grep -E "step|rec_|wgrad|dx_gemm|inproj" >> $OUT; done
cat $OUT | grep -E "passed|failed|NL=|rec_|B=256"
