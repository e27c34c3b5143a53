#!/bin/bash
# build-variant sweep on the GPU box: rebuilds wave.cu with the given -D sets and probes the XLong step
OUT=gpurun_out; mkdir -p $OUT
for V in "$@"; do
  echo "== variant: $V"
  HPMN_NVCC_EXTRA="$V" python -m hpmn_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 120 python -m tests.probe_xlong 256 5 2>&1 | grep -E "B=256|rec_|wgrad|dx_gemm|scatter"
done
python -m hpmn_b200.build --force > /dev/null 2>&1
