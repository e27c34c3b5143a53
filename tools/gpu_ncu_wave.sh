#!/bin/bash
# one ncu --set full capture (with source counters) of the forward and backward wavefront kernels at XLong shape
TAG=${1:-r1_v9}; NL=${2:-5}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wave_" -s 4 -c 2 -o $OUT/${TAG}_wave_nl${NL} -f python -m tests.probe_xlong 256 $NL > $OUT/${TAG}_wave_nl${NL}.log 2>&1; echo "rc=$?"
ls -la $OUT/${TAG}_wave_nl${NL}.ncu-rep
