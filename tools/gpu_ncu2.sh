#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for NL in 1 5; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wave_" -s 4 -c 2 -o $OUT/wave_nl${NL} -f python -m tests.probe_xlong 256 $NL > $OUT/wave_nl${NL}.log 2>&1; echo "NL=$NL rc=$?"
done
ls -la $OUT/*.ncu-rep
