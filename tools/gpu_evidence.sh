#!/bin/bash
# tools/gpu_evidence.sh : round-2 measurement pass on one B200 (run under gpurun): gather L2-fill experiment, microbench sweep of the
# tensor-core recurrence, ncu captures (launch list of the bench step; full sets of the tcgen05 recurrence kernels and the gather)
mkdir -p gpurun_out
for v in 0 1; do
  HPMN_GATHER_L2_64=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_gather_l2_$v.json 2> gpurun_out/r2_gather_l2_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_gather_l2_$v.json").read().strip().splitlines()[-1])
print("L2_64=$v step %.4f ms  gather %.4f ms %.0f GB/s" % (d["ms_per_step"], d["kernels"]["gather_fwd"]["ms_per_step"], d["kernels"]["gather_fwd"]["GBps"]))
PY
done
for v in 0 1; do
  HPMN_GATHER_L2_64=$v timeout 300 ncu --set full --clock-control none -k regex:gather_fwd -s 3 -c 1 -o gpurun_out/r2_gather_l2_$v python -m tools.probe 256 5 > gpurun_out/r2_gather_ncu_$v.log 2>&1
done
timeout 900 python -m tools.tcrec_bench fwd > gpurun_out/r2_tcrec_bench_fwd.log 2>&1
timeout 900 python -m tools.tcrec_bench fwdbwd > gpurun_out/r2_tcrec_bench_fwdbwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tcrec_ -s 2 -c 4 -o gpurun_out/r2_tcrec_h64_sat python -m tools.tcrec_probe 18944 128 64 2 bwd > gpurun_out/r2_tcrec_ncu_h64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tcrec_ -s 2 -c 4 -o gpurun_out/r2_tcrec_h32_sat python -m tools.tcrec_probe 18944 128 32 2 bwd > gpurun_out/r2_tcrec_ncu_h32.log 2>&1
tail -3 gpurun_out/r2_tcrec_bench_fwd.log | cut -c1-200
