#!/bin/bash
# tools/gpu_evidence2.sh : round-2 evidence pass on one B200 (under gpurun): microbench sweeps of the tensor-core recurrence, the
# whole step at large batch with and without it, ncu captures (tcgen05 recurrence kernels at a saturating batch; launch list of the
# bench step), SASS mnemonics
mkdir -p gpurun_out
timeout 900 python -m tools.tcrec_bench fwd > gpurun_out/r2_tcrec_bench_fwd.log 2>&1
timeout 900 python -m tools.tcrec_bench fwdbwd > gpurun_out/r2_tcrec_bench_fwdbwd.log 2>&1
for tc in 0 1; do HPMN_TCREC=$tc timeout 600 python -m tools.microbench big > gpurun_out/r2_microbench_step_tc$tc.log 2>&1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tcrec_ -s 2 -c 4 -o gpurun_out/r2_tcrec_h64_sat python -m tools.tcrec_probe 18944 128 64 2 bwd > gpurun_out/r2_tcrec_ncu_h64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tcrec_ -s 2 -c 4 -o gpurun_out/r2_tcrec_h32_sat python -m tools.tcrec_probe 18944 128 32 2 bwd > gpurun_out/r2_tcrec_ncu_h32.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 56 -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
python -m tools.tcrec_stamps 256 512 32 > gpurun_out/r2_stamps_h32.txt 2>&1
python -m tools.tcrec_stamps 256 512 64 > gpurun_out/r2_stamps_h64.txt 2>&1
tail -2 gpurun_out/r2_microbench_step_tc1.log | cut -c1-300
