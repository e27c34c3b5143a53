#!/bin/bash
# final single-GPU pass: GPU tests, smoke, bench (product and reference arm), ncu launch list of the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_h_pytest.txt; cat gpurun_out/r2_h_pytest.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/r2_h_bench.json 2> gpurun_out/r2_h_bench.err; head -c 300 gpurun_out/r2_h_bench.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_h_bench_reference_arm.json 2>/dev/null; head -c 250 gpurun_out/r2_h_bench_reference_arm.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 48 -c 52 --csv --log-file gpurun_out/r2_h_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_h_ncu_bench.log 2>&1
grep -c "_kernel" gpurun_out/r2_h_launches.csv || true
