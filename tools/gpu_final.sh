#!/bin/bash
# Round-end style call: GPU tests, smoke, both bench arms with the default flags, per-config benches, ncu launch list.
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.txt
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err; echo "ref rc=$?"
for C in taobao amazon; do timeout 300 python bench.py --config $C --no-cpu-baseline > $OUT/${TAG}_bench_$C.json 2>> $OUT/${TAG}_bench.err; done
python - <<PY
import json
for f in ["bench", "bench_reference_arm", "bench_taobao", "bench_amazon"]:
    try:
        d = json.loads(open("$OUT/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms %.4f e2e %s cpu %s" % (d["value"], d.get("ms_per_step", 0), d.get("e2e", {}).get("value"), d.get("cpu_baseline")))
        if f == "bench":
            for k, v in d["kernels"].items(): print("  %-12s %.4f ms %s" % (k, v["ms_per_step"], ("%.0f GB/s" % v["GBps"]) if "GBps" in v else ""))
            print("  roofline", {k: v for k, v in d["roofline"].items() if k not in ("note",)}); print("  clocks", d["clocks"])
    except Exception as e:
        print(f, "parse failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
