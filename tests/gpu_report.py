"""python -m tests.gpu_report : print CUDA-vs-oracle errors for a sweep of small configurations
(diagnostic twin of tests/test_gpu_parity.py; writes gpurun_out/parity_report.json)."""
import json
import os
import sys
import traceback

from hpmn_b200.layout import HpmnShape
from tests._parity import run_case

CASES = {
    "amazon_like_H32_F3": HpmnShape(B=8, T=20, F=3, E=16, H=32, periods=[2, 5], L=3, hops=3, V=500),
    "amazon_like_H18_F2": HpmnShape(B=8, T=20, F=2, E=16, H=18, periods=[2, 2], L=3, hops=3, V=500),
    "industry_pad": HpmnShape(B=6, T=29, F=2, E=16, H=32, periods=[2, 2, 2], L=4, hops=3, V=300, front_pad=3,
                              mask_id0=False, last_offset=2),
    "taobao_like_F4": HpmnShape(B=5, T=36, F=4, E=16, H=32, periods=[2, 2, 3], L=4, hops=2, V=400),
    "single_layer": HpmnShape(B=3, T=7, F=2, E=8, H=16, periods=[], L=1, hops=1, V=50),
}

if __name__ == "__main__":
    rep = {}
    for name, sh in CASES.items():
        try:
            r = run_case(sh, ragged=sh.mask_id0)
            rep[name] = r
            big = {k: v for k, v in r.items() if isinstance(v, float) and (v > 1e-4 if not k.startswith("grad") and k != "dtable" else v > 1e-3)}
            print(name, "pred %.2e logit %.2e memory %.2e covreg %.2e grad_worst %.2e (%s) dtable %.2e" % (
                r["pred"], r["logit"], r.get("memory", -1), r["covreg"], r["grad_worst"], r["grad_worst_name"], r["dtable"]))
            if big:
                print("   OVER TOLERANCE:", json.dumps(big, indent=1))
        except Exception:
            traceback.print_exc()
            rep[name] = {"error": traceback.format_exc()}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.json", "w") as f:
        json.dump(rep, f, indent=1)
