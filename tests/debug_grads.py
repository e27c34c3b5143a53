"""python -m tests.debug_grads : per-tensor gradient errors for a few small cases (diagnostic)."""
import sys
from hpmn_b200.layout import HpmnShape
from tests._parity import run_case
cases = {
    "F2": HpmnShape(B=8, T=20, F=2, E=16, H=32, periods=[2, 2], L=3, hops=2, V=500),
    "F3": HpmnShape(B=9, T=20, F=3, E=16, H=32, periods=[2, 5], L=3, hops=2, V=500),
    "F2_big": HpmnShape(B=64, T=64, F=2, E=16, H=32, periods=[2, 2], L=3, hops=2, V=500),
}
for n, sh in cases.items():
    r = run_case(sh)
    print(n, "pred %.1e dtable %.1e" % (r["pred"], r["dtable"]))
    for k, v in r.items():
        if k.startswith("grad:") and "GRU" in k:
            print("   %-50s %.3e" % (k[5:], v))
