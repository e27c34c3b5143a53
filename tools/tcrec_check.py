"""python -m tools.tcrec_check : error of the tensor-core recurrence per case and per layer (diagnostic twin of
tests/test_gpu_tcrec.py; prints instead of asserting)."""
import os
import sys
import traceback

import numpy as np

os.environ.setdefault("HPMN_TCREC", "1")
from tests.test_gpu_tcrec import BWD_CASES, FWD_CASES, run_memory_bwd, run_memory_fwd  # noqa: E402

def check_bwd(names):
    for n in names:
        try:
            got, ref = run_memory_bwd(BWD_CASES[n])
            worst = max((np.linalg.norm(got[k] - v) / (np.linalg.norm(v) + 1e-12), k) for k, v in ref.items())
            print("%-24s worst rel L2 %.3e (%s)   dx %.3e" % (n, worst[0], worst[1].split("/")[-2] + "/" + worst[1].split("/")[-1] if "/" in worst[1] else worst[1],
                                                              np.linalg.norm(got["dx"] - ref["dx"]) / (np.linalg.norm(ref["dx"]) + 1e-12)), flush=True)
            if worst[0] > 1e-3:
                for k, v in ref.items():
                    print("    %-50s %.3e" % (k, np.linalg.norm(got[k] - v) / (np.linalg.norm(v) + 1e-12)))
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bwd":
        check_bwd(sys.argv[2:] or sorted(BWD_CASES))
        sys.exit(0)
    names = sys.argv[1:] or sorted(FWD_CASES)
    for n in names:
        try:
            mem, ref, launches = run_memory_fwd(FWD_CASES[n])
            err = np.abs(mem - ref) / (np.abs(ref) + 1e-6)
            viol = np.abs(mem - ref) / (1e-4 * np.abs(ref) + 1e-6)          # the test's criterion: <= 1
            per_layer = " ".join("L%d:%.2f/%.1e" % (k, viol[:, k].max(), np.abs(mem - ref)[:, k].max()) for k in range(mem.shape[1]))
            print("%-24s launches %d  max rel %.3e  crit/abs %s  nan=%d  |mem|max %.3f |ref|max %.3f" % (
                n, launches, err.max(), per_layer, int(np.isnan(mem).sum()), np.nanmax(np.abs(mem)), np.abs(ref).max()), flush=True)
            if err.max() > 1e-3:
                b = int(np.argmax(err.reshape(err.shape[0], -1).max(1)))
                print("   worst row", b, "mem", mem[b, 0, :6], "ref", ref[b, 0, :6])
        except Exception:
            traceback.print_exc()
