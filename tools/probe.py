"""python -m tools.probe [B] [L] : time the XLong-shape step per kernel family (the command the ncu captures under profiles/ run; earlier rounds: tests.probe_xlong)."""
import json
import sys
import time

import numpy as np
import torch

from hpmn_b200.engine import HpmnEngine
from hpmn_b200.layout import HpmnShape

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NL = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sh = HpmnShape(B=B, T=1001, F=2, E=16, H=32, periods=[2] * (NL - 1), L=NL, hops=3, V=3308019, front_pad=23,
               mask_id0=False, last_offset=2)
eng = HpmnEngine(sh, memory_reg=5e-5)
rng = np.random.default_rng(0)
ids = [torch.as_tensor(rng.integers(1, sh.V, size=(B, sh.T, sh.F), dtype=np.int32), device=eng.device) for _ in range(4)]
lab = torch.as_tensor(rng.integers(0, 2, size=B).astype(np.int32), device=eng.device)
for i in range(3):
    eng.forward_backward(ids[i % 4], lab)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    eng.forward_backward(ids[i % 4], lab)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("B=%d step %.3f ms  -> %.0f samples/s  launches/step %d" % (B, ms, B / ms * 1e3, eng.launch_count() // (n + 3)))
eng.profile(True)
for i in range(n):
    eng.forward_backward(ids[i % 4], lab)
prof = eng.profile_read()
eng.profile(False)
tot = sum(v[0] for v in prof.values())
for k, (t, c) in prof.items():
    print("  %-12s %8.3f ms/step  (%4.1f%%)  %d calls" % (k, t / n, 100 * t / max(tot, 1e-9), c // n))
print("  sum %.3f ms/step" % (tot / n), "scalars", eng.scalars.cpu().numpy())
