"""HPMN_TCR_DEBUG=1 python -m tools.tcrec_stamps B T H : per-step timeline of the tensor-core recurrence (layer 0, CTA 0) from
the clock64 stamps of its hand-offs; mean cycles per interval over steps 64..T-1."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["HPMN_TCR_DEBUG"] = "1"
os.environ.setdefault("HPMN_TCREC", "1")
from hpmn_b200 import _lib  # noqa: E402
from hpmn_b200.layout import HpmnShape, param_layout  # noqa: E402

B, T, H = (int(v) for v in sys.argv[1:4])
lib = _lib.lib()
sh = HpmnShape(B=B, T=T, F=2, E=16, H=H, periods=[], L=1, hops=1, V=10)
c = sh.to_c()
ctx = C.c_void_p(); _lib.check(lib.hpmn_create(C.byref(ctx), 0))
lay, npar = param_layout(sh)
g = torch.Generator(device="cuda").manual_seed(1)
params = (torch.rand(npar, device="cuda", generator=g) - 0.5) * 0.3
x = torch.randn(B, sh.Tpad, sh.D, device="cuda", generator=g) * 0.3
ws = torch.empty(lib.hpmn_workspace_bytes(C.byref(c), 1), dtype=torch.uint8, device="cuda")
mem = torch.zeros(B, 1, H, device="cuda")
for i in range(3):
    _lib.check(lib.hpmn_memory_fwd(ctx, C.byref(c), x.data_ptr(), params.data_ptr(), mem.data_ptr(), ws.data_ptr(), None), ctx)
torch.cuda.synchronize()
n = min(T, 2048)
buf = (C.c_longlong * (n * 16))()
_lib.check(lib.hpmn_debug_tcr_stamps(buf, n * 16))
s = np.frombuffer(buf, dtype=np.int64).reshape(n, 16)
lo = min(64, n // 2)
e, m = s[lo:, :8], s[lo:, 8:]
print("B=%d T=%d H=%d   cycles per step: %.0f" % (B, T, H, np.diff(e[:, 0]).mean()))
names_e = ["wait bar_r", "r part -> arrive rh", "wait bar_u", "u part", "wait bar_2", "c part -> arrive h", "stores"]
for i, nm in enumerate(names_e):
    print("  epilogue  %-22s %7.0f" % (nm, (e[:, i + 1] - e[:, i]).mean()))
names_m = ["wait bar_h", "issue r + commit", "issue u + commit", "issue x(t+1)", "wait bar_rh", "issue c + commit"]
for i, nm in enumerate(names_m):
    print("  mma       %-22s %7.0f" % (nm, (m[:, i + 1] - m[:, i]).mean()))
# cross-thread latencies (same SM clock)
print("  epi arrive h(t-1) -> mma sees it      %7.0f" % (m[1:, 1] - e[:-1, 6]).mean())
print("  mma commit r      -> epi sees it      %7.0f" % (e[:, 1] - m[:, 2]).mean())
print("  epi arrive rh     -> mma sees it      %7.0f" % (m[:, 5] - e[:, 2]).mean())
print("  mma commit c      -> epi sees it      %7.0f" % (e[:, 5] - m[:, 6]).mean())
