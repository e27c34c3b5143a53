"""python -m tools.tcrec_probe B T H L [bwd] : a few K2 (+K4) calls on the tensor-core recurrence (the ncu target for tcrec.cu)."""
import ctypes as C
import os
import sys

import torch

os.environ.setdefault("HPMN_TCREC", "1")
from hpmn_b200 import _lib  # noqa: E402
from hpmn_b200.layout import HpmnShape, param_layout  # noqa: E402

B, T, H, L = (int(v) for v in sys.argv[1:5])
bwd = len(sys.argv) > 5 and sys.argv[5] == "bwd"
lib = _lib.lib()
sh = HpmnShape(B=B, T=T, F=2, E=16, H=H, periods=[2] * (L - 1), L=L, hops=1, V=10)
c = sh.to_c()
ctx = C.c_void_p(); _lib.check(lib.hpmn_create(C.byref(ctx), 0))
lay, npar = param_layout(sh)
g = torch.Generator(device="cuda").manual_seed(1)
params = (torch.rand(npar, device="cuda", generator=g) - 0.5) * 0.3
x = torch.randn(B, sh.Tpad, sh.D, device="cuda", generator=g) * 0.3
ws = torch.empty(lib.hpmn_workspace_bytes(C.byref(c), 1), dtype=torch.uint8, device="cuda")
mem = torch.zeros(B, L, H, device="cuda")
dmem = torch.randn(B, L, H, device="cuda", generator=g) * 0.1
dx = torch.empty(B, sh.Tpad, sh.D, device="cuda")
grads = torch.zeros(npar, device="cuda")
for i in range(3):
    _lib.check(lib.hpmn_memory_fwd(ctx, C.byref(c), x.data_ptr(), params.data_ptr(), mem.data_ptr(), ws.data_ptr(), None), ctx)
    if bwd:
        _lib.check(lib.hpmn_memory_bwd(ctx, C.byref(c), x.data_ptr(), params.data_ptr(), dmem.data_ptr(), dx.data_ptr(), grads.data_ptr(),
                                       ws.data_ptr(), None), ctx)
torch.cuda.synchronize()
print("ok", float(mem.abs().sum()))
