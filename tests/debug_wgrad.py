"""python -m tests.debug_wgrad : tcgen05 vs FFMA GRU weight-gradient kernels on random buffers (diagnostic)."""
import ctypes as C
import numpy as np
import torch
from hpmn_b200 import _lib
from hpmn_b200.layout import HpmnShape, param_layout

lib = _lib.lib()
ctx = C.c_void_p(); _lib.check(lib.hpmn_create(C.byref(ctx), 0))
for (B, T) in ((1, 32), (2, 64), (16, 256)):
    sh = HpmnShape(B=B, T=T, F=2, E=16, H=32, periods=[2], L=2, hops=1, V=100)
    lay, n = param_layout(sh)
    c = sh.to_c()
    M = B * T
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(M, 32, generator=g).cuda()
    st = torch.rand(M, 128, generator=g).cuda()
    da = torch.randn(M, 96, generator=g).cuda()
    out = []
    for use_tc in (0, 1):
        grads = torch.zeros(n, device="cuda")
        _lib.check(lib.hpmn_debug_wgrad(ctx, C.byref(c), 0, x.data_ptr(), 32, st.data_ptr(), da.data_ptr(), grads.data_ptr(), use_tc, None), ctx)
        torch.cuda.synchronize()
        out.append(grads.cpu().numpy())
    # torch reference
    hprev = torch.zeros(M, 32, device="cuda"); hprev[1:] = st[:-1, :32]
    hprev.view(B, T, 32)[:, 0] = 0
    A = torch.cat([x, hprev], 1).double(); Arh = torch.cat([x, hprev * st[:, 32:64]], 1).double()
    dWg = (A.T @ da[:, :64].double()).cpu().numpy(); dWc = (Arh.T @ da[:, 64:].double()).cpu().numpy()
    off, shp = lay["User/GRU0/rnn/gru_cell/gates/kernel"]
    offc, shpc = lay["User/GRU0/rnn/gru_cell/candidate/kernel"]
    offb, _ = lay["User/GRU0/rnn/gru_cell/gates/bias"]
    for name, o in zip(("ffma", "tc"), out):
        eg = np.abs(o[off:off + 64 * 64].reshape(64, 64) - dWg).max() / np.abs(dWg).max()
        ec = np.abs(o[offc:offc + 64 * 32].reshape(64, 32) - dWc).max() / np.abs(dWc).max()
        eb = np.abs(o[offb:offb + 64] - da[:, :64].double().sum(0).cpu().numpy()).max()
        print("B=%d T=%d %-5s dWg err %.2e  dWc err %.2e  dbg err %.2e   |tc out| max %.3e" % (B, T, name, eg, ec, eb, np.abs(o).max()))
    if B == 1:
        t = out[1][off:off + 64 * 64].reshape(64, 64)
        print(" tc dWg[0:4,0:4]\n", t[:4, :4], "\n ref\n", dWg[:4, :4])
