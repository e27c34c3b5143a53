"""python -m tools.tcrec_bench [fwd|fwdbwd] : K2 (hpmn_memory_fwd) [+ K4 hpmn_memory_bwd] timed alone, tensor-core recurrence
(tcrec.cu) against the warp-per-sample FFMA kernels, over batch / hidden size / sequence length.  CUDA events, 2 warm-up +
n timed calls, inputs rotate over 2 x buffers.  One JSON line per configuration; all lines -> gpurun_out/tcrec_bench.json."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

from hpmn_b200 import _lib
from hpmn_b200.layout import HpmnShape, param_layout

MODE = sys.argv[1] if len(sys.argv) > 1 else "fwd"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def run(B, T, H, tc, L=5, F=2, n=None):
    os.environ["HPMN_TCREC"] = "1" if tc else "0"
    lib = _lib.lib()
    sh = HpmnShape(B=B, T=T, F=F, E=16, H=H, periods=[2] * (L - 1), L=L, hops=1, V=10)
    c = sh.to_c()
    wsb = lib.hpmn_workspace_bytes(C.byref(c), 1)
    if wsb == 0 or wsb > 120e9:
        return None
    ctx = C.c_void_p(); _lib.check(lib.hpmn_create(C.byref(ctx), 0))
    lay, npar = param_layout(sh)
    g = torch.Generator(device="cuda").manual_seed(1)
    params = (torch.rand(npar, device="cuda", generator=g) - 0.5) * 0.3
    xs = [torch.randn(B, sh.Tpad, sh.D, device="cuda", generator=g) * 0.3 for _ in range(2)]
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    mem = torch.zeros(B, L, H, device="cuda")
    dmem = torch.randn(B, L, H, device="cuda", generator=g) * 0.1
    dx = torch.empty(B, sh.Tpad, sh.D, device="cuda") if MODE == "fwdbwd" else None
    grads = torch.zeros(npar, device="cuda")

    def step(i):
        x = xs[i % 2]
        _lib.check(lib.hpmn_memory_fwd(ctx, C.byref(c), x.data_ptr(), params.data_ptr(), mem.data_ptr(), ws.data_ptr(), None), ctx)
        if MODE == "fwdbwd":
            _lib.check(lib.hpmn_memory_bwd(ctx, C.byref(c), x.data_ptr(), params.data_ptr(), dmem.data_ptr(), dx.data_ptr(),
                                           grads.data_ptr(), ws.data_ptr(), None), ctx)
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    n = n or max(3, min(20, int(2e8 / (B * T))))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lib.hpmn_destroy(ctx)
    fl = B * sh.gru_flops_fwd_per_sample() * (3 if MODE == "fwdbwd" else 1)
    rows = B * sum(sh.steps())
    row = dict(mode=MODE, kernels="tcgen05" if tc else "ffma", B=B, T=T, H=H, L=L, ms=ms, samples_per_s=B / ms * 1e3,
               gru_algorithmic_TFLOPs=fl / (ms * 1e-3) / 1e12, frac_of_bf16_peak=fl / (ms * 1e-3) / 1e12 / PEAKS["bf16_tflops"],
               issued_tf32_TFLOPs=3 * fl / (ms * 1e-3) / 1e12 if tc else None,
               state_GBps=rows * 4 * H * 4 * (2 if MODE == "fwdbwd" else 1) / (ms * 1e-3) / 1e9)
    del ws, xs
    torch.cuda.empty_cache()
    return row


if __name__ == "__main__":
    out = []
    cfgs = []
    for H in (32, 64):
        for B in (256, 1024, 4096, 16384, 65536):
            for T in (64, 256, 1024, 4096):
                if B * T > 16384 * 1024 * (2 if MODE == "fwd" else 1):
                    continue
                cfgs.append((B, T, H))
    if len(sys.argv) > 2 and sys.argv[2] == "quick":
        cfgs = [(256, 1024, 32), (4096, 1024, 32), (16384, 1024, 32), (18944, 256, 32), (4096, 1024, 64), (16384, 256, 64), (18944, 256, 64)]
    for (B, T, H) in cfgs:
        for tc in ((True, False) if H <= 32 else (True,)):
            try:
                r = run(B, T, H, tc)
            except Exception as e:  # noqa: BLE001
                r = dict(B=B, T=T, H=H, kernels="tcgen05" if tc else "ffma", error=str(e)[:200])
            if r:
                out.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/tcrec_bench_%s.json" % MODE, "w"), indent=1)
