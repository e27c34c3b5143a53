"""python -m tools.microbench : BASELINE.json configs[4] -- gather + periodic-GRU microbench.

Sweeps seq_len 64..4096, batch 256..16384 and hidden {16, 32, 64} (H = 64 and B >= 2048 run the memory on the tensor-core
recurrence, csrc/tcrec.cu; H = 128 is not built: DESIGN.md section 1) on synthetic ids, 5 layers period 2, V = 4 M, and reports
per configuration: step time, samples/s, gather GB/s against
the measured HBM peak, and the GRU's algorithmic TFLOP/s (fwd+bwd = 3 x fwd) over the time of the kernels that do GRU work.
Writes gpurun_out/microbench.json; one line per configuration on stdout."""
import json
import os
import sys

import numpy as np
import torch

from hpmn_b200.data_loader import synthetic_ids
from hpmn_b200.engine import HpmnEngine
from hpmn_b200.layout import HpmnShape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0

CONFIGS = ([(256, T, H) for H in (16, 32, 64) for T in (64, 256, 1024, 4096)] + [(B, 1024, 32) for B in (1024, 4096, 16384)] +
           [(B, 1024, 64) for B in (4096, 16384)] + [(4096, 1024, 16)])
if len(sys.argv) > 1 and sys.argv[1] == "big":
    CONFIGS = [(B, 1024, 32) for B in (1024, 4096, 16384)] + [(16384, 256, 32), (65536, 256, 32)]
V = 4000000
out = []
for (B, T, H) in CONFIGS:
    sh = HpmnShape(B=B, T=T, F=2, E=16, H=H, periods=[2, 2, 2, 2], L=5, hops=3, V=V, mask_id0=False)
    eng = HpmnEngine(sh, memory_reg=5e-5)
    ids = [torch.as_tensor(synthetic_ids(B, T, 2, V, seed=i), device=eng.device) for i in range(3)]
    lab = torch.zeros(B, dtype=torch.int32, device=eng.device)
    n = 5 if B * T >= 4096 * 1024 else 20
    for i in range(3):
        eng.forward_backward(ids[i % 3], lab)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        eng.forward_backward(ids[i % 3], lab)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    eng.profile(True)
    for i in range(n):
        eng.forward_backward(ids[i % 3], lab)
    prof = {k: v[0] / n for k, v in eng.profile_read().items()}
    eng.profile(False)
    gru_ms = sum(prof[k] for k in ("inproj_gemm", "rec_fwd", "rec_bwd", "dx_gemm", "gru_wgrad"))
    gru_tflops = 3 * B * sh.gru_flops_fwd_per_sample() / (gru_ms * 1e-3) / 1e12
    gather_gbs = (sh.gather_bytes() + B * sh.Tpad * sh.D * 4) / (prof["gather_fwd"] * 1e-3) / 1e9
    row = dict(B=B, T=T, H=H, ms_per_step=ms, samples_per_s=B / ms * 1e3, gather_GBps=gather_gbs, gather_frac_hbm=gather_gbs / HBM,
               gru_ms=gru_ms, gru_algorithmic_TFLOPs=gru_tflops, rec_fwd_ms=prof["rec_fwd"], rec_bwd_ms=prof["rec_bwd"],
               wgrad_ms=prof["gru_wgrad"], steps_per_layer=sh.steps())
    out.append(row)
    print(json.dumps(row), flush=True)
    eng.close(); del eng, ids
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)
