"""Kernel timeline of the bench step (N = 1): start / end of every kernel and memset of a few steady-state steps on every
stream, from CUPTI through torch.profiler (works for the kernels of libhpmn_b200.so: they run in this process).  Answers
"where do the microseconds between the kernels go" -- what the bracketed per-family times of bench.py cannot show.

    python tools/timeline.py [--steps 6] [--out gpurun_out/timeline.txt]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    from hpmn_b200.data_loader import synthetic_ids
    from hpmn_b200.engine import HpmnEngine
    from hpmn_b200.layout import HpmnShape
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/timeline.txt")
    ap.add_argument("--host", action="store_true", help="the host-buffer path (hpmn_step_host_begin / _end, two steps in flight)")
    args = ap.parse_args()
    cfg = bench.CONFIGS["xlong"]
    B = args.batch or cfg["batch"]
    sh = HpmnShape(B=B, T=cfg["T"], F=cfg["F"], E=cfg["E"], H=cfg["H"], periods=cfg["periods"], L=cfg["L"], hops=cfg["hops"],
                   V=cfg["V"], front_pad=cfg["front_pad"], mask_id0=cfg["mask_id0"], last_offset=cfg["last_offset"])
    from hpmn_b200 import dist as hd
    rank, local_rank, world = hd.init_process_group("nccl")      # torchrun: one rank per GPU, the step includes the exchange
    eng = HpmnEngine(sh, device=local_rank, memory_reg=cfg["memory_reg"], seed=4321, symmetric=world > 1)
    if world > 1:
        hd.GradExchange(eng, mode=os.environ.get("HPMN_EXCHANGE", "auto")).attach()
    dev = eng.device
    NB = 4
    d_ids = [torch.from_numpy(synthetic_ids(B, sh.T, sh.F, sh.V, seed=1234 + 97 * rank + i)).to(dev) for i in range(NB)]
    d_lab = [torch.from_numpy(np.random.default_rng(99 + i).integers(0, 2, size=B).astype(np.int32)).to(dev) for i in range(NB)]

    def step(i):
        eng.forward_backward(d_ids[i % NB], d_lab[i % NB], keep_prob=0.5, seed=i * world + rank, loss_batch=B * world)
        if world > 1:
            hd.exchange_grads(eng)

    h_ids = [t.cpu().pin_memory() for t in d_ids]
    h_lab = [t.cpu().pin_memory() for t in d_lab]

    def host_steps(n):
        feeds = ((h_ids[i % NB], h_lab[i % NB]) for i in range(n))
        for _ in eng.step_host_stream(feeds, True, 0.5, seed0=1, loss_batch=B):
            pass

    for i in range(8):
        step(i)
    host_steps(4)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        if args.host:
            host_steps(args.steps)
        else:
            for i in range(args.steps):
                step(i)
        torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as tmpdir:
        tmp = os.path.join(tmpdir, "trace.json")
        prof.export_chrome_trace(tmp)
        with open(tmp) as f:
            ev = [e for e in json.load(f)["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    # one steady-state step: from the 2nd-to-last gather to the last gather
    gathers = [i for i, e in enumerate(ev) if "gather_fwd" in e["name"]]
    lo, hi = gathers[-2], gathers[-1]
    # the step's first events (memsets) sit in front of the gather
    while lo > 0 and "gather_fwd" not in ev[lo - 1]["name"] and ev[lo - 1]["ts"] > ev[lo]["ts"] - 30:
        lo -= 1
    t0 = ev[lo]["ts"]
    lines = ["%-46s %7s %9s %9s %8s %s" % ("kernel", "stream", "start us", "end us", "dur us", "gap to previous end on this stream")]
    last_end = {}
    for e in ev[lo:hi]:
        st = e["args"].get("stream")
        s, d = e["ts"] - t0, e["dur"]
        gap = s - last_end[st] if st in last_end else float("nan")
        last_end[st] = s + d
        name = e["name"].replace("hpmn::", "").split("(")[0][:46]
        lines.append("%-46s %7s %9.1f %9.1f %8.1f %8.1f" % (name, st, s, s + d, d, gap))
    lines.append("step period (gather to gather): %.1f us" % (ev[hi]["ts"] - ev[gathers[-2]]["ts"]))
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        open(args.out, "w").write("\n".join(lines) + "\n")
        print("\n".join(lines))
    if world > 1:
        hd.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
