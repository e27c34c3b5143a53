"""bench.py -- HPMN fwd+bwd samples/sec on the XLong-shape synthetic workload (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]              # product arm (libhpmn_b200.so on B200)
    python bench.py --impl reference [--gpus N] [--steps K] ...      # reference arm: the TF1-equivalent CPU restatement
    torchrun --nproc-per-node N ... bench.py --gpus N ...            # N > 1: one rank per GPU, NCCL all-reduce per step

One "step" = embedding gather -> 5-layer periodic GRU memory -> covreg + 3-hop attention -> head -> log-loss and
the full backward down to the dense embedding-table gradient (tf.gradients of code/hpmn.py:211; optimizer excluded),
plus, when N > 1, the single flat gradient all-reduce.  Batch 256 per GPU (weak scaling).
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for how each field is produced.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "hpmn_fwd_bwd_samples_per_sec_xlong"
UNIT = "samples/s"

CONFIGS = {
    # BASELINE.json configs[3]: XLong-shape synthetic (code/hpmn.py:643-662 user side)
    "xlong": dict(T=1001, F=2, E=16, H=32, periods=[2, 2, 2, 2], L=5, hops=3, V=3308019, front_pad=23, mask_id0=False,
                  last_offset=2, memory_reg=5e-5, batch=256),
    # configs[2]: Taobao-shape synthetic (T 300 -> 304)
    "taobao": dict(T=300, F=2, E=16, H=32, periods=[2, 2, 2], L=4, hops=3, V=4000000, front_pad=4, mask_id0=True,
                   last_offset=1, memory_reg=1e-5, batch=256),
    # configs[1]: Amazon-shape synthetic
    "amazon": dict(T=100, F=2, E=16, H=18, periods=[2, 2], L=3, hops=3, V=65536, front_pad=0, mask_id0=True,
                   last_offset=1, memory_reg=1e-5, batch=128),
}


def workload_name(cfg_name, cfg, B):
    return "%s-synthetic B=%d/GPU T=%d->%d F=%d E=%d H=%d L=%d periods=%s hops=%d V=%d" % (
        cfg_name, B, cfg["T"], cfg["T"] + cfg["front_pad"], cfg["F"], cfg["E"], cfg["H"], cfg["L"], cfg["periods"],
        cfg["hops"], cfg["V"])


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed regions (and one untimed clock window of the
    same step) run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t0 = time.time()
        while len(self.rows) < 3 and time.time() - t0 < 3.0:      # a slow first nvidia-smi query must not leave the line without clocks
            time.sleep(0.05)
        time.sleep(0.25)
        self.proc.terminate()
        inside = [r for (t, r) in self.rows if any(a <= t <= b for a, b in windows)] or [r for _, r in self.rows]
        sm = [float(r[1]) for r in inside if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(inside)}


# --------------------------------------------------------------------------------------------------
def oracle_inputs(cfg, B, seed_data=1234, seed_params=4321, V_cap=None):
    from oracle import hpmn_oracle as O
    V = min(cfg["V"], V_cap) if V_cap else cfg["V"]
    sh = O.OracleShape(B=B, T=cfg["T"], F=cfg["F"], E=cfg["E"], H=cfg["H"], periods=cfg["periods"], L=cfg["L"],
                       hops=cfg["hops"], V=V, front_pad=cfg["front_pad"], mask_id0=cfg["mask_id0"],
                       last_offset=cfg["last_offset"])
    params, table = O.init_params(sh, seed=seed_params, mode="tf_default")
    ids, labels = O.synthetic_batch(sh, seed=seed_data, ragged=False)
    return sh, params, table, ids, labels


NB = 8   # distinct id batches of the product arm (rotated so that no step re-reads its ids from L2)


_RECORD = []     # the JSON line of this run (printed by main() on the real stdout)


def make_config(cfg_name, cfg, B, world, keep_prob, ids="uniform on [1,V)"):
    """`config` of the JSON line -- identical for the product and the reference arm (same workload, same step)."""
    return {"workload": workload_name(cfg_name, cfg, B), "global_batch": B * world, "ids": ids,
            "parallelism": "dp%d batch-sharded, replicated table, one gradient exchange per step" % world,
            "l2": "inputs larger than L2 (212 MB table + 0.56 GB streamed activations per step, %d rotating id batches)" % NB,
            "keep_prob": keep_prob, "optimizer": "excluded (fwd+bwd metric)"}


def run_reference(args, cfg_name, cfg):
    """Reference arm: the reference's own CPU path on the SAME configuration as the product arm -- every row of the
    B=256 batch, the full V x 16 table, T=1024 recurrence, dropout keep_prob of the train feed, fwd+bwd, optimizer
    excluded.  TF1.4/py2 cannot be installed, so this is the TF1-equivalent restatement
    (oracle/tf1_restatement.py, kind "port") on all host threads.  The table gradient is sparse (indices, rows), which
    is what tf.gradients hands back for tf.nn.embedding_lookup (IndexedSlices) before the clip of code/hpmn.py:212."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import torch
    from oracle import tf1_restatement as R
    B = args.batch or cfg["batch"]
    rows = args.ref_rows or B
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sh, params, table, ids, labels = oracle_inputs(cfg, rows)
    p = R._t(params, torch.float32)
    tb = torch.tensor(table, dtype=torch.float32, requires_grad=True)
    tid, tl = torch.tensor(ids, dtype=torch.int64), torch.tensor(labels, dtype=torch.float32)
    times = []
    for i in range(args.warmup + args.steps):
        for v in p.values():
            v.grad = None
        tb.grad = None
        t0 = time.perf_counter()
        masks = None
        if args.keep_prob < 1.0:
            masks = ((torch.rand(rows, 200) < args.keep_prob).float(), (torch.rand(rows, 80) < args.keep_prob).float())
        out = R.forward_torch(sh, p, tb, tid, tl, cfg["memory_reg"], args.keep_prob, masks, sparse_grad=True)
        out["loss"].backward()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = float(sum(times))
    value = rows * len(times) / total
    sample = "%d of %d rows per step (%s), full V=%d table, full T=%d recurrence, fwd+bwd, fp32 torch-CPU op-per-timestep loop, sparse table gradient" % (
        rows, B, "the whole batch" if rows == B else "a sample", sh.V, cfg["T"] + cfg["front_pad"])
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(cfg_name, cfg, B, max(world, args.gpus), args.keep_prob),
            "note": "TF1-equivalent restatement, not TensorFlow (TF1.4/py2 not installable; DESIGN.md); one CPU process "
                    "(rank 0) runs one rank's batch",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _RECORD.append(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def run_product(args, cfg_name, cfg):
    import torch
    from hpmn_b200 import dist as hd
    from hpmn_b200.data_loader import synthetic_ids
    from hpmn_b200.engine import HpmnEngine
    from hpmn_b200.layout import HpmnShape

    rank, local_rank, world = hd.init_process_group("nccl")
    if world != args.gpus and rank == 0 and world > 1:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)
    B = args.batch or cfg["batch"]
    sh = HpmnShape(B=B, T=cfg["T"], F=cfg["F"], E=cfg["E"], H=cfg["H"], periods=cfg["periods"], L=cfg["L"],
                   hops=cfg["hops"], V=cfg["V"], front_pad=cfg["front_pad"], mask_id0=cfg["mask_id0"],
                   last_offset=cfg["last_offset"])
    eng = HpmnEngine(sh, device=local_rank, memory_reg=cfg["memory_reg"], seed=4321, symmetric=world > 1)   # replicated parameters
    exch = hd.GradExchange(eng, mode=os.environ.get("HPMN_EXCHANGE", "auto")).attach()
    dev = eng.device
    # NB distinct id batches; together with the 212 MB table and the ~0.56 GB of streamed activations the
    # per-step working set is far larger than the 126 MB L2, so no explicit flush is needed
    h_ids = [torch.from_numpy(synthetic_ids(B, sh.T, sh.F, sh.V, seed=1234 + 97 * rank + i, zipf=args.zipf, uid_col=args.uid_col)).pin_memory()
             for i in range(NB)]
    h_lab = [torch.from_numpy(np.random.default_rng(99 + i + rank).integers(0, 2, size=B).astype(np.int32)).pin_memory()
             for i in range(NB)]
    d_ids = [t.to(dev) for t in h_ids]
    d_lab = [t.to(dev) for t in h_lab]
    loss_batch = B * world

    def step_dev(i):
        eng.forward_backward(d_ids[i % NB], d_lab[i % NB], keep_prob=args.keep_prob, seed=i * world + rank, loss_batch=loss_batch)
        if world > 1:
            hd.exchange_grads(eng)                    # the step's gradient exchange (GradExchange: peer rows / in-switch all-reduce)

    def run_host(steps):
        # hpmn_step_host_begin / _end through HpmnEngine.step_host_stream: pinned host ids/labels in, host results out, every
        # step pays its own H2D (staged on the library's copy stream while the previous step computes) and D2H; two steps are
        # in flight, so the host's enqueue time of step i+1 hides behind the device time of step i
        feeds = ((h_ids[i % NB], h_lab[i % NB]) for i in range(steps))
        after = (lambda i: hd.exchange_grads(eng)) if world > 1 else None
        n = 0
        for scal, pred in eng.step_host_stream(feeds, True, args.keep_prob, seed0=1 + rank * 1000003, loss_batch=loss_batch, after_step=after):
            n += 1
        assert n == steps

    def timed(fn, steps, whole=False):
        hd.barrier(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        if whole:
            fn(steps)
        else:
            for i in range(steps):
                fn(i)
        e1.record()
        torch.cuda.synchronize(dev); hd.barrier()
        w1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)    # max over ranks
        return float(ms.item()), (w0, w1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # before the warm-up: nvidia-smi needs ~0.5 s to deliver its first row, the timed regions are ~0.2 s each
    for i in range(args.warmup):
        step_dev(i)
    run_host(min(args.warmup, 3))
    windows = []
    l0 = eng.launch_count()
    ms_dev, w = timed(step_dev, args.steps); windows.append(w)
    launches = eng.launch_count() - l0
    ms_e2e, w = timed(run_host, args.steps, whole=True); windows.append(w)
    # per-kernel-family device time: same step, same inputs, CUDA-event brackets on the launch stream
    eng.profile(True)
    _, w = timed(step_dev, args.steps); windows.append(w)
    prof = eng.profile_read()
    eng.profile(False)
    # clock window: the same step, untimed, for at least 0.6 s, so that the 100 ms sampler sees the load even when the timed
    # regions above were shorter than its period (identical on every rank: the step contains the collective)
    _, w = timed(step_dev, max(args.steps, int(0.6e3 / max(ms_dev / args.steps, 1e-3)))); windows.append(w)
    clocks = sampler.stop(windows) if rank == 0 else None
    scal = eng.scalars.cpu().numpy()
    if rank != 0:
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops", 1590.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    steps_all = sh.steps()
    fl_rec_fwd = B * sum(s * 2 * sh.H * 3 * sh.H for s in steps_all)                 # recurrent half, per launch set
    fl_gru_fwd = B * sh.gru_flops_fwd_per_sample()
    fl_upper_x = B * sum(s * 2 * sh.H * 3 * sh.H for s in steps_all[1:])             # x-half of layers >= 1 (in the wavefront kernels)
    fl_l0_x = fl_gru_fwd - fl_rec_fwd - fl_upper_x                                   # x-half of layer 0 (tcgen05 GEMM)
    fams = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items() if v[1]}
    gather_bytes = sh.gather_bytes() + B * sh.Tpad * sh.D * 4
    scatter_bytes = sh.gather_bytes() + 2 * 4 * sh.E * B * sh.T * sh.F
    if "gather_fwd" in fams:
        fams["gather_fwd"]["GBps"] = gather_bytes / (fams["gather_fwd"]["ms_per_step"] * 1e-3) / 1e9
        fams["gather_fwd"]["frac_of_hbm_peak"] = fams["gather_fwd"]["GBps"] / hbm_peak
    if "scatter_add" in fams:
        fams["scatter_add"]["GBps"] = scatter_bytes / (fams["scatter_add"]["ms_per_step"] * 1e-3) / 1e9
        fams["scatter_add"]["frac_of_hbm_peak"] = fams["scatter_add"]["GBps"] / hbm_peak
    dom = max(fams, key=lambda k: fams[k]["ms_per_step"])
    dom_ms = fams[dom]["ms_per_step"]
    wave = fams.get("rec_fwd", {}).get("launches_per_step", 0) == 1 and sh.L > 1   # fused wavefront kernels in use
    flops_by_family = {"rec_fwd": fl_rec_fwd + (fl_upper_x if wave else 0), "rec_bwd": fl_rec_fwd + (fl_upper_x if wave else 0),
                       "inproj_gemm": fl_l0_x if wave else fl_gru_fwd - fl_rec_fwd,
                       "dx_gemm": fl_l0_x if wave else fl_gru_fwd - fl_rec_fwd, "gru_wgrad": fl_gru_fwd}
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        ncu_traffic = {}
    if dom in flops_by_family:
        achieved = flops_by_family[dom] / (dom_ms * 1e-3) / 1e12
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12       # FFMA lanes x 2 flop x max SM clock
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": achieved / tf_peak, "traffic": ncu_traffic.get(dom) if cfg_name == "xlong" and B == 256 else None,
                "peak_source": peak_src, "algorithmic_flops_per_step": flops_by_family[dom], "kernel_ms_per_step": dom_ms,
                "frac_of_fp32_ffma_peak": achieved / fp32_peak,
                "note": "dense contraction, so reported against the measured tensor peak; the kernel itself is an fp32 FFMA2 "
                        "recurrence (1e-4 parity through 1024 dependent steps rules out single-pass tf32/bf16; M>=64 tcgen05 tiles "
                        "would occupy 2-4 SMs at B=256) and is latency-bound by construction -- DESIGN.md section 4. "
                        "The dense halves run on tcgen05 (inproj_gemm, dx_gemm, gru_wgrad in `kernels`).",
                "traffic_source": ncu_traffic.get("source")}
    else:
        byts = gather_bytes if dom == "gather_fwd" else scatter_bytes
        achieved = byts / (dom_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src, "kernel_ms_per_step": dom_ms}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import tf1_restatement as R
        rows = args.cpu_rows or B
        osh, params, table, ids, labels = oracle_inputs(cfg, rows)
        r = R.time_cpu_baseline(osh, params, table, ids, labels, iters=3, warmup=1, budget_s=40, memory_reg=cfg["memory_reg"],
                                keep_prob=args.keep_prob, sparse_grad=True)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": "%d of %d rows, full V=%d table, full T=%d recurrence, %d timed fwd+bwd passes (median), fp32 torch-CPU "
                         "restatement of the TF1 graph (TensorFlow 1.4 itself is not installable here)" % (rows, B, sh.V, sh.Tpad, r["iters"])}
    value = B * world * args.steps / (ms_dev * 1e-3)
    e2e = B * world * args.steps / (ms_e2e * 1e-3)
    h2d = B * sh.T * sh.F * 4 + B * 4
    d2h = eng.result_block_bytes          # scalars | pred | logit | w_hop0 in one copy (256-byte aligned parts)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": make_config(cfg_name, cfg, B, world, args.keep_prob,
                                  ("Zipf(%g)" % args.zipf if args.zipf else "uniform on [1,V)") + (", column 0 = one uid per sample" if args.uid_col else "")),
            "exchange": {"mode": exch.mode, "why": exch.why},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "path": "hpmn_step_host_begin/_end (C ABI), two steps in flight: pinned host ids/labels -> H2D (double-buffered via hpmn_prefetch_host) -> fwd+bwd -> D2H scalars,pred,logit,weights -> host waits for each step's results"},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "kernels": fams,
            "check": {"logloss": float(scal[0]), "covreg": float(scal[1])}}
    _RECORD.append(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="xlong", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="rows per GPU (default: the config's)")
    ap.add_argument("--keep-prob", type=float, default=0.5, help="dropout keep prob of the train feed (code/hpmn.py:480)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the cpu_baseline sample (default: the whole batch)")
    ap.add_argument("--ref-rows", type=int, default=0, help="rows per step of the reference arm (default: the whole batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--zipf", type=float, default=0.0, help="ids ~ Zipf(a) instead of uniform (popularity skew; SURVEY.md 8d: 1.05)")
    ap.add_argument("--uid-col", action="store_true", help="column 0 constant per sample like the real XLong feed (uid, item)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    # stdout carries exactly ONE line, the JSON record: whatever libraries print on file descriptor 1 while the run is set
    # up (NCCL's version banner under torchrun) goes to stderr; the record is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args, args.config, cfg)
        else:
            run_product(args, args.config, cfg)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
        if _RECORD:
            print(_RECORD[-1], flush=True)


if __name__ == "__main__":
    main()
