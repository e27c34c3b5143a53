"""torchrun --nproc-per-node N -m tools.exchange_bench : the gradient exchange of the XLong step timed alone, piece by piece
(symmetric-memory barrier, in-switch all-reduce of the dense block, peer-row scatter, whole exchange per mode).  CUDA events on the
launching stream, max over ranks; rank 0 prints."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

from hpmn_b200 import _lib
from hpmn_b200 import dist as hd
from hpmn_b200.data_loader import synthetic_ids
from hpmn_b200.engine import HpmnEngine
from hpmn_b200.layout import HpmnShape


def timed(fn, reps, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, local_rank, world = hd.init_process_group("nccl")
    dev = torch.device("cuda", local_rank)
    sh = HpmnShape(B=256, T=1001, F=2, E=16, H=32, periods=[2, 2, 2, 2], L=5, hops=3, V=3308019, front_pad=23, mask_id0=False, last_offset=2)
    eng = HpmnEngine(sh, device=local_rank, memory_reg=5e-5, symmetric=True)
    ids = torch.as_tensor(synthetic_ids(sh.B, sh.T, sh.F, sh.V, seed=5 + rank), device=dev)
    lab = torch.zeros(sh.B, dtype=torch.int32, device=dev)
    out = {}
    for mode in ("rows", "nvls", "nccl"):
        os.environ["HPMN_EXCHANGE_OVERLAP"] = "0"
        ex = hd.GradExchange(eng, mode=mode).attach()
        eng.set_comm_stream(None)
        eng.forward_backward(ids, lab, loss_batch=sh.B * world)
        torch.cuda.synchronize(dev)
        out["exchange " + mode + " (ran as " + ex.mode + ")"] = timed(ex.exchange, 20, dev)
        if mode == "rows" and ex.mode == "rows":
            out["barrier"] = timed(lambda: ex.hdl_ws.barrier(channel=0), 50, dev)
            out["nvls dense block (0.4 MB)"] = timed(lambda: ex._nvls(min((eng.n_params + 3) & ~3, eng.table_off)), 50, dev)
            ids_off, dx_off, dlast_off = eng.table_grad_sources(sh.B)
            ids_off = eng.ids_sym.data_ptr() - eng.ws_sym.data_ptr()
            cs = eng._cshape(sh.B)
            peers = [(rank + k) % world for k in range(1, world)]
            bases = [int(ex.hdl_ws.buffer_ptrs[p]) for p in peers]
            arr = C.c_void_p * (world - 1)

            def scat():
                st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                _lib.check(eng.lib.hpmn_gather_bwd_multi(eng.ctx, C.byref(cs), world - 1, arr(*[b + ids_off for b in bases]),
                                                         arr(*[b + dx_off for b in bases]), arr(*[b + dlast_off for b in bases]),
                                                         C.c_void_p(eng.dtable.data_ptr()), st), eng.ctx)
            out["peer-row scatter, %d peers x 35.6 MB" % (world - 1)] = timed(scat, 20, dev)

            def scat_local():
                st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                b0 = eng.ws_sym.data_ptr()
                _lib.check(eng.lib.hpmn_gather_bwd_multi(eng.ctx, C.byref(cs), 1, arr(*([b0 + ids_off] * (world - 1))),
                                                         arr(*([b0 + dx_off] * (world - 1))), arr(*([b0 + dlast_off] * (world - 1))),
                                                         C.c_void_p(eng.dtable.data_ptr()), st), eng.ctx)
            out["same scatter on LOCAL rows, 1 source"] = timed(scat_local, 20, dev)
    if rank == 0:
        for k, v in out.items():
            print("N=%d  %-46s %8.1f us" % (world, k, v * 1e3), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
