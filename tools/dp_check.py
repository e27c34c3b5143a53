"""torchrun --nproc-per-node N -m tools.dp_check : the data-parallel contract on real GPUs.  Every rank runs its row shard, the
gradients are exchanged with each GradExchange mode (nccl / nvls / rows) and compared with the single-GPU gradient of the whole
batch computed on the same GPU.  Exit code 1 on a mismatch; prints one line per mode (rank 0)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from hpmn_b200 import dist as hd
from hpmn_b200.engine import HpmnEngine
from hpmn_b200.layout import HpmnShape
from oracle import hpmn_oracle as O
from tests._parity import oracle_shape


def main():
    rank, local_rank, world = hd.init_process_group("nccl")
    dev = torch.device("cuda", local_rank)
    Bg = 16 * world
    sh = HpmnShape(B=Bg, T=61, F=2, E=16, H=32, periods=[2, 2, 2], L=4, hops=3, V=5000, front_pad=3, mask_id0=False, last_offset=2)
    osh = oracle_shape(sh)
    params, table = O.init_params(osh, mode="stress")
    ids, labels = O.synthetic_batch(osh, ragged=False)
    full = HpmnEngine(sh, device=local_rank, memory_reg=1e-3, table=table, params=params)
    full.forward_backward(torch.as_tensor(ids, device=dev), torch.as_tensor(labels, device=dev))
    torch.cuda.synchronize()
    ref = full.flat_grad.clone()
    lo, hi = hd.shard_range(Bg, rank, world)
    ok = True
    for mode in ("nccl", "nvls", "rows", "auto"):
        eng = HpmnEngine(sh.with_batch(hi - lo), device=local_rank, memory_reg=1e-3, table=table, params=params, symmetric=True)
        ex = hd.GradExchange(eng, mode=mode).attach()
        d_ids, d_lab = torch.as_tensor(ids[lo:hi], device=dev), torch.as_tensor(labels[lo:hi], device=dev)
        errs = []
        for it in range(3):                    # repeated steps: the barriers must also order step i+1 against step i
            eng.forward_backward(d_ids, d_lab, loss_batch=Bg)
            hd.exchange_grads(eng)
            torch.cuda.synchronize()
            errs.append(float((eng.flat_grad - ref).norm() / ref.norm()))
        # host entry point (ids in the workspace slot)
        eng.step_host(ids[lo:hi], labels[lo:hi], with_backward=True, loss_batch=Bg)
        hd.exchange_grads(eng)
        torch.cuda.synchronize()
        errs.append(float((eng.flat_grad - ref).norm() / ref.norm()))
        e = torch.tensor([max(errs)], device=dev)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("mode %-5s -> ran as %-5s %s  max rel err over ranks/steps %.3e" % (mode, ex.mode, ("(" + ex.why + ")") if ex.why else "", float(e)), flush=True)
        ok = ok and float(e) < 1e-5
        dist.barrier()
        eng.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
