"""Data-parallel plumbing: one process per GPU, batch rows sharded across ranks, embedding table and
dense parameters replicated, ONE all-reduce per step over the flat [dense grads | table grad] buffer
(SURVEY.md section 8e).  The reference has no distributed code at all; this is new.

Every rank back-propagates  sum_local logloss_i / B_global + memory_reg * sum_local covreg_b  (the
`loss_batch` knob of the C ABI), so a SUM all-reduce reproduces the single-GPU gradient exactly; clip and
Adam then run identically on every rank.  The helpers are backend-agnostic (nccl on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by it."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str = "nccl") -> Tuple[int, int, int]:
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def rank_world() -> Tuple[int, int]:
    """(rank, world) of the default process group; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the global batch owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_flat(flat_grad: torch.Tensor) -> torch.Tensor:
    """The step's single collective: SUM over ranks of the flat gradient buffer, in place."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


def allreduce_grads(eng) -> None:
    """The step's collectives for an HpmnEngine whose comm stream is set (eng.set_comm_stream): ONE all-reduce of the
    embedding-table gradient (212 MB at XLong) on the comm stream -- the library made that stream wait for the scatter, which
    is queued in front of the GRU weight-gradient reduction, so the transfer runs beside the rest of the backward pass -- and
    one of the 0.4 MB dense block on the caller's stream once the call has finished.  Both are complete for the caller's
    stream on return.  Without a comm stream this is allreduce_flat(eng.flat_grad)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return
    comm = getattr(eng, "comm_stream", None)
    if comm is None:
        dist.all_reduce(eng.flat_grad, op=dist.ReduceOp.SUM)
        return
    with torch.cuda.stream(comm):
        dist.all_reduce(eng.dtable, op=dist.ReduceOp.SUM)
    dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM)
    torch.cuda.current_stream(eng.device).wait_stream(comm)


def exchange_grads(eng, ids=None) -> None:
    """The step's gradient exchange for an HpmnEngine: after it every rank holds the sum over ranks of the dense gradients
    and of the embedding-table gradient.  Default: the flat all-reduce (allreduce_grads)."""
    allreduce_grads(eng)


def allreduce_scalars(scalars: torch.Tensor) -> torch.Tensor:
    """logloss (already divided by the global batch), covreg and loss are sums over ranks."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(scalars, op=dist.ReduceOp.SUM)
    return scalars


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
