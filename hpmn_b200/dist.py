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


class GradExchange:
    """The step's gradient exchange over NVLink / NVSwitch for an HpmnEngine built with symmetric=True (csrc/comm.cu).

    mode "nvls"  -- in-switch all-reduce of the flat [dense | table] gradient: rank r reduces slice r with multimem.ld_reduce
                    and broadcasts it with multimem.st; ~|buffer| in and out per GPU whatever N (a ring moves 2(N-1)/N of it).
    mode "rows"  -- peer-row exchange: every rank points its embedding scatter-add at each PEER's (ids, dX rows, dlast) through
                    the symmetric mapping, so only the touched rows cross NVLink (36 MB per peer at XLong instead of the 212 MB
                    dense table gradient) and they are added while they arrive; the 0.4 MB dense block goes through "nvls".
    mode "nccl"  -- the flat ncclAllReduce (fallback when symmetric memory / multicast is unavailable).
    mode "auto"  -- cost model: rows while (N-1) * (ids + rows) < 1.3 x the dense buffer (measured on 8 x B200, XLong: rows move
                    249 MB per GPU in 0.47 ms, the in-switch all-reduce 212 MB in 0.55 ms, ncclAllReduce 0.60 ms), else nvls.
    Every rank must call exchange() once per step, after its backward call, on the stream that ran it."""

    def __init__(self, eng, mode: str = "auto", group=None):
        self.eng = eng
        self.rank, self.world = rank_world()
        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.mode = "nccl"
        self.why = ""
        self.hdl_grad = self.hdl_ws = None
        if self.world <= 1:
            self.mode = "none"
            return
        if mode == "nccl" or not getattr(eng, "symmetric", False) or eng.device.type != "cuda":
            self.why = "engine not in symmetric memory" if mode != "nccl" else "requested"
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            self.hdl_grad = symm_mem.rendezvous(eng.flat_grad_sym, self.group)
            self.hdl_ws = symm_mem.rendezvous(eng.ws_sym, self.group)
            mc = int(self.hdl_grad.multicast_ptr)
        except Exception as e:  # noqa: BLE001 -- no symmetric memory on this system: keep the NCCL path
            self.why = "symmetric memory unavailable: %s" % (str(e)[:120],)
            self.hdl_grad = self.hdl_ws = None
            return
        self.mc_ptr = mc
        sh = eng.shape
        row_bytes = sh.B * sh.T * sh.F * 4 + sh.B * sh.Tpad * sh.D * 4
        dense_bytes = eng.flat_grad_sym.numel() * 4
        if mode == "auto":
            mode = "rows" if (self.world - 1) * row_bytes < dense_bytes * 1.3 else "nvls"
        if mode == "nvls" and mc == 0:
            mode, self.why = "rows", "no multicast support"
        if mode == "rows":
            try:
                eng.table_grad_sources()
            except Exception as e:  # noqa: BLE001 -- row groups: dX is not one block
                mode, self.why = ("nvls" if mc else "nccl"), str(e)[:120]
        self.mode = mode

    def _nvls(self, n_floats: int, offset_floats: int = 0):
        eng = self.eng
        _lib_check = __import__("hpmn_b200._lib", fromlist=["check"]).check
        import ctypes as C
        st = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
        ctas = int(os.environ.get("HPMN_NVLS_CTAS", "0"))
        _lib_check(eng.lib.hpmn_nvls_allreduce(eng.ctx, C.c_void_p(self.mc_ptr + 4 * offset_floats), n_floats, self.rank, self.world, ctas, st),
                   eng.ctx)

    def attach(self) -> "GradExchange":
        """Overlap: the table part of the exchange runs on its own stream, which the library releases as soon as the local
        scatter-add is done (hpmn_set_comm_stream) -- i.e. beside the GRU weight-gradient reduction that ends the backward call."""
        if self.mode in ("nvls", "rows") and os.environ.get("HPMN_EXCHANGE_OVERLAP", "1") != "0":
            self.comm = torch.cuda.Stream(device=self.eng.device)
            self.eng.set_comm_stream(self.comm)
        self.eng.exchange = self
        return self

    def exchange(self) -> None:
        eng = self.eng
        if self.mode == "none":
            return
        if self.mode == "nccl":
            allreduce_grads(eng)
            return
        import ctypes as C
        from . import _lib
        main = torch.cuda.current_stream(eng.device)
        comm = getattr(self, "comm", None)
        n_dense = min((eng.n_params + 3) & ~3, eng.table_off)
        n_all = eng.flat_grad_sym.numel()
        if self.mode == "rows":
            ids_t, B = eng.last_ids
            ids_off, dx_off, dlast_off = eng.table_grad_sources(B)
            if ids_t is not None:
                ids_off = ids_t.data_ptr() - eng.ws_sym.data_ptr()  # the device-path feed was copied into the symmetric id slot
            cs = eng._cshape(B)
        # ---- table gradient: on the comm stream (released by the library behind the local scatter) when attached, else in line
        with torch.cuda.stream(comm if comm is not None else main):
            self.hdl_ws.barrier(channel=0)                         # every rank's table gradient / dX rows / ids are final
            if self.mode == "nvls":
                self._nvls(n_all - eng.table_off, eng.table_off)
            else:
                st = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
                n = self.world - 1
                peers = [(self.rank + k) % self.world for k in range(1, self.world)]     # staggered start per rank
                bases = [int(self.hdl_ws.buffer_ptrs[p]) for p in peers]
                arr = C.c_void_p * n
                _lib.check(eng.lib.hpmn_gather_bwd_multi(eng.ctx, C.byref(cs), n, arr(*[b + ids_off for b in bases]),
                                                         arr(*[b + dx_off for b in bases]), arr(*[b + dlast_off for b in bases]),
                                                         C.c_void_p(eng.dtable.data_ptr()), st), eng.ctx)
            if comm is not None:
                comm.wait_stream(main)                             # the dense gradients are final at the end of the backward call
            self.hdl_ws.barrier(channel=1)                         # ... on every rank; and every rank is done reading peers' rows
            self._nvls(n_dense)
            self.hdl_ws.barrier(channel=2)                         # every slice of the dense block has been broadcast
        if comm is not None:
            main.wait_stream(comm)


def exchange_grads(eng, ids=None) -> None:
    """The step's gradient exchange for an HpmnEngine: after it every rank holds the sum over ranks of the dense gradients
    and of the embedding-table gradient.  Uses the engine's GradExchange when one is attached (eng.exchange), else the flat
    all-reduce (allreduce_grads)."""
    ex = getattr(eng, "exchange", None)
    if ex is not None:
        ex.exchange()
    else:
        allreduce_grads(eng)


def allreduce_scalars(scalars: torch.Tensor) -> torch.Tensor:
    """logloss (already divided by the global batch), covreg and loss are sums over ranks."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(scalars, op=dist.ReduceOp.SUM)
    return scalars


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
