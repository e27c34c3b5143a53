"""HpmnEngine: owns the device buffers (PyTorch is only the allocator / stream / NCCL plumbing) and calls
the C ABI of libhpmn_b200.so.  One engine per process / GPU.  It plays the role of the TF session the
reference builds in /root/reference/code/hpmn.py:57-62,80-89: `forward` is the eval fetch of
hpmn.py:365-367 / 511-513, `forward_backward` is compute_gradients of hpmn.py:211, `apply_gradients`
is the clip + Adam of hpmn.py:212-214."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .layout import HpmnShape, param_layout


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class HpmnEngine:
    def __init__(self, shape: HpmnShape, device: int = 0, memory_reg: float = 1e-5, l2_reg: float = 0.0,
                 table: Optional[np.ndarray] = None, params: Optional[Dict[str, np.ndarray]] = None,
                 seed: int = 4321, symmetric: bool = False):
        """symmetric: allocate the gradient buffer and the workspace in symmetric memory (torch.distributed._symmetric_memory)
        so that peer ranks can map them -- needed by the NVLink gradient exchanges of hpmn_b200.dist.GradExchange."""
        if not torch.cuda.is_available():
            raise RuntimeError("hpmn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.shape = shape
        self.cshape = shape.to_c()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        ctx = C.c_void_p()
        _lib.check(self.lib.hpmn_create(C.byref(ctx), device))
        self.ctx = ctx
        self.layout, self.n_params = param_layout(shape)
        n_c = self.lib.hpmn_param_count(C.byref(self.cshape))
        if n_c != self.n_params:
            raise RuntimeError("parameter layout mismatch: python %d vs C %d" % (self.n_params, n_c))
        self.memory_reg, self.l2_reg = float(memory_reg), float(l2_reg)
        f32 = dict(dtype=torch.float32, device=self.device)
        # trainables + table live in ONE flat buffer [dense params | table] so that the gradient twin is the
        # single all-reduce message of the data-parallel step and Adam is one sweep
        # The table starts on a 256-byte boundary (zero floats pad the dense block): a 64-byte row that straddles
        # two 64-byte DRAM atoms doubles the HBM traffic of the gather and of the scatter-add (ncu r1: 2.3x).
        self.n_table = shape.V * shape.E
        self.table_off = (self.n_params + 63) & ~63
        self.flat = torch.zeros(self.table_off + self.n_table, **f32)
        self.symmetric = bool(symmetric)
        n_flat = (self.flat.numel() + 3) & ~3
        if self.symmetric:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.flat_grad_sym = symm_mem.empty(n_flat, dtype=torch.float32, device=self.device)
                self.flat_grad_sym.zero_()
                self.flat_grad = self.flat_grad_sym[: self.flat.numel()]
            except Exception as e:  # noqa: BLE001 -- no symmetric memory on this system: plain buffers, NCCL exchange
                import warnings
                warnings.warn("symmetric memory unavailable (%s): gradients will be exchanged with ncclAllReduce" % (str(e)[:120],))
                self.symmetric = False
        if not self.symmetric:
            self.flat_grad = torch.zeros_like(self.flat)
        self.comm_stream = None
        self.params = self.flat[: self.n_params]
        self.table = self.flat[self.table_off:].view(shape.V, shape.E)
        self.grads = self.flat_grad[: self.n_params]
        self.dtable = self.flat_grad[self.table_off:].view(shape.V, shape.E)
        assert self.table.data_ptr() % 256 == 0 and self.dtable.data_ptr() % 256 == 0
        self.adam_m: Optional[torch.Tensor] = None
        self.adam_v: Optional[torch.Tensor] = None
        self.adam_t = 0
        ws_bytes = self.lib.hpmn_workspace_bytes(C.byref(self.cshape), 1)
        if ws_bytes == 0:
            raise ValueError("invalid shape for libhpmn_b200: %r" % (shape,))
        if self.symmetric:
            import torch.distributed._symmetric_memory as symm_mem
            # + one id batch: the device-path feed is copied here so that peers can read it (hpmn_b200.dist.GradExchange)
            self.ids_bytes = shape.B * shape.T * shape.F * 4
            self.ws_sym = symm_mem.empty(((ws_bytes + 255) & ~255) + self.ids_bytes, dtype=torch.uint8, device=self.device)
            self.workspace = self.ws_sym[:ws_bytes]
            self.ids_sym = self.ws_sym[(ws_bytes + 255) & ~255:].view(torch.int32).view(shape.B, shape.T, shape.F)
        if not self.symmetric:
            self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            self.ids_sym = None
        self.last_ids = None          # (tensor or None, rows): ids the last backward call consumed (None: the host entry point's slot)
        B, L, H = shape.B, shape.L, shape.H
        self.scalars = torch.zeros(4, **f32)
        self.pred = torch.zeros(B, **f32)
        self.logit = torch.zeros(B, **f32)
        self.w_hop0 = torch.zeros(B, L, **f32)
        self.memory = torch.zeros(B, L, H, **f32)
        self._out = _lib.hpmn_outputs(_ptr(self.scalars), _ptr(self.pred), _ptr(self.logit), _ptr(self.w_hop0),
                                      _ptr(self.memory))
        # pinned host mirrors for the host-buffer entry point (feed_dict / fetches of hpmn.py:474-482)
        self.h_ids = torch.empty(B, shape.T, shape.F, dtype=torch.int32).pin_memory()
        self.h_labels = torch.empty(B, dtype=torch.int32).pin_memory()
        # host result buffers mirror the library's device staging block (hpmn_output_block): one D2H copy per step; two sets,
        # because step_host_stream keeps two steps in flight
        offs, tot = (C.c_size_t * 4)(), C.c_size_t()
        _lib.check(self.lib.hpmn_output_block(C.byref(self.cshape), offs, C.byref(tot)), self.ctx)

        def result_block():
            blk = torch.zeros(int(tot.value), dtype=torch.uint8).pin_memory()
            f = lambda o, n: blk[int(o): int(o) + 4 * n].view(torch.float32)      # noqa: E731
            sc, pr, lg, w0 = f(offs[0], 4), f(offs[1], B), f(offs[2], B), f(offs[3], B * L).view(B, L)
            return blk, sc, pr, lg, w0, _lib.hpmn_outputs(_ptr(sc), _ptr(pr), _ptr(lg), _ptr(w0), C.c_void_p(0))
        self.result_block_bytes = int(tot.value)
        self._hblk, self.h_scalars, self.h_pred, self.h_logit, self.h_w_hop0, self._out_host = result_block()
        self._hblk2, self.h_scalars2, self.h_pred2, self.h_logit2, self.h_w_hop02, self._out_host2 = result_block()
        self.init_parameters(seed)
        if params is not None:
            self.load_named(params)
        if table is not None:
            self.table.copy_(torch.as_tensor(table, dtype=torch.float32))

    # ------------------------------------------------------------------ parameters
    def view(self, name: str) -> torch.Tensor:
        off, shp = self.layout[name]
        n = int(np.prod(shp))
        return self.params[off: off + n].view(*shp)

    def grad_view(self, name: str) -> torch.Tensor:
        off, shp = self.layout[name]
        n = int(np.prod(shp))
        return self.grads[off: off + n].view(*shp)

    def init_parameters(self, seed: int = 4321):
        """TF1.4 defaults: glorot-uniform kernels (get_variable / dense default), GRU gate bias 1.0
        (util.py:84-86), other biases 0, BN gamma 1 / beta 0."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name, (off, shp) in self.layout.items():
            v = self.view(name)
            if name.endswith("kernel") or name.endswith("/map"):
                lim = float(np.sqrt(6.0 / (shp[0] + shp[1])))
                v.copy_((torch.rand(*shp, generator=g) * 2 - 1) * lim)
            elif name.endswith("gates/bias") or name.endswith("gamma"):
                v.fill_(1.0)
            else:
                v.zero_()
        lim = float(np.sqrt(6.0 / (self.shape.V + self.shape.E)))
        chunk = 1 << 20
        for r0 in range(0, self.shape.V, chunk):
            r1 = min(self.shape.V, r0 + chunk)
            self.table[r0:r1].copy_((torch.rand(r1 - r0, self.shape.E, generator=g) * 2 - 1) * lim)

    def load_named(self, params: Dict[str, np.ndarray]):
        for name, arr in params.items():
            if name in self.layout:
                self.view(name).copy_(torch.as_tensor(np.asarray(arr), dtype=torch.float32))

    def named_parameters(self) -> Dict[str, np.ndarray]:
        return {n: self.view(n).detach().cpu().numpy().copy() for n in self.layout}

    def named_grads(self) -> Dict[str, np.ndarray]:
        return {n: self.grad_view(n).detach().cpu().numpy().copy() for n in self.layout}

    # ------------------------------------------------------------------ calls
    def _hyper(self, keep_prob: float, seed: int, loss_batch: int) -> "_lib.hpmn_hyper":
        return _lib.hpmn_hyper(self.memory_reg, self.l2_reg, float(keep_prob), int(seed) & (2 ** 64 - 1), int(loss_batch))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _cshape(self, B: int):
        """The engine is sized for shape.B rows; any smaller batch reuses the same buffers (the TF graph of the
        reference has a `None` batch dimension, hpmn.py:248-263)."""
        if B == self.shape.B:
            return self.cshape
        if B <= 0 or B > self.shape.B:
            raise ValueError("batch %d outside (0, %d]" % (B, self.shape.B))
        c = self.shape.to_c()
        c.B = B
        return c

    def set_comm_stream(self, stream: Optional["torch.cuda.Stream"]):
        """Multi-GPU overlap: `stream` waits, inside every backward call, for the point where the table gradient is final
        (hpmn_set_comm_stream); hpmn_b200.dist.allreduce_grads() all-reduces `dtable` on it beside the rest of the backward
        pass.  None clears the hook."""
        self.comm_stream = stream
        _lib.check(self.lib.hpmn_set_comm_stream(self.ctx, C.c_void_p(stream.cuda_stream if stream is not None else None)), self.ctx)

    def forward(self, ids: torch.Tensor, labels: torch.Tensor, keep_prob: float = 1.0, seed: int = 0, loss_batch: int = 0):
        """ids [B,T,F] int32 cuda, labels [B] int32 cuda; results land in self.pred / logit / w_hop0 / scalars."""
        hy = self._hyper(keep_prob, seed, loss_batch)
        _lib.check(self.lib.hpmn_forward(self.ctx, C.byref(self._cshape(ids.shape[0])), C.byref(hy), _ptr(ids), _ptr(labels),
                                         _ptr(self.params), _ptr(self.table), C.byref(self._out), _ptr(self.workspace),
                                         self._stream()), self.ctx)

    def forward_backward(self, ids: torch.Tensor, labels: torch.Tensor, keep_prob: float = 1.0, seed: int = 0,
                         loss_batch: int = 0, zero_dtable: bool = True):
        hy = self._hyper(keep_prob, seed, loss_batch)
        if self.ids_sym is not None:          # peers read this rank's ids during the peer-row exchange
            self.ids_sym[: ids.shape[0]].copy_(ids)
            ids = self.ids_sym[: ids.shape[0]]
        self.last_ids = (ids, ids.shape[0])
        _lib.check(self.lib.hpmn_forward_backward(self.ctx, C.byref(self._cshape(ids.shape[0])), C.byref(hy), _ptr(ids), _ptr(labels),
                                                  _ptr(self.params), _ptr(self.table), _ptr(self.grads), _ptr(self.dtable),
                                                  int(zero_dtable), C.byref(self._out), _ptr(self.workspace),
                                                  self._stream()), self.ctx)

    def step_host(self, ids: np.ndarray, labels: np.ndarray, with_backward: bool = True, keep_prob: float = 1.0,
                  seed: int = 0, loss_batch: int = 0, zero_dtable: bool = True):
        """Host buffers in, host results out (pinned staging; H2D + compute + D2H inside; returns when the host results have
        landed -- the gradient buffers are device tensors and complete in stream order on the current stream)."""
        B = int(np.shape(ids)[0])
        self.h_ids.numpy().reshape(-1)[: B * self.shape.T * self.shape.F] = np.asarray(ids, dtype=np.int32).reshape(-1)
        self.h_labels.numpy()[:B] = np.asarray(labels, dtype=np.int32)
        return self.step_host_pinned(with_backward, keep_prob, seed, loss_batch, zero_dtable, B)

    def step_host_pinned(self, with_backward: bool = True, keep_prob: float = 1.0, seed: int = 0, loss_batch: int = 0,
                         zero_dtable: bool = True, B: Optional[int] = None, h_ids: Optional[torch.Tensor] = None,
                         h_labels: Optional[torch.Tensor] = None, prefetch_next=None):
        """Same, with the feed already in pinned host memory: self.h_ids / self.h_labels (first B rows, contiguous)
        or caller-owned pinned int32 tensors.  prefetch_next = (pinned ids, pinned labels) of the NEXT batch: its H2D copy
        is queued on the library's copy stream while this step computes (double-buffered feed)."""
        B = self.shape.B if B is None else B
        h_ids = self.h_ids if h_ids is None else h_ids
        h_labels = self.h_labels if h_labels is None else h_labels
        hy = self._hyper(keep_prob, seed, loss_batch)
        cs = self._cshape(B)
        st = self._stream()
        self.last_ids = (None, B)
        _lib.check(self.lib.hpmn_step_host_begin(self.ctx, C.byref(cs), C.byref(hy), _ptr(h_ids),
                                                 _ptr(h_labels), _ptr(self.params), _ptr(self.table), _ptr(self.grads),
                                                 _ptr(self.dtable), int(zero_dtable), int(with_backward),
                                                 C.byref(self._out_host), _ptr(self.workspace), st), self.ctx)
        if prefetch_next is not None:
            self.prefetch_host(prefetch_next[0], prefetch_next[1], B)
        _lib.check(self.lib.hpmn_step_host_end(self.ctx, C.byref(cs), C.byref(self._out_host), st), self.ctx)
        return self.h_scalars.numpy(), self.h_pred.numpy()[:B]

    def step_host_stream(self, feeds, with_backward: bool = True, keep_prob: float = 1.0, seed0: int = 0, loss_batch: int = 0,
                         after_step=None):
        """Host buffers in, host results out for a SEQUENCE of batches, software-pipelined: `feeds` yields (pinned ids, pinned
        labels); step i+1 is enqueued (its H2D copy double-buffered on the copy stream) before the host waits for the results of
        step i, so the host's launch time never sits between two steps.  Yields (scalars, pred) numpy views of step i -- valid until
        the step after next is enqueued.  after_step(i), if given, is called right after step i is enqueued (gradient exchange,
        optimizer).  Every step still pays its own H2D and D2H inside whatever region the caller times."""
        outs = ((self._out_host, self.h_scalars, self.h_pred), (self._out_host2, self.h_scalars2, self.h_pred2))
        st = self._stream()
        pending = None
        it = iter(feeds)
        nxt = next(it, None)
        if nxt is not None:
            self.prefetch_host(nxt[0], nxt[1], int(nxt[1].shape[0]))
        i = 0
        while nxt is not None:
            cur, nxt = nxt, next(it, None)
            B = int(cur[1].shape[0])
            out, hs, hp = outs[i & 1]
            cs = self._cshape(B)
            hy = self._hyper(keep_prob, seed0 + i, loss_batch)
            self.last_ids = (None, B)
            _lib.check(self.lib.hpmn_step_host_begin(self.ctx, C.byref(cs), C.byref(hy), _ptr(cur[0]), _ptr(cur[1]), _ptr(self.params),
                                                     _ptr(self.table), _ptr(self.grads), _ptr(self.dtable), 1, int(with_backward),
                                                     C.byref(out), _ptr(self.workspace), st), self.ctx)
            if after_step is not None:
                after_step(i)
            if nxt is not None:
                self.prefetch_host(nxt[0], nxt[1], int(nxt[1].shape[0]))
            if pending is not None:
                pcs, pout, phs, php, pB = pending
                _lib.check(self.lib.hpmn_step_host_end(self.ctx, C.byref(pcs), C.byref(pout), st), self.ctx)
                yield phs.numpy(), php.numpy()[:pB]
            pending = (cs, out, hs, hp, B)
            i += 1
        if pending is not None:
            pcs, pout, phs, php, pB = pending
            _lib.check(self.lib.hpmn_step_host_end(self.ctx, C.byref(pcs), C.byref(pout), st), self.ctx)
            yield phs.numpy(), php.numpy()[:pB]

    def prefetch_host(self, h_ids: torch.Tensor, h_labels: torch.Tensor, B: Optional[int] = None):
        """Start the H2D copy of the NEXT batch (pinned int32 tensors) on the library's copy stream; the following
        step_host_pinned(..., h_ids=, h_labels=) with the same tensors consumes it without copying again."""
        B = self.shape.B if B is None else B
        _lib.check(self.lib.hpmn_prefetch_host(self.ctx, C.byref(self._cshape(B)), _ptr(h_ids), _ptr(h_labels),
                                               _ptr(self.workspace)), self.ctx)

    def table_grad_sources(self, B: Optional[int] = None):
        """Byte offsets inside the workspace of (id slot of the host entry point, dX of layer 0, dlast) -- what the embedding
        scatter of the last backward call consumed (hpmn_table_grad_sources)."""
        B = self.shape.B if B is None else B
        a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _lib.check(self.lib.hpmn_table_grad_sources(self.ctx, C.byref(self._cshape(B)), C.byref(a), C.byref(b), C.byref(c)), self.ctx)
        return int(a.value), int(b.value), int(c.value)

    def check_ids(self):
        """Device-path twin of the check hpmn_step_host_end does: forward / forward_backward embed an id outside
        [0, feature_size) as zeros and drop its gradient (they cannot raise without a sync); this reads the flag the
        gather left in scalars[3] (one 16-byte D2H, synchronises the stream) and raises like TF's GatherV2 does on CPU."""
        if float(self.scalars[_lib.S_IDERR].item()) != 0.0:
            raise _lib.HpmnError(_lib.HPMN_EINVAL, "an id is outside [0, feature_size=%d)" % self.shape.V)

    def apply_gradients(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, clip: float = 1.0):
        """clip_by_value(g,-1,1) + dense Adam over [dense params | table] (hpmn.py:209-214; the clip densifies the
        embedding gradient in TF1.4, so every table row is touched each step)."""
        if self.adam_m is None:
            self.adam_m = torch.zeros_like(self.flat)
            self.adam_v = torch.zeros_like(self.flat)
        self.adam_t += 1
        _lib.check(self.lib.hpmn_clip_adam(self.ctx, _ptr(self.flat), _ptr(self.flat_grad), _ptr(self.adam_m),
                                           _ptr(self.adam_v), self.flat.numel(), self.adam_t, lr, beta1, beta2, eps, clip,
                                           self._stream()), self.ctx)

    # ------------------------------------------------------------------ measurement
    def launch_count(self) -> int:
        return int(self.lib.hpmn_launch_count(self.ctx))

    def profile(self, on: bool):
        _lib.check(self.lib.hpmn_profile_enable(self.ctx, int(on)), self.ctx)

    def profile_read(self):
        ms = (C.c_float * len(_lib.K_FAMILIES))()
        calls = (C.c_int64 * len(_lib.K_FAMILIES))()
        _lib.check(self.lib.hpmn_profile_read(self.ctx, ms, calls), self.ctx)
        return {n: (float(ms[i]), int(calls[i])) for i, n in enumerate(_lib.K_FAMILIES)}

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.hpmn_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
