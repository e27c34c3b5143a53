"""HpmnDualEngine: the `user=True, item=True` graph of the reference (/root/reference/code/hpmn.py:432-465, 284-320): two
hierarchical periodic memories -- scope "User" over `user_inp` and scope "item" over `item_inp` -- read by their own multi-hop
attention, over ONE shared embedding table; `repre = concat([user_repre, item_repre])` feeds the prediction head and
`memory_loss = umloss + imloss`.

No reference configuration turns the item side on (hpmn.py:592,620,659), so this is composed on the host from the K1-K5 entry
points of the C ABI instead of being a fused step of its own:

    per side   hpmn_gather_fwd -> hpmn_memory_fwd -> hpmn_attn_fwd            (covreg accumulates into one scalar)
    head       hpmn_head_wide_fwd / _bwd over R = (H + D_user) + (H + D_item)
    per side   hpmn_attn_bwd -> hpmn_memory_bwd -> hpmn_gather_bwd            (both sides accumulate into one table gradient)

PyTorch only owns the buffers; every kernel is libhpmn_b200's."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .layout import HpmnShape, param_layout

HEAD_NAMES = ["output/bn1/gamma", "output/bn1/beta", "output/fc1/kernel", "output/fc1/bias", "output/fc2/kernel", "output/fc2/bias",
              "output/fc3/kernel", "output/fc3/bias"]


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _Side:
    """Buffers of one memory side.  Its flat parameter block uses the single-side layout of the C ABI (the trailing output/*
    tensors of that layout are unused here: the head is shared and wider)."""

    def __init__(self, lib, shape: HpmnShape, device: torch.device):
        self.shape, self.c = shape, shape.to_c()
        self.layout, self.n = param_layout(shape)
        f32 = dict(dtype=torch.float32, device=device)
        self.params = torch.zeros(self.n, **f32)
        self.grads = torch.zeros(self.n, **f32)
        ws = lib.hpmn_workspace_bytes(C.byref(self.c), 1)
        if ws == 0:
            raise ValueError("invalid shape for libhpmn_b200: %r" % (shape,))
        self.ws = torch.empty(ws, dtype=torch.uint8, device=device)
        B, L, H, D = shape.B, shape.L, shape.H, shape.D
        self.x = torch.empty(B, shape.Tpad, D, **f32)
        self.dx = torch.empty(B, shape.Tpad, D, **f32)
        self.memory = torch.empty(B, L, H, **f32)
        self.dmemory = torch.empty(B, L, H, **f32)
        self.repre = torch.empty(B, H + D, **f32)
        self.drepre = torch.empty(B, H + D, **f32)
        self.dlast = torch.empty(B, D, **f32)
        self.w_hop0 = torch.empty(B, L, **f32)
        self.names = [n for n in self.layout if not n.startswith("output/")]

    def cshape(self, n: int):
        """Any batch up to shape.B reuses the buffers (rows are independent)."""
        if n == self.shape.B:
            return self.c
        if n <= 0 or n > self.shape.B:
            raise ValueError("batch %d outside (0, %d]" % (n, self.shape.B))
        c = self.shape.to_c()
        c.B = n
        return c

    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, shp = self.layout[name]
        return buf[off: off + int(np.prod(shp))].view(*shp)


class HpmnDualEngine:
    def __init__(self, user_shape: HpmnShape, item_shape: HpmnShape, device: int = 0, memory_reg: float = 1e-5,
                 l2_reg: float = 0.0, table: Optional[np.ndarray] = None, params: Optional[Dict[str, np.ndarray]] = None, seed: int = 4321):
        if not torch.cuda.is_available():
            raise RuntimeError("hpmn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if (user_shape.B, user_shape.E, user_shape.H, user_shape.V) != (item_shape.B, item_shape.E, item_shape.H, item_shape.V):
            raise ValueError("both sides share the batch, the embedding table and the hidden size")
        if user_shape.scope == item_shape.scope:
            raise ValueError("the two sides need distinct variable scopes (reference: 'User' and 'item')")
        self.lib = _lib.lib()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        ctx = C.c_void_p()
        _lib.check(self.lib.hpmn_create(C.byref(ctx), device))
        self.ctx = ctx
        self.memory_reg, self.l2_reg = float(memory_reg), float(l2_reg)
        self.user, self.item = _Side(self.lib, user_shape, self.device), _Side(self.lib, item_shape, self.device)
        self.B, self.V, self.E = user_shape.B, user_shape.V, user_shape.E
        self.R = (user_shape.H + user_shape.D) + (item_shape.H + item_shape.D)
        f32 = dict(dtype=torch.float32, device=self.device)
        n_head = self.lib.hpmn_head_wide_param_count(self.R)
        if n_head <= 0:
            raise ValueError("head input width %d is not supported" % self.R)
        offs, sizes = (C.c_int64 * 8)(), (C.c_int64 * 8)()
        self.lib.hpmn_head_wide_param_offsets(self.R, offs, sizes)
        shapes = [(self.R,), (self.R,), (self.R, 200), (200,), (200, 80), (80,), (80, 1), (1,)]
        self.head_layout = {n: (int(offs[i]), shapes[i]) for i, n in enumerate(HEAD_NAMES)}
        self.hparams = torch.zeros(n_head, **f32)
        self.hgrads = torch.zeros(n_head, **f32)
        self.hws = torch.empty(self.lib.hpmn_head_wide_workspace_bytes(self.B, self.R), dtype=torch.uint8, device=self.device)
        self.table = torch.zeros(self.V, self.E, **f32)
        self.dtable = torch.zeros(self.V, self.E, **f32)
        self.repre = torch.empty(self.B, self.R, **f32)
        self.drepre = torch.empty(self.B, self.R, **f32)
        self.scalars = torch.zeros(4, **f32)
        self.pred = torch.zeros(self.B, **f32)
        self.logit = torch.zeros(self.B, **f32)
        self._adam = None
        self.adam_t = 0
        self.init_parameters(seed)
        if params is not None:
            self.load_named(params)
        if table is not None:
            self.table.copy_(torch.as_tensor(table, dtype=torch.float32))

    # ------------------------------------------------------------------ parameters (TF variable names)
    def _views(self, grads: bool = False):
        out = {}
        for side in (self.user, self.item):
            for n in side.names:
                out[n] = side.view(side.grads if grads else side.params, n)
        buf = self.hgrads if grads else self.hparams
        for n, (off, shp) in self.head_layout.items():
            out[n] = buf[off: off + int(np.prod(shp))].view(*shp)
        return out

    def init_parameters(self, seed: int = 4321):
        """TF1.4 defaults (see HpmnEngine.init_parameters)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name, v in self._views().items():
            shp = tuple(v.shape)
            if name.endswith("kernel") or name.endswith("/map"):
                lim = float(np.sqrt(6.0 / (shp[0] + shp[1])))
                v.copy_((torch.rand(*shp, generator=g) * 2 - 1) * lim)
            elif name.endswith("gates/bias") or name.endswith("gamma"):
                v.fill_(1.0)
            else:
                v.zero_()
        lim = float(np.sqrt(6.0 / (self.V + self.E)))
        self.table.copy_((torch.rand(self.V, self.E, generator=g) * 2 - 1) * lim)

    def load_named(self, params: Dict[str, np.ndarray]):
        views = self._views()
        for name, arr in params.items():
            if name in views:
                views[name].copy_(torch.as_tensor(np.asarray(arr), dtype=torch.float32).reshape(views[name].shape))

    def named_parameters(self) -> Dict[str, np.ndarray]:
        return {n: v.detach().cpu().numpy().copy() for n, v in self._views().items()}

    def named_grads(self) -> Dict[str, np.ndarray]:
        return {n: v.detach().cpu().numpy().copy() for n, v in self._views(True).items()}

    # ------------------------------------------------------------------ the step
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _hyper(self, keep_prob, seed):
        return _lib.hpmn_hyper(self.memory_reg, 0.0, float(keep_prob), int(seed) & (2 ** 64 - 1), 0)

    def forward(self, user_ids: torch.Tensor, item_ids: torch.Tensor, labels: torch.Tensor, keep_prob: float = 1.0, seed: int = 0):
        """ids [B,T_side,F_side] int32 cuda; results in self.pred / logit / scalars (logloss, covreg of both sides, loss)."""
        lib, ctx, st = self.lib, self.ctx, self._stream()
        n = int(labels.shape[0])
        self.scalars.zero_()
        for side, ids in ((self.user, user_ids), (self.item, item_ids)):
            c = C.byref(side.cshape(n))
            _lib.check(lib.hpmn_gather_fwd(ctx, c, _ptr(ids), _ptr(self.table), _ptr(side.x), st), ctx)
            _lib.check(lib.hpmn_memory_fwd(ctx, c, _ptr(side.x), _ptr(side.params), _ptr(side.memory), _ptr(side.ws), st), ctx)
            _lib.check(lib.hpmn_attn_fwd(ctx, c, _ptr(side.memory), _ptr(side.x), _ptr(side.params), _ptr(side.repre), _ptr(side.w_hop0),
                                         _ptr(self.scalars), _ptr(side.ws), st), ctx)
        Ru = self.user.repre.shape[1]
        self.repre[:n, :Ru].copy_(self.user.repre[:n])           # concat([user_repre, item_repre]), hpmn.py:454
        self.repre[:n, Ru:].copy_(self.item.repre[:n])
        hy = self._hyper(keep_prob, seed)
        _lib.check(lib.hpmn_head_wide_fwd(ctx, n, self.R, C.byref(hy), _ptr(self.repre), _ptr(labels), _ptr(self.hparams),
                                          _ptr(self.pred), _ptr(self.logit), _ptr(self.scalars), _ptr(self.hws), st), ctx)
        self.scalars[2] = self.scalars[0] + self.memory_reg * self.scalars[1]      # loss = logloss + memory_reg * memory_loss
        if self.l2_reg != 0.0:      # + l2_reg * tf.nn.l2_loss(v) over every trainable, the table included (hpmn.py:204-205)
            for p in self._l2_buffers():
                self.scalars[2] += 0.5 * self.l2_reg * (p[0] * p[0]).sum()

    def _l2_buffers(self):
        """(variables, gradients) the l2 term covers; the unused output/* slots of the side blocks are zeros"""
        return [(self.user.params, self.user.grads), (self.item.params, self.item.grads), (self.hparams, self.hgrads),
                (self.table.view(-1), self.dtable.view(-1))]

    def forward_backward(self, user_ids: torch.Tensor, item_ids: torch.Tensor, labels: torch.Tensor, keep_prob: float = 1.0,
                         seed: int = 0):
        """+ gradients of the loss w.r.t. every trainable of both sides, the head and the table (tf.gradients, hpmn.py:211)."""
        self.forward(user_ids, item_ids, labels, keep_prob, seed)
        lib, ctx, st = self.lib, self.ctx, self._stream()
        n = int(labels.shape[0])
        hy = self._hyper(keep_prob, seed)
        self.hgrads.zero_(); self.dtable.zero_()
        _lib.check(lib.hpmn_head_wide_bwd(ctx, n, self.R, C.byref(hy), _ptr(self.repre), _ptr(labels), _ptr(self.hparams),
                                          _ptr(self.drepre), _ptr(self.hgrads), _ptr(self.hws), st), ctx)
        Ru = self.user.repre.shape[1]
        self.user.drepre[:n].copy_(self.drepre[:n, :Ru])
        self.item.drepre[:n].copy_(self.drepre[:n, Ru:])
        for side, ids in ((self.user, user_ids), (self.item, item_ids)):
            c = C.byref(side.cshape(n))
            side.grads.zero_()
            _lib.check(lib.hpmn_attn_bwd(ctx, c, C.byref(hy), _ptr(side.memory), _ptr(side.x), _ptr(side.params), _ptr(side.drepre),
                                         _ptr(side.dmemory), _ptr(side.dlast), _ptr(side.grads), _ptr(side.ws), st), ctx)
            _lib.check(lib.hpmn_memory_bwd(ctx, c, _ptr(side.x), _ptr(side.params), _ptr(side.dmemory), _ptr(side.dx), _ptr(side.grads),
                                           _ptr(side.ws), st), ctx)
            _lib.check(lib.hpmn_gather_bwd(ctx, c, _ptr(ids), _ptr(side.dx), _ptr(side.dlast), _ptr(self.dtable), st), ctx)
        if self.l2_reg != 0.0:
            for p, g in self._l2_buffers():
                g.add_(p, alpha=self.l2_reg)

    def apply_gradients(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, clip: float = 1.0):
        """clip_by_value + Adam over every trainable (hpmn.py:209-214); the unused output/* slots of the side blocks have zero
        gradients and stay at their initial value."""
        bufs = self._l2_buffers()
        if self._adam is None:
            self._adam = [(torch.zeros_like(p), torch.zeros_like(p)) for p, _ in bufs]
        self.adam_t += 1
        for (p, g), (m, v) in zip(bufs, self._adam):
            _lib.check(self.lib.hpmn_clip_adam(self.ctx, _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), self.adam_t, lr, beta1, beta2,
                                               eps, clip, self._stream()), self.ctx)

    def launch_count(self) -> int:
        return int(self.lib.hpmn_launch_count(self.ctx))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.hpmn_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
