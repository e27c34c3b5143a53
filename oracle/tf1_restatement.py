"""TF1-equivalent restatement of the HPMN graph in torch-CPU  --  TEST / BASELINE INFRASTRUCTURE ONLY.

PARITY: graph wiring pinned to the reference's own code/hpmn.py (tests/golden/refgraph_*.npz, tests/test_reference_graph.py),
TF1.4 op arithmetic unpinned (see oracle/hpmn_oracle.py header): TensorFlow 1.4 cannot be installed here, so
this is a restatement, NOT TensorFlow.  It exists for two reasons:
  * an independent check of the hand-derived adjoint in hpmn_oracle.backward (torch autograd);
  * the CPU baseline ("cpu_baseline" / `bench.py --impl reference`): the graph is executed the way
    TF1's executor runs it -- one explicit iteration per time step per layer with separate
    concat / matmul / bias-add / sigmoid / tanh ops like the while_loop body of
    /root/reference/code/rnn.py:732-793 driving the cell of code/util.py:81-110, layer after layer
    like code/hpmn.py:113-131, autograd backward standing in for tf.gradients -- in fp32 on all
    host threads.
Only tests/, __graft_entry__.smoke() and bench.py may import this file.
"""
from __future__ import annotations

import time
from typing import Dict

import numpy as np
import torch

from .hpmn_oracle import BN_EPS, LOGLOSS_EPS, OracleShape


def _t(p: Dict[str, np.ndarray], dtype, requires_grad=True):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in p.items()}


def side_torch(sh: OracleShape, p, table, ids, sparse_grad=False):
    """One memory side (scope sh.scope) of the graph: embedding -> build_memory -> get_covreg -> query_memory
    (hpmn.py:414-430 / 266-282, 113-131, 161-170, 172-182).  Returns x, memory, covreg, q, w_hop0, last, repre."""
    H, sc = sh.H, sh.scope
    B = ids.shape[0]
    # hpmn.py:414-430 / 266-282
    x = torch.nn.functional.embedding(ids, table, sparse=sparse_grad)
    if sh.mask_id0:
        x = x * (ids != 0).unsqueeze(-1).to(x.dtype)
    x = x.reshape(B, sh.T, sh.D)
    if sh.front_pad:
        x = torch.cat([torch.zeros(B, sh.front_pad, sh.D, dtype=x.dtype), x], dim=1)   # hpmn.py:288-289
    # hpmn.py:113-131
    inp = x
    finals = []
    for k in range(sh.L):
        base = "%s/GRU%d/rnn/gru_cell/" % (sc, k)
        Wg, bg = p[base + "gates/kernel"], p[base + "gates/bias"]
        Wc, bc = p[base + "candidate/kernel"], p[base + "candidate/bias"]
        xt = inp.transpose(0, 1)                      # time-major like rnn.py:560-563
        h = torch.zeros(B, H, dtype=x.dtype)          # zero_state, rnn.py:588
        outs = []
        # input_ta.unstack(input), rnn.py:722-725: one tensor per step (its gradient is a stack, O(T); indexing xt[s]
        # would make autograd materialise a full-size zero tensor per step, O(T^2), which TF's TensorArray does not)
        for xs in xt.unbind(0):                       # while_loop body, rnn.py:780-793
            g = torch.sigmoid(torch.matmul(torch.cat([xs, h], dim=1), Wg) + bg)
            r, u = torch.split(g, H, dim=1)
            c = torch.tanh(torch.matmul(torch.cat([xs, r * h], dim=1), Wc) + bc)
            h = u * h + (1 - u) * c
            outs.append(h)
        outputs = torch.stack(outs, dim=0).transpose(0, 1)     # rnn.py:620-622
        finals.append(h.unsqueeze(1))
        if k < sh.L - 1:
            pk = sh.periods[k]
            S = outputs.shape[1] // pk
            outputs = outputs.reshape(B, S, pk, H)
            inp = outputs[:, :, pk - 1, :]
    memory = torch.cat(finals, dim=1)
    # hpmn.py:161-170
    mc = memory - memory.mean(dim=2, keepdim=True)
    C = torch.matmul(mc, mc.transpose(1, 2)) / float(H)
    C = C - torch.diag_embed(torch.diagonal(C, dim1=1, dim2=2))
    covreg = torch.sqrt((C * C).sum(dim=(1, 2))).sum()
    # hpmn.py:172-182 / 133-146
    last = x[:, -sh.last_offset, :]
    q = torch.matmul(last, p[sc + "/dense/kernel"]) + p[sc + "/dense/bias"]
    w0 = None
    for hop in range(sh.hops):
        n = 3 * hop
        Q = q.unsqueeze(1).expand(-1, sh.L, -1)
        a = torch.cat([Q, memory, Q - memory, Q * memory], dim=-1)
        z1 = torch.relu(torch.matmul(a, p["%s/dense_%d/kernel" % (sc, n + 1)]) + p["%s/dense_%d/bias" % (sc, n + 1)])
        z2 = torch.relu(torch.matmul(z1, p["%s/dense_%d/kernel" % (sc, n + 2)]) + p["%s/dense_%d/bias" % (sc, n + 2)])
        s3 = torch.matmul(z2, p["%s/dense_%d/kernel" % (sc, n + 3)]) + p["%s/dense_%d/bias" % (sc, n + 3)]
        w = torch.softmax(s3.reshape(B, sh.L), dim=1)
        read = (memory * w.unsqueeze(2)).sum(dim=1)
        q = torch.matmul(q, p[sc + "/map"]) + read
        if hop == 0:
            w0 = w
    repre = torch.cat([q, last], dim=-1)
    return dict(x=x, memory=memory, covreg=covreg, q=q, w_hop0=w0, last=last, repre=repre)


def head_torch(p, repre, labels, keep_prob=1.0, masks=None):
    """build_fc_net, hpmn.py:190-202."""
    bn = repre / float(np.sqrt(1.0 + BN_EPS)) * p["output/bn1/gamma"] + p["output/bn1/beta"]
    f1 = torch.nn.functional.elu(torch.matmul(bn, p["output/fc1/kernel"]) + p["output/fc1/bias"])
    if masks is not None:
        f1 = f1 * masks[0] / keep_prob
    f2 = torch.nn.functional.elu(torch.matmul(f1, p["output/fc2/kernel"]) + p["output/fc2/bias"])
    if masks is not None:
        f2 = f2 * masks[1] / keep_prob
    logit = (torch.matmul(f2, p["output/fc3/kernel"]) + p["output/fc3/bias"]).reshape(-1)
    pred = torch.sigmoid(logit)
    ll = (-labels * torch.log(pred + LOGLOSS_EPS) - (1 - labels) * torch.log(1 - pred + LOGLOSS_EPS)).mean()
    return logit, pred, ll


def forward_torch(sh: OracleShape, p, table, ids, labels, memory_reg=1e-5, keep_prob=1.0, masks=None, sparse_grad=False):
    """p: dict of torch tensors, table: torch [V,E]; ids: torch int64 [B,T,F]; labels: torch float [B].
    sparse_grad: the table gradient is a sparse (indices, rows) tensor -- what tf.gradients returns for
    tf.nn.embedding_lookup (IndexedSlices, before the clip of code/hpmn.py:212 densifies it)."""
    s = side_torch(sh, p, table, ids, sparse_grad)
    logit, pred, ll = head_torch(p, s["repre"], labels, keep_prob, masks)
    loss = ll + memory_reg * s["covreg"]
    return dict(x=s["x"], memory=s["memory"], covreg=s["covreg"], q=s["q"], w_hop0=s["w_hop0"], logit=logit, pred=pred, logloss=ll,
                loss=loss)


def forward_torch_dual(sh_user: OracleShape, sh_item: OracleShape, p, table, ids_user, ids_item, labels, memory_reg=1e-5,
                       keep_prob=1.0, masks=None):
    """user=True, item=True (hpmn.py:432-465): both memory sides over the shared embedding table,
    repre = concat([user_repre, item_repre]), memory_loss = umloss + imloss."""
    u = side_torch(sh_user, p, table, ids_user)
    i = side_torch(sh_item, p, table, ids_item)
    repre = torch.cat([u["repre"], i["repre"]], dim=-1)
    logit, pred, ll = head_torch(p, repre, labels, keep_prob, masks)
    covreg = u["covreg"] + i["covreg"]
    loss = ll + memory_reg * covreg
    return dict(user=u, item=i, covreg=covreg, logit=logit, pred=pred, logloss=ll, loss=loss)


def forward_backward_numpy(sh: OracleShape, params, table, ids, labels, memory_reg=1e-5,
                           dtype=torch.float64, keep_prob=1.0, masks=None):
    """Convenience wrapper used by the tests: numpy in, numpy out, gradients by autograd."""
    p = _t(params, dtype)
    tb = torch.tensor(table, dtype=dtype, requires_grad=True)
    tmasks = None if masks is None else tuple(torch.tensor(m, dtype=dtype) for m in masks)
    out = forward_torch(sh, p, tb, torch.tensor(ids, dtype=torch.int64), torch.tensor(labels, dtype=dtype),
                        memory_reg, keep_prob, tmasks)
    out["loss"].backward()
    grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items()}
    res = {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    return res, grads, tb.grad.numpy()


def time_cpu_baseline(sh: OracleShape, params, table, ids, labels, iters=3, warmup=1, threads=None,
                      budget_s=None, memory_reg=1e-5, keep_prob=1.0, sparse_grad=True):
    """fwd+bwd samples/sec of the restatement in fp32 on `threads` host threads (default: all).
    Returns dict(value, ms_per_step, cores, iters).  The optimizer is excluded, like the GPU arm."""
    import os
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    p = _t(params, torch.float32)
    tb = torch.tensor(table, dtype=torch.float32, requires_grad=True)
    tid = torch.tensor(ids, dtype=torch.int64)
    tl = torch.tensor(labels, dtype=torch.float32)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + iters):
        for v in p.values():
            v.grad = None
        tb.grad = None
        t0 = time.perf_counter()
        masks = None
        if keep_prob < 1.0:      # tf.nn.dropout(keep_prob) of the train feed, code/hpmn.py:191-194,480
            masks = ((torch.rand(ids.shape[0], 200) < keep_prob).float(), (torch.rand(ids.shape[0], 80) < keep_prob).float())
        out = forward_torch(sh, p, tb, tid, tl, memory_reg, keep_prob, masks, sparse_grad=sparse_grad)
        out["loss"].backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if budget_s is not None and len(times) >= 1 and time.perf_counter() - t_start > budget_s:
            break
    med = float(np.median(times))
    return dict(value=ids.shape[0] / med, ms_per_step=med * 1e3, cores=threads, iters=len(times))
