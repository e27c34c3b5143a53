"""CPU oracle for the HPMN forward/backward hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference TF1.4 graph built by
/root/reference/code/hpmn.py.  It is the checker the CUDA path is compared with; it is
never the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product package (hpmn_b200/) must not.

PARITY: GRAPH WIRING PINNED, TF OP ARITHMETIC UNPINNED.  The reference ships no tests, golden vectors or recorded
outputs for this path, TensorFlow 1.4 and Python 2 cannot be installed in this image, and the
arithmetic of the ops lives in that un-vendored third-party dependency (readme.md:17 pins it only as "Tensorflow 1.4").
What is pinned: tests/golden/refgraph_*.npz are produced by importing the reference's code/hpmn.py UNMODIFIED on a TF1-API
stand-in (tests/golden/tf1_shim.py, generator tests/golden/make_reference_graph_fixture.py) and recording what its own
graph computes; tests/test_reference_graph.py checks this file against them (outputs 1e-13, gradients 1e-8, two
clip + Adam steps).  What is not: the arithmetic inside GRUCell / dense / batch_normalization / log_loss / Adam, which the
stand-in restates from TF1.4's published behaviour just like this file does -- no real TF1.4 session ever ran.
The restatement follows, in order of authority:
  1. code/hpmn.py:113-214, 266-320, 414-465   graph wiring, shapes, constants          (pinned, see above)
  2. code/util.py:81-110 (minus line 108)     in-tree copy of the TF1.4 GRUCell arithmetic
  3. code/rnn.py:588, 627-807                 dynamic_rnn loop semantics (zero state, no length mask)
  4. code/util.py:152-159                     front padding of the input tuples
  5. published TF1.4 defaults (glorot-uniform kernels, gate bias 1, BN eps 1e-3 momentum .99
     training=False, log_loss eps 1e-7 mean reduction, dropout scales by 1/keep_prob)
Each function cites the reference lines it restates.  Self-made golden vectors (fp64 run of this
file, tests/golden/{amazon,taobao,xlong}_*.npz) pin the CUDA path to THIS restatement.

Everything is written batch-vectorised with an explicit Python loop over time steps, in a
caller-chosen dtype (np.float64 for the checker, np.float32 to measure fp32 round-off).
Parameters are a dict keyed by the TF variable names the reference's scopes would produce
(SURVEY.md appendix B).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ATT_FC1 = 80   # hpmn.py:137
ATT_FC2 = 40   # hpmn.py:138
HEAD_FC1 = 200  # hpmn.py:191
HEAD_FC2 = 80   # hpmn.py:193
BN_EPS = 1e-3   # tf.layers.batch_normalization default epsilon [TF1.4]
LOGLOSS_EPS = 1e-7  # tf.losses.log_loss default epsilon [TF1.4]


@dataclass
class OracleShape:
    """Static configuration of one memory side (the `User` scope of hpmn.py:436-442 / 287-295)."""
    B: int            # batch
    T: int            # id steps fed by the loader (user_maxlen, hpmn.py:249)
    F: int            # id features per step (user_dim)
    E: int            # embedding_size
    H: int            # hidden_size
    periods: Sequence[int]  # li_layer[:L-1] of hpmn.py:113 (period of layer k, k < L-1)
    L: int            # num_layer
    hops: int         # self.hop
    V: int            # feature_size
    front_pad: int = 0      # Hpmn_Industry prepends 23 zero steps (hpmn.py:288-289); Hpmn: 0
    mask_id0: bool = True   # Hpmn multiplies by mask_lookup_table (hpmn.py:417-423); Industry: no mask
    last_offset: int = 1    # Hpmn: uinp[:, -1] (hpmn.py:439); Industry: uinp[:, -2] (hpmn.py:292)
    scope: str = "User"

    @property
    def D(self) -> int:
        return self.F * self.E

    @property
    def Tpad(self) -> int:
        return self.T + self.front_pad

    def steps(self) -> List[int]:
        """Steps run by each layer: maxlen /= li_layer[i] (hpmn.py:122-123)."""
        s, out = self.Tpad, []
        for k in range(self.L):
            out.append(s)
            if k < self.L - 1:
                p = self.periods[k]
                if s % p:
                    raise ValueError("layer %d: %d steps not divisible by period %d" % (k, s, p))
                s //= p
        return out


def param_names(sh: OracleShape) -> "OrderedDict[str, Tuple[int, ...]]":
    """TF variable names -> shapes (scopes at hpmn.py:117,173-174,137-139,190-195,433-465)."""
    o: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    H, D = sh.H, sh.D
    for k in range(sh.L):
        din = D if k == 0 else H
        base = "%s/GRU%d/rnn/gru_cell/" % (sh.scope, k)
        o[base + "gates/kernel"] = (din + H, 2 * H)
        o[base + "gates/bias"] = (2 * H,)
        o[base + "candidate/kernel"] = (din + H, H)
        o[base + "candidate/bias"] = (H,)
    o[sh.scope + "/dense/kernel"] = (D, H)
    o[sh.scope + "/dense/bias"] = (H,)
    o[sh.scope + "/map"] = (H, H)
    n = 1
    for _ in range(sh.hops):
        for (a, b) in ((4 * H, ATT_FC1), (ATT_FC1, ATT_FC2), (ATT_FC2, 1)):
            o["%s/dense_%d/kernel" % (sh.scope, n)] = (a, b)
            o["%s/dense_%d/bias" % (sh.scope, n)] = (b,)
            n += 1
    R = H + D
    o["output/bn1/gamma"] = (R,)
    o["output/bn1/beta"] = (R,)
    o["output/fc1/kernel"] = (R, HEAD_FC1)
    o["output/fc1/bias"] = (HEAD_FC1,)
    o["output/fc2/kernel"] = (HEAD_FC1, HEAD_FC2)
    o["output/fc2/bias"] = (HEAD_FC2,)
    o["output/fc3/kernel"] = (HEAD_FC2, 1)
    o["output/fc3/bias"] = (1,)
    return o


def _glorot(rng, shape):
    fan_in, fan_out = shape[0], shape[1]
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape)


def init_params(sh: OracleShape, seed: int = 4321, mode: str = "tf_default",
                dtype=np.float32) -> Tuple[Dict[str, np.ndarray], np.ndarray]:
    """TF1.4 default initialisation (glorot-uniform kernels, gate bias 1.0 per util.py:84-86,
    other biases 0, BN gamma 1 / beta 0) or a "stress" initialisation with larger embeddings,
    non-trivial biases and BN affine so that parity runs exercise every nonlinearity.
    Returns (params dict, embedding table [V,E])."""
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    for name, shape in param_names(sh).items():
        if name.endswith("kernel") or name.endswith("/map"):
            w = _glorot(rng, shape)
            if mode == "stress":
                w = w * 1.5
        elif name.endswith("gates/bias"):
            w = np.ones(shape)
            if mode == "stress":
                w = w + rng.uniform(-0.3, 0.3, size=shape)
        elif name.endswith("gamma"):
            w = np.ones(shape)
            if mode == "stress":
                w = w + rng.uniform(-0.2, 0.2, size=shape)
        else:
            w = np.zeros(shape)
            if mode == "stress":
                w = rng.uniform(-0.1, 0.1, size=shape)
        p[name] = np.ascontiguousarray(w, dtype=dtype)
    if mode == "stress":
        table = rng.uniform(-0.5, 0.5, size=(sh.V, sh.E))
    else:
        lim = np.sqrt(6.0 / (sh.V + sh.E))  # glorot-uniform on [V,E] (hpmn.py:415-416) [TF1.4 default]
        table = rng.uniform(-lim, lim, size=(sh.V, sh.E))
    return p, np.ascontiguousarray(table, dtype=dtype)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------

def embed(sh: OracleShape, table: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """hpmn.py:414-430 (Hpmn: gather * id-0 mask) / 266-282 (Industry: plain gather), then the
    zero front padding of hpmn.py:288-289.  ids [B,T,F] int -> x [B,Tpad,F*E]."""
    B, T, F = ids.shape
    x = table[ids.reshape(-1)].reshape(B, T, F, sh.E)
    if sh.mask_id0:
        x = x * (ids != 0)[..., None].astype(table.dtype)
    x = x.reshape(B, T, F * sh.E)
    if sh.front_pad:
        x = np.concatenate([np.zeros((B, sh.front_pad, F * sh.E), dtype=x.dtype), x], axis=1)
    return x


def gru_layer_fwd(x: np.ndarray, Wg, bg, Wc, bc):
    """One tf.nn.dynamic_rnn(GRUCell(H)) from zero state with no sequence_length
    (hpmn.py:118-120; cell arithmetic util.py:81-110 w/o line 108; loop rnn.py:732-793).
    x [B,S,Din] -> outputs [B,S,H] plus the saved gates."""
    B, S, Din = x.shape
    H = bc.shape[0]
    dt = x.dtype
    h = np.zeros((B, H), dtype=dt)
    hs = np.empty((B, S, H), dtype=dt)
    rs = np.empty((B, S, H), dtype=dt)
    us = np.empty((B, S, H), dtype=dt)
    cs = np.empty((B, S, H), dtype=dt)
    for s in range(S):
        xs = x[:, s]
        g = _sigmoid(np.concatenate([xs, h], axis=1) @ Wg + bg)      # util.py:95
        r, u = g[:, :H], g[:, H:]                                      # util.py:96 (r first, then u)
        c = np.tanh(np.concatenate([xs, r * h], axis=1) @ Wc + bc)    # util.py:98-107
        h = u * h + (1.0 - u) * c                                      # util.py:109
        hs[:, s], rs[:, s], us[:, s], cs[:, s] = h, r, u, c
    return hs, rs, us, cs


def build_memory_fwd(sh: OracleShape, p, x):
    """hpmn.py:113-131: L stacked GRUs, layer k+1 sees every periods[k]-th output of layer k,
    memory = stack of final states."""
    saved = []
    inp = x
    finals = []
    for k in range(sh.L):
        base = "%s/GRU%d/rnn/gru_cell/" % (sh.scope, k)
        hs, rs, us, cs = gru_layer_fwd(inp, p[base + "gates/kernel"], p[base + "gates/bias"],
                                       p[base + "candidate/kernel"], p[base + "candidate/bias"])
        saved.append((inp, hs, rs, us, cs))
        finals.append(hs[:, -1])
        if k < sh.L - 1:
            pk = sh.periods[k]
            inp = hs[:, pk - 1::pk]            # reshape [-1,S/p,p,H] + gather(p-1), hpmn.py:124-128
    memory = np.stack(finals, axis=1)          # [B,L,H]  hpmn.py:121,129
    return memory, saved


def covreg_fwd(memory):
    """hpmn.py:161-170."""
    H = memory.shape[2]
    mc = memory - memory.mean(axis=2, keepdims=True)
    C = mc @ mc.transpose(0, 2, 1) / memory.dtype.type(H)
    L = C.shape[1]
    off = C * (1.0 - np.eye(L, dtype=memory.dtype))
    nrm = np.sqrt((off * off).sum(axis=(1, 2)))
    return nrm.sum(), (mc, off, nrm)


def attention_fwd(memory, q, A1, a1, A2, a2, A3, a3):
    """hpmn.py:133-146 (key = value = memory)."""
    B, L, H = memory.shape
    Q = np.broadcast_to(q[:, None, :], (B, L, H))
    inp = np.concatenate([Q, memory, Q - memory, Q * memory], axis=-1)
    z1 = np.maximum(inp @ A1 + a1, 0)
    z2 = np.maximum(z1 @ A2 + a2, 0)
    s = (z2 @ A3 + a3)[..., 0]
    s = s - s.max(axis=1, keepdims=True)
    e = np.exp(s)
    w = e / e.sum(axis=1, keepdims=True)
    read = (memory * w[..., None]).sum(axis=1)
    return read, w, (inp, z1, z2)


def query_memory_fwd(sh: OracleShape, p, memory, last):
    """hpmn.py:172-182."""
    sc = sh.scope
    q = last @ p[sc + "/dense/kernel"] + p[sc + "/dense/bias"]
    qs, hop_saved, weights = [q], [], []
    for hop in range(sh.hops):
        n = 3 * hop
        read, w, sv = attention_fwd(memory, q,
                                    p["%s/dense_%d/kernel" % (sc, n + 1)], p["%s/dense_%d/bias" % (sc, n + 1)],
                                    p["%s/dense_%d/kernel" % (sc, n + 2)], p["%s/dense_%d/bias" % (sc, n + 2)],
                                    p["%s/dense_%d/kernel" % (sc, n + 3)], p["%s/dense_%d/bias" % (sc, n + 3)])
        q = q @ p[sc + "/map"] + read
        qs.append(q)
        hop_saved.append((w,) + sv)
        weights.append(w)
    return q, weights[0], (qs, hop_saved)


def _elu(x):
    return np.where(x > 0, x, np.exp(np.minimum(x, 0)) - 1.0)


def head_fwd(p, repre, labels, keep_prob=1.0, masks=None):
    """hpmn.py:190-202.  BN is built with the default training=False, so it normalises with the
    never-updated moving stats (mean 0, var 1): y = gamma * x / sqrt(1 + 1e-3) + beta.
    masks: optional (m1 [B,200], m2 [B,80]) 0/1 keep masks; dropout scales kept units by 1/keep_prob."""
    dt = repre.dtype
    inv = dt.type(1.0) / np.sqrt(dt.type(1.0) + dt.type(BN_EPS))
    bn = repre * inv * p["output/bn1/gamma"] + p["output/bn1/beta"]
    a1 = bn @ p["output/fc1/kernel"] + p["output/fc1/bias"]
    f1 = _elu(a1)
    d1 = f1 if masks is None else f1 * masks[0].astype(dt) / dt.type(keep_prob)
    a2 = d1 @ p["output/fc2/kernel"] + p["output/fc2/bias"]
    f2 = _elu(a2)
    d2 = f2 if masks is None else f2 * masks[1].astype(dt) / dt.type(keep_prob)
    logit = (d2 @ p["output/fc3/kernel"] + p["output/fc3/bias"])[:, 0]
    pred = _sigmoid(logit)
    y = labels.astype(dt)
    eps = dt.type(LOGLOSS_EPS)
    ll = (-y * np.log(pred + eps) - (1.0 - y) * np.log(1.0 - pred + eps)).mean()   # tf.losses.log_loss
    return logit, pred, ll, (bn, a1, f1, d1, a2, f2, d2)


def forward(sh: OracleShape, params, table, ids, labels, memory_reg=1e-5, l2_reg=0.0,
            keep_prob=1.0, masks=None, dtype=np.float64):
    """Whole graph of hpmn.py:432-465 / 284-320 (user=True, item=False) + build_fc_net(reg=True)."""
    p = {k: v.astype(dtype) for k, v in params.items()}
    tb = table.astype(dtype)
    x = embed(sh, tb, ids)
    memory, gru_saved = build_memory_fwd(sh, p, x)
    covreg, cov_saved = covreg_fwd(memory)
    last = x[:, -sh.last_offset, :]
    q, w0, att_saved = query_memory_fwd(sh, p, memory, last)
    repre = np.concatenate([q, last], axis=-1)
    logit, pred, ll, head_saved = head_fwd(p, repre, labels, keep_prob, masks)
    loss = ll + memory_reg * covreg
    if l2_reg:
        loss = loss + l2_reg * 0.5 * (sum((v * v).sum() for v in p.values()) + (tb * tb).sum())
    return dict(x=x, memory=memory, covreg=covreg, last=last, q=q, w_hop0=w0, repre=repre,
                logit=logit, pred=pred, logloss=ll, loss=loss,
                _saved=(p, tb, gru_saved, cov_saved, att_saved, head_saved))


# --------------------------------------------------------------------------------------
# backward (hand-derived adjoint; SURVEY.md appendix C for the GRU step)
# --------------------------------------------------------------------------------------

def gru_layer_bwd(x, hs, rs, us, cs, Wg, Wc, dhs):
    """Adjoint of gru_layer_fwd.  dhs [B,S,H] is the gradient arriving at every output of the
    layer (upper layer's dx at firing steps + the memory-slot gradient at the last step)."""
    B, S, Din = x.shape
    H = hs.shape[2]
    dt = x.dtype
    dWg = np.zeros_like(Wg); dbg = np.zeros(2 * H, dtype=dt)
    dWc = np.zeros_like(Wc); dbc = np.zeros(H, dtype=dt)
    dx = np.empty_like(x)
    dh_next = np.zeros((B, H), dtype=dt)
    for s in range(S - 1, -1, -1):
        h_prev = hs[:, s - 1] if s > 0 else np.zeros((B, H), dtype=dt)
        r, u, c, xs = rs[:, s], us[:, s], cs[:, s], x[:, s]
        dh = dhs[:, s] + dh_next
        dc = dh * (1.0 - u)
        du = dh * (h_prev - c)
        dh_prev = dh * u
        dac = dc * (1.0 - c * c)
        xrh = np.concatenate([xs, r * h_prev], axis=1)
        dWc += xrh.T @ dac; dbc += dac.sum(axis=0)
        dxrh = dac @ Wc.T
        dx_s = dxrh[:, :Din]
        drh = dxrh[:, Din:]
        dr = drh * h_prev
        dh_prev = dh_prev + drh * r
        dag = np.concatenate([dr * r * (1.0 - r), du * u * (1.0 - u)], axis=1)
        xh = np.concatenate([xs, h_prev], axis=1)
        dWg += xh.T @ dag; dbg += dag.sum(axis=0)
        dxh = dag @ Wg.T
        dx[:, s] = dx_s + dxh[:, :Din]
        dh_next = dh_prev + dxh[:, Din:]
    return dx, dWg, dbg, dWc, dbc


def backward(sh: OracleShape, fwd, ids, labels, memory_reg=1e-5, l2_reg=0.0, keep_prob=1.0,
             masks=None, loss_scale_B: Optional[int] = None, guard_zero_norm: bool = False):
    """Gradient of `loss` w.r.t. every trainable variable (tf.gradients at hpmn.py:211) BEFORE the
    clip of hpmn.py:212.  Returns (grads dict, dtable [V,E] dense -- the clip densifies the
    IndexedSlices in TF1.4).  loss_scale_B: batch size the log-loss mean divides by (the global
    batch when a rank only holds a shard); default: local B.
    guard_zero_norm: tf.norm's gradient is 0/0 = NaN when a sample's off-diagonal covariance is exactly zero
    (always the case for L == 1); False reproduces that, True yields 0 like the CUDA path."""
    p, tb, gru_saved, cov_saved, att_saved, head_saved = fwd["_saved"]
    dt = tb.dtype
    B = ids.shape[0]
    Bn = dt.type(loss_scale_B if loss_scale_B else B)
    H, D, sc = sh.H, sh.D, sh.scope
    g: Dict[str, np.ndarray] = {k: np.zeros_like(v) for k, v in p.items()}

    # ---- head (hpmn.py:190-202)
    bn, a1, f1, d1, a2, f2, d2 = head_saved
    pred = fwd["pred"]
    y = labels.astype(dt)
    eps = dt.type(LOGLOSS_EPS)
    dpred = (-y / (pred + eps) + (1.0 - y) / (1.0 - pred + eps)) / Bn
    dlogit = dpred * pred * (1.0 - pred)
    g["output/fc3/kernel"] += d2.T @ dlogit[:, None]
    g["output/fc3/bias"] += dlogit.sum(keepdims=True)
    dd2 = dlogit[:, None] @ p["output/fc3/kernel"].T
    df2 = dd2 if masks is None else dd2 * masks[1].astype(dt) / dt.type(keep_prob)
    da2 = df2 * np.where(a2 > 0, 1.0, f2 + 1.0)
    g["output/fc2/kernel"] += d1.T @ da2
    g["output/fc2/bias"] += da2.sum(axis=0)
    dd1 = da2 @ p["output/fc2/kernel"].T
    df1 = dd1 if masks is None else dd1 * masks[0].astype(dt) / dt.type(keep_prob)
    da1 = df1 * np.where(a1 > 0, 1.0, f1 + 1.0)
    g["output/fc1/kernel"] += bn.T @ da1
    g["output/fc1/bias"] += da1.sum(axis=0)
    dbn = da1 @ p["output/fc1/kernel"].T
    inv = dt.type(1.0) / np.sqrt(dt.type(1.0) + dt.type(BN_EPS))
    repre = fwd["repre"]
    g["output/bn1/gamma"] += (dbn * repre * inv).sum(axis=0)
    g["output/bn1/beta"] += dbn.sum(axis=0)
    drepre = dbn * inv * p["output/bn1/gamma"]
    dq = drepre[:, :H].copy()
    dlast = drepre[:, H:].copy()

    # ---- query_memory (hpmn.py:172-182), hops in reverse
    memory = fwd["memory"]
    L = memory.shape[1]
    dmem = np.zeros_like(memory)
    qs, hop_saved = att_saved
    Hmap = p[sc + "/map"]
    for hop in range(sh.hops - 1, -1, -1):
        n = 3 * hop
        A1 = p["%s/dense_%d/kernel" % (sc, n + 1)]
        A2 = p["%s/dense_%d/kernel" % (sc, n + 2)]
        A3 = p["%s/dense_%d/kernel" % (sc, n + 3)]
        w, inp, z1, z2 = hop_saved[hop]
        qin = qs[hop]
        # q_out = qin @ Hmap + read
        g[sc + "/map"] += qin.T @ dq
        dqin = dq @ Hmap.T
        dread = dq
        # read = sum_l w_l m_l
        dmem += dread[:, None, :] * w[..., None]
        dw = (memory * dread[:, None, :]).sum(axis=2)
        dscore = w * (dw - (dw * w).sum(axis=1, keepdims=True))
        ds3 = dscore[..., None]                                   # [B,L,1]
        g["%s/dense_%d/kernel" % (sc, n + 3)] += np.einsum("bli,blo->io", z2, ds3)
        g["%s/dense_%d/bias" % (sc, n + 3)] += ds3.sum(axis=(0, 1))
        dz2 = (ds3 @ A3.T) * (z2 > 0)
        g["%s/dense_%d/kernel" % (sc, n + 2)] += np.einsum("bli,blo->io", z1, dz2)
        g["%s/dense_%d/bias" % (sc, n + 2)] += dz2.sum(axis=(0, 1))
        dz1 = (dz2 @ A2.T) * (z1 > 0)
        g["%s/dense_%d/kernel" % (sc, n + 1)] += np.einsum("bli,blo->io", inp, dz1)
        g["%s/dense_%d/bias" % (sc, n + 1)] += dz1.sum(axis=(0, 1))
        dinp = dz1 @ A1.T                                         # [B,L,4H]
        dQ = dinp[..., :H] + dinp[..., 2 * H:3 * H] + dinp[..., 3 * H:] * memory
        dmem += dinp[..., H:2 * H] - dinp[..., 2 * H:3 * H] + dinp[..., 3 * H:] * qin[:, None, :]
        dq = dqin + dQ.sum(axis=1)
    # q0 = last @ Wq + bq
    last = fwd["last"]
    g[sc + "/dense/kernel"] += last.T @ dq
    g[sc + "/dense/bias"] += dq.sum(axis=0)
    dlast += dq @ p[sc + "/dense/kernel"].T

    # ---- covreg (hpmn.py:161-170): d(sum_b ||offdiag(C_b)||_F)
    mc, off, nrm = cov_saved
    with np.errstate(divide="ignore", invalid="ignore"):
        dC = off / nrm[:, None, None]           # 0/0 -> nan exactly as tf.norm's gradient would
    if guard_zero_norm:
        dC = np.where(nrm[:, None, None] > 0, dC, 0.0)
    dmc = dt.type(2.0) * (dC @ mc) / dt.type(H)
    dmem += dt.type(memory_reg) * (dmc - dmc.mean(axis=2, keepdims=True))

    # ---- build_memory (hpmn.py:113-131), top layer first
    dx_up = None
    for k in range(sh.L - 1, -1, -1):
        inp, hs, rs, us, cs = gru_saved[k]
        dhs = np.zeros_like(hs)
        dhs[:, -1] += dmem[:, k]
        if k < sh.L - 1:
            pk = sh.periods[k]
            dhs[:, pk - 1::pk] += dx_up
        base = "%s/GRU%d/rnn/gru_cell/" % (sc, k)
        dx_up, dWg, dbg, dWc, dbc = gru_layer_bwd(inp, hs, rs, us, cs, p[base + "gates/kernel"],
                                                  p[base + "candidate/kernel"], dhs)
        g[base + "gates/kernel"] += dWg; g[base + "gates/bias"] += dbg
        g[base + "candidate/kernel"] += dWc; g[base + "candidate/bias"] += dbc
    dx = dx_up                                   # [B,Tpad,D]
    dx[:, -sh.last_offset, :] += dlast

    # ---- embedding (hpmn.py:414-430 / 266-282): scatter-add, id-0 rows masked, front pad dropped
    dxe = dx[:, sh.front_pad:, :].reshape(B, sh.T, sh.F, sh.E)
    if sh.mask_id0:
        dxe = dxe * (ids != 0)[..., None].astype(dt)
    dtable = np.zeros_like(tb)
    np.add.at(dtable, ids.reshape(-1), dxe.reshape(-1, sh.E))

    if l2_reg:
        for k2 in g:
            g[k2] += dt.type(l2_reg) * p[k2]
        dtable += dt.type(l2_reg) * tb
    return g, dtable


# --------------------------------------------------------------------------------------
# update step (hpmn.py:209-214)  [TF1.4 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); v -= lr_t*m/(sqrt(v)+eps)]
# --------------------------------------------------------------------------------------

def clip_adam_step(var, grad, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """clip_by_value(grad,-1,1) then one dense AdamOptimizer apply (hpmn.py:212-214).  t starts at 1."""
    dt = var.dtype
    gcl = np.clip(grad, -1.0, 1.0)
    m[...] = b1 * m + (1 - b1) * gcl
    v[...] = b2 * v + (1 - b2) * gcl * gcl
    lr_t = dt.type(lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t))
    var[...] = var - lr_t * m / (np.sqrt(v) + dt.type(eps))


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------

def synthetic_batch(sh: OracleShape, seed: int = 1234, ragged: bool = True):
    """ids uniform on [1,V); for F >= 3 column 0 is constant along t (the uid column,
    preprocess_amazon.py:162); with ragged=True each sample keeps a random-length suffix and the
    prefix is id 0 (front padding, util.py:152-159).  Labels Bernoulli(0.5)."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(1, sh.V, size=(sh.B, sh.T, sh.F), dtype=np.int64)
    if sh.F >= 3:
        ids[:, :, 0] = ids[:, :1, 0]
    if ragged:
        lens = rng.integers(min(5, sh.T), sh.T + 1, size=sh.B)
        for b in range(sh.B):
            ids[b, : sh.T - lens[b]] = 0
    labels = rng.integers(0, 2, size=sh.B).astype(np.int32)
    return ids.astype(np.int32), labels
