"""Shared parity harness: run the CUDA path (through the C ABI) and the oracle on identical seeded
inputs and report the error of every output and gradient.  Used by the -m gpu tests, by
__graft_entry__.smoke()."""
from __future__ import annotations

import numpy as np

from oracle import hpmn_oracle as O


def oracle_shape(sh) -> O.OracleShape:
    return O.OracleShape(B=sh.B, T=sh.T, F=sh.F, E=sh.E, H=sh.H, periods=list(sh.periods), L=sh.L, hops=sh.hops, V=sh.V,
                         front_pad=sh.front_pad, mask_id0=sh.mask_id0, last_offset=sh.last_offset, scope=sh.scope)


def rel_max(a, ref):
    """max |a-ref| / (|ref| + 1e-6)  """
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(a - ref) / (np.abs(ref) + 1e-6))) if a.size else 0.0


def rel_l2(a, ref, floor=1e-6):
    """||a-ref|| / (||ref|| + floor).  The floor matters for tensors whose true gradient is identically zero
    (the last attention-MLP bias: softmax is shift invariant, so d loss / d bias == 0 and fp32 leaves ~1e-9)."""
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.linalg.norm(a - ref) / (np.linalg.norm(ref) + floor))


def abs_scaled(a, ref):
    """max |a-ref| / max|ref| -- for tensors with entries near zero where a per-element ratio is meaningless"""
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(a - ref)) / (np.max(np.abs(ref)) + 1e-30))


def run_case(sh, memory_reg=1e-3, mode="stress", ragged=True, seed_data=1234, seed_params=4321, keep_prob=1.0,
             host_path=False, device=0):
    """Returns dict of errors (CUDA vs fp64 oracle) for one configuration."""
    import torch
    from hpmn_b200.engine import HpmnEngine

    osh = oracle_shape(sh)
    params, table = O.init_params(osh, seed=seed_params, mode=mode, dtype=np.float32)
    ids, labels = O.synthetic_batch(osh, seed=seed_data, ragged=ragged)
    fwd = O.forward(osh, params, table, ids, labels, memory_reg=memory_reg, dtype=np.float64)
    g_ref, dt_ref = O.backward(osh, fwd, ids, labels, memory_reg=memory_reg, guard_zero_norm=(sh.L == 1))

    eng = HpmnEngine(sh, device=device, memory_reg=memory_reg, table=table, params=params)
    if host_path:
        eng.step_host(ids, labels, with_backward=True, keep_prob=keep_prob)
        scal, pred = eng.h_scalars.numpy().copy(), eng.h_pred.numpy().copy()
        logit, w0 = eng.h_logit.numpy().copy(), eng.h_w_hop0.numpy().copy()
        torch.cuda.synchronize()
        memory = None
    else:
        d_ids = torch.as_tensor(ids, device=eng.device)
        d_lab = torch.as_tensor(labels, device=eng.device)
        eng.forward_backward(d_ids, d_lab, keep_prob=keep_prob)
        torch.cuda.synchronize()
        scal, pred = eng.scalars.cpu().numpy(), eng.pred.cpu().numpy()
        logit, w0 = eng.logit.cpu().numpy(), eng.w_hop0.cpu().numpy()
        memory = eng.memory.cpu().numpy()
    out = {}
    if memory is not None:
        out["memory"] = rel_max(memory, fwd["memory"])
    out["pred"] = rel_max(pred, fwd["pred"])
    out["logit"] = rel_max(logit, fwd["logit"])
    out["w_hop0"] = rel_max(w0, fwd["w_hop0"])
    out["logloss"] = rel_max(scal[0], fwd["logloss"])
    out["covreg"] = rel_max(scal[1], fwd["covreg"])
    out["loss"] = rel_max(scal[2], fwd["loss"])
    grads = eng.named_grads()
    # per-tensor L2 error with the floor of tests/test_gpu_parity.py::_grad_close: a tensor whose true gradient is
    # identically zero (the last attention-MLP bias: softmax is shift invariant) is measured against the scale of the
    # largest gradient entry of the step instead of against its own ~1e-9 round-off
    gmax = max(float(np.abs(v).max()) for v in g_ref.values())
    worst, worst_name = 0.0, ""
    for k, v in g_ref.items():
        e = rel_l2(grads[k], v, floor=1e-4 * gmax)
        out["grad:" + k] = e
        if e > worst:
            worst, worst_name = e, k
    out["grad_worst"] = worst
    out["grad_worst_name"] = worst_name
    out["dtable"] = rel_l2(eng.dtable.cpu().numpy(), dt_ref, floor=1e-4 * gmax)
    out["launches"] = eng.launch_count()
    eng.close()
    return out
