"""-m gpu: the CUDA path against what the reference's OWN graph code computes (tests/golden/refgraph_*.npz, written by
tests/golden/make_reference_graph_fixture.py: /root/reference/code/hpmn.py imported unmodified on the TF1-API stand-in).
Variables are loaded by their TF names, the feeds are the fixture's, and every fetch of the reference's train() / eval()
is compared: prediction, log_loss, memory_loss, cross_entropy, hop-0 attention weights, compute_gradients() of every
trainable, and the variables after two train_step runs (clip_by_value + Adam).

Tolerances: those of tests/test_gpu_parity.py (outputs 1e-4 relative + 1e-6, gradients 1e-3 in L2 per tensor).  After
the optimiser steps an element-wise bound is used instead: Adam's first updates are lr * g / (|g| + eps) ~ +-lr, so an
entry whose gradient is fp32 round-off (|g| ~ 1e-9: e.g. the last attention bias, whose true gradient is 0) may move by
a fraction of lr in either direction; at least 99.5 % of the entries of every tensor must agree to 2 % of lr and none
may differ by more than the two steps can move it."""
import numpy as np
import pytest

from hpmn_b200.layout import HpmnShape
from tests import _refgraph as G
from tests.test_gpu_parity import _close, _grad_close

pytestmark = pytest.mark.gpu


def _update_close(got, after, lr, name):
    d = np.abs(np.asarray(got, np.float64) - np.asarray(after, np.float64))
    frac = float(np.mean(d <= 0.02 * lr + 1e-7 * np.abs(after)))
    assert frac >= 0.995 and d.max() <= 2 * 2.05 * lr, "%s: %.4f of the entries agree, max |d| %.3e (lr %.0e)" % (name, frac, d.max(), lr)


@pytest.mark.parametrize("name", ["refgraph_amazon", "refgraph_xlong"])
@pytest.mark.parametrize("host", [False, True], ids=["device_feed", "host_feed"])
def test_cuda_path_reproduces_the_reference_graph(name, host):
    import torch
    from hpmn_b200.engine import HpmnEngine
    z, c = G.load(name)
    sh = HpmnShape(**G.side_kwargs(c)[0])
    p = {k: v.astype(np.float32) for k, v in G.trainables(z).items()}
    tb = z["var:" + G.TABLE]
    eng = HpmnEngine(sh, device=0, memory_reg=c["memory_reg"], l2_reg=c["l2_reg"], table=tb, params=p)
    assert set(eng.named_parameters()) == set(p)

    def step(i):
        ids, lab = z["user_inp%d" % i], z["label%d" % i]
        if host:
            eng.step_host(ids, lab, with_backward=True)
            return eng.h_scalars.numpy().copy(), eng.h_pred.numpy().copy(), eng.h_w_hop0.numpy().reshape(sh.B, sh.L).copy()
        eng.forward_backward(torch.as_tensor(ids, device=eng.device), torch.as_tensor(lab, device=eng.device))
        torch.cuda.synchronize()
        return eng.scalars.cpu().numpy(), eng.pred.cpu().numpy(), eng.w_hop0.cpu().numpy()

    scal, pred, w0 = step(0)
    _close(pred, z["prediction"], "prediction")
    _close(w0, z["user_weights"], "user_weights")
    _close(scal[:3], [z["log_loss"], z["memory_loss"], z["cross_entropy"]], "log_loss, memory_loss, cross_entropy")
    ref = {k: z["grad:" + k] for k in p}
    ref[G.TABLE] = G.dense_rows(z, "grad", sh.V, sh.E)
    got = eng.named_grads(); got[G.TABLE] = eng.dtable.cpu().numpy()
    _grad_close(got, ref)
    untouched = np.setdiff1d(np.arange(sh.V), z["grad_rows"])
    assert not got[G.TABLE][untouched].any()

    eng.apply_gradients(c["lr"])
    step(1)
    eng.apply_gradients(c["lr"])
    torch.cuda.synchronize()
    now = eng.named_parameters()
    for k in p:
        _update_close(now[k], z["after:" + k], c["lr"], k)
    _update_close(eng.table.cpu().numpy(), G.dense_rows(z, "after", sh.V, sh.E, base=tb), c["lr"], G.TABLE)
    eng.step_host(z["user_inp0"], z["label0"], with_backward=False)
    assert np.abs(eng.h_pred.numpy()[: sh.B] - z["prediction_after"]).max() < 2e-3
    eng.close()


def test_tensor_core_recurrence_reproduces_the_reference_graph_at_1024_steps(monkeypatch):
    """HPMN_TCREC=1 forces the tcgen05 recurrence (normally chosen from 8192 rows per GPU) onto the XLong fixture: 1001 + 23
    steps, 5 layers, 3xTF32 gate GEMMs through 1024 dependent steps, against the reference's own graph code"""
    import torch
    from hpmn_b200.engine import HpmnEngine
    monkeypatch.setenv("HPMN_TCREC", "1")
    z, c = G.load("refgraph_xlong")
    sh = HpmnShape(**G.side_kwargs(c)[0])
    p = {k: v.astype(np.float32) for k, v in G.trainables(z).items()}
    eng = HpmnEngine(sh, device=0, memory_reg=c["memory_reg"], table=z["var:" + G.TABLE], params=p)
    eng.forward_backward(torch.as_tensor(z["user_inp0"], device=eng.device), torch.as_tensor(z["label0"], device=eng.device))
    torch.cuda.synchronize()
    _close(eng.pred.cpu().numpy(), z["prediction"], "prediction")
    _close(eng.w_hop0.cpu().numpy(), z["user_weights"], "user_weights")
    _close(eng.scalars.cpu().numpy()[:3], [z["log_loss"], z["memory_loss"], z["cross_entropy"]], "log_loss, memory_loss, cross_entropy")
    ref = {k: z["grad:" + k] for k in p}
    ref[G.TABLE] = G.dense_rows(z, "grad", sh.V, sh.E)
    got = eng.named_grads(); got[G.TABLE] = eng.dtable.cpu().numpy()
    _grad_close(got, ref)
    eng.close()


def test_cuda_path_reproduces_the_reference_graph_hidden_64():
    """hidden_size = 64: the tensor-core recurrence (tcrec.cu), the 64-lane attention kernels and the sliced weight-gradient
    kernel against the reference's own graph code"""
    import torch
    from hpmn_b200.engine import HpmnEngine
    z, c = G.load("refgraph_h64")
    sh = HpmnShape(**G.side_kwargs(c)[0])
    p = {k: v.astype(np.float32) for k, v in G.trainables(z).items()}
    eng = HpmnEngine(sh, device=0, memory_reg=c["memory_reg"], table=z["var:" + G.TABLE], params=p)
    eng.forward_backward(torch.as_tensor(z["user_inp0"], device=eng.device), torch.as_tensor(z["label0"], device=eng.device))
    torch.cuda.synchronize()
    _close(eng.pred.cpu().numpy(), z["prediction"], "prediction")
    _close(eng.w_hop0.cpu().numpy(), z["user_weights"], "user_weights")
    _close(eng.scalars.cpu().numpy()[:3], [z["log_loss"], z["memory_loss"], z["cross_entropy"]], "log_loss, memory_loss, cross_entropy")
    ref = {k: z["grad:" + k].astype(np.float64) for k in p}
    ref[G.TABLE] = G.dense_rows(z, "grad", sh.V, sh.E)
    got = eng.named_grads(); got[G.TABLE] = eng.dtable.cpu().numpy()
    _grad_close(got, ref)
    eng.close()


def test_cuda_path_reproduces_the_reference_graph_both_sides_with_l2():
    import torch
    from hpmn_b200.dual import HpmnDualEngine
    z, c = G.load("refgraph_dual")
    ku, ki = G.side_kwargs(c)
    us, it = HpmnShape(**ku), HpmnShape(**ki)
    p = {k: v.astype(np.float32) for k, v in G.trainables(z).items()}
    tb = z["var:" + G.TABLE]
    eng = HpmnDualEngine(us, it, device=0, memory_reg=c["memory_reg"], l2_reg=c["l2_reg"], table=tb, params=p)
    assert set(eng.named_parameters()) == set(p)
    dev = eng.device

    def step(i):
        eng.forward_backward(torch.as_tensor(z["user_inp%d" % i], device=dev), torch.as_tensor(z["item_inp%d" % i], device=dev),
                             torch.as_tensor(z["label%d" % i], device=dev))
        torch.cuda.synchronize()

    step(0)
    _close(eng.pred.cpu().numpy()[: us.B], z["prediction"], "prediction")
    _close(eng.user.w_hop0.cpu().numpy(), z["user_weights"], "user_weights")
    _close(eng.item.w_hop0.cpu().numpy(), z["item_weights"], "item_weights")
    _close(eng.scalars.cpu().numpy()[:3], [z["log_loss"], z["memory_loss"], z["cross_entropy"]], "log_loss, memory_loss, cross_entropy")
    ref = {k: z["grad:" + k] for k in p}
    ref[G.TABLE] = G.dense_rows(z, "grad", us.V, us.E)
    got = eng.named_grads(); got[G.TABLE] = eng.dtable.cpu().numpy()
    _grad_close(got, ref)
    eng.apply_gradients(c["lr"])
    step(1)
    eng.apply_gradients(c["lr"])
    torch.cuda.synchronize()
    now = eng.named_parameters()
    for k in p:
        _update_close(now[k], z["after:" + k], c["lr"], k)
    _update_close(eng.table.cpu().numpy(), G.dense_rows(z, "after", us.V, us.E, base=tb), c["lr"], G.TABLE)
    eng.close()


def test_model_class_with_the_reference_constructor_arguments_reproduces_the_reference_eval(tmp_path):
    """hpmn_b200.model.Hpmn built with the argument list the fixture's reference model was built with (hpmn.py:577-596),
    weights loaded by TF name, evaluated through predict_on_batch() -- the `sess.run([memory_loss, prediction])` of
    hpmn.py:509-511."""
    import torch
    from hpmn_b200.model import Hpmn
    z, c = G.load("refgraph_amazon")
    model = Hpmn(str(tmp_path), None, None, c["V"], c["user_dim"], c["item_dim"], c["user_maxlen"], c["item_maxlen"], c["lr"], c["H"],
                 c["E"], c["hop"], c["user_layers"], c["item_layers"], c["user_num_layers"], c["item_num_layers"], c["user"], c["item"],
                 None, c["l2_reg"], c["memory_reg"], max_batch=8)
    model.engine.load_named({k: v.astype(np.float32) for k, v in G.trainables(z).items()})
    model.engine.table.copy_(torch.as_tensor(z["var:" + G.TABLE]))
    B = c["B"]
    data = (z["label0"].tolist(), z["user_inp0"], [c["user_maxlen"]] * B, z["item_inp0"], [c["item_maxlen"]] * B)
    mem, pred, w0 = model.predict_on_batch(data)
    _close(pred, z["prediction"], "prediction")
    _close(w0, z["user_weights"], "user_weights")
    _close([mem], [z["memory_loss"]], "memory_loss")
