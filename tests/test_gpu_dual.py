"""-m gpu: `item=True` -- both memory sides + the wider head (hpmn_b200/dual.py over the K1-K5 entry points and
hpmn_head_wide_*) against the fp64 autograd restatement of hpmn.py:432-465 (oracle/tf1_restatement.forward_torch_dual)."""
import numpy as np
import pytest

from hpmn_b200.layout import HpmnShape, param_names
from oracle import hpmn_oracle as O
from oracle import tf1_restatement as R
from tests._parity import oracle_shape

pytestmark = pytest.mark.gpu


def _case(B=6, mode="stress", industry=False):
    if industry:    # Hpmn_Industry: no id-0 mask, front padding on both sides, user target = second to last step (hpmn.py:284-304)
        us = HpmnShape(B=B, T=29, F=2, E=16, H=32, periods=[2, 2, 2], L=4, hops=3, V=300, front_pad=3, mask_id0=False, last_offset=2)
        it = HpmnShape(B=B, T=10, F=1, E=16, H=32, periods=[3, 2], L=3, hops=3, V=300, front_pad=2, mask_id0=False, last_offset=1,
                       scope="Item")
    else:           # Hpmn amazon: user side F=3 periods [2,5], item side F=2 (hpmn.py:576-595)
        us = HpmnShape(B=B, T=20, F=3, E=16, H=32, periods=[2, 5], L=3, hops=3, V=300)
        it = HpmnShape(B=B, T=20, F=2, E=16, H=32, periods=[2, 2], L=3, hops=3, V=300, scope="item")
    rng = np.random.default_rng(5)
    params = {}
    for sh in (us, it):
        p, table = O.init_params(oracle_shape(sh), seed=4321 + len(params), mode=mode, dtype=np.float32)
        params.update({k: v for k, v in p.items() if not k.startswith("output/")})
    Rtot = (us.H + us.D) + (it.H + it.D)
    head_shapes = {"output/bn1/gamma": (Rtot,), "output/bn1/beta": (Rtot,), "output/fc1/kernel": (Rtot, 200), "output/fc1/bias": (200,),
                   "output/fc2/kernel": (200, 80), "output/fc2/bias": (80,), "output/fc3/kernel": (80, 1), "output/fc3/bias": (1,)}
    for k, shp in head_shapes.items():
        params[k] = (rng.standard_normal(shp) * (0.3 if k.endswith("kernel") else 0.1) + (1.0 if k.endswith("gamma") else 0.0)).astype(np.float32)
    ids_u, labels = O.synthetic_batch(oracle_shape(us), seed=11, ragged=us.mask_id0)
    ids_i, _ = O.synthetic_batch(oracle_shape(it), seed=12, ragged=it.mask_id0)
    return us, it, params, table, ids_u, ids_i, labels


@pytest.mark.parametrize("industry", [False, True], ids=["hpmn", "hpmn_industry"])
def test_dual_memory_forward_backward_matches_autograd(industry):
    import torch
    from hpmn_b200.dual import HpmnDualEngine
    us, it, params, table, ids_u, ids_i, labels = _case(industry=industry)
    mreg = 1e-3
    p64 = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in params.items()}
    tb = torch.tensor(table.astype(np.float64), requires_grad=True)
    out = R.forward_torch_dual(oracle_shape(us), oracle_shape(it), p64, tb, torch.tensor(ids_u, dtype=torch.int64),
                               torch.tensor(ids_i, dtype=torch.int64), torch.tensor(labels, dtype=torch.float64), memory_reg=mreg)
    out["loss"].backward()
    eng = HpmnDualEngine(us, it, device=0, memory_reg=mreg, table=table, params=params)
    dev = eng.device
    eng.forward_backward(torch.as_tensor(ids_u, device=dev), torch.as_tensor(ids_i, device=dev), torch.as_tensor(labels, device=dev))
    torch.cuda.synchronize()
    ref_logit = out["logit"].detach().numpy()
    got_logit = eng.logit.cpu().numpy()
    assert np.all(np.abs(got_logit - ref_logit) <= 1e-4 * np.abs(ref_logit) + 1e-6)
    s = eng.scalars.cpu().numpy()
    np.testing.assert_allclose(s[:3], [out[k].item() for k in ("logloss", "covreg", "loss")], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(eng.user.w_hop0.cpu().numpy(), out["user"]["w_hop0"].detach().numpy(), rtol=1e-4, atol=1e-6)
    got = eng.named_grads()
    assert set(got) == set(params)
    ref = {k: v.grad.numpy() for k, v in p64.items()}
    ref["Embedding/emb_mtx"] = tb.grad.numpy(); got["Embedding/emb_mtx"] = eng.dtable.cpu().numpy()
    gmax = max(np.abs(v).max() for v in ref.values())
    for k, v in ref.items():
        err = np.linalg.norm(got[k].astype(np.float64) - v)
        assert err <= 1e-3 * (np.linalg.norm(v) + 1e-4 * gmax), "%s: %.3e vs ||ref|| %.3e" % (k, err, np.linalg.norm(v))
    # a smaller batch reuses the buffers and gives the same rows
    eng.forward(torch.as_tensor(ids_u[:3], device=dev), torch.as_tensor(ids_i[:3], device=dev), torch.as_tensor(labels[:3], device=dev))
    torch.cuda.synchronize()
    assert np.allclose(eng.logit.cpu().numpy()[:3], got_logit[:3], rtol=1e-6, atol=1e-7)
    eng.close()


def test_model_classes_train_with_item_side(tmp_path):
    """Hpmn(user=True, item=True) and Hpmn(user=False, item=True): the constructor switches of hpmn.py:452-462."""
    from hpmn_b200.data_loader import synthetic_dataset
    from hpmn_b200.model import Hpmn
    train = synthetic_dataset(96, 20, 3, 500, user_T=20, user_F=2, seed=1)
    for i, t in enumerate(train):      # give the user_part real ids (synthetic_dataset leaves it zero)
        rng = np.random.default_rng(i)
        train[i] = (t[0], t[1], t[2], rng.integers(1, 500, size=(20, 2)).tolist(), 20)
    for user, item in ((True, True), (False, True)):
        m = Hpmn(str(tmp_path / ("m%d%d" % (user, item))), train, train[:32], 500, 3, 2, 20, 20, 0.003, 32, 16, 3, [2, 5, 1], [2, 2, 1],
                 3, 3, user, item, l2_reg=0., memory_reg=1e-5, max_batch=64)
        m.eval_every = 2
        best = m.train(2, 32)
        auc, loss, mem = m.eval(train[:32], 32)
        assert np.isfinite(loss) and 0.0 <= auc <= 1.0 and np.isfinite(mem) and 0.0 <= best <= 1.0
