"""The oracle against the reference's OWN graph code.

tests/golden/refgraph_*.npz hold what /root/reference/code/hpmn.py -- imported unmodified, its `tensorflow` resolved to
the TF1-API stand-in tests/golden/tf1_shim.py -- computes for three configurations: variables by TF name, feeds,
prediction / log_loss / memory_loss / cross_entropy / attention weights, compute_gradients() of every trainable, and
the variables after two train_step runs.  These tests pin the oracle's restatement of the graph WIRING (hpmn.py:113-214,
266-320, 414-465) to that; the arithmetic inside the TF ops is restated on both sides (see the shim's header)."""
import os

import numpy as np
import pytest

from oracle import hpmn_oracle as O
from oracle import tf1_restatement as R
from tests import _refgraph as G


@pytest.mark.parametrize("name", ["refgraph_amazon", "refgraph_xlong"])
def test_oracle_reproduces_the_reference_graph_single_side(name):
    z, c = G.load(name)
    sh = O.OracleShape(**G.side_kwargs(c)[0])
    assert list(O.param_names(sh)) == [k for k in map(str, z["trainable"]) if k != G.TABLE and k not in set(map(str, z["unused"]))]
    p = G.trainables(z)
    for k, shp in O.param_names(sh).items():
        assert p[k].shape == tuple(shp), k
    tb = z["var:" + G.TABLE].astype(np.float64)
    f = O.forward(sh, p, tb, z["user_inp0"], z["label0"], memory_reg=c["memory_reg"], dtype=np.float64)
    np.testing.assert_allclose(f["pred"], z["prediction"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(f["w_hop0"], z["user_weights"], rtol=0, atol=1e-13)
    np.testing.assert_allclose([f["logloss"], f["covreg"], f["loss"]], [z["log_loss"], z["memory_loss"], z["cross_entropy"]],
                               rtol=1e-12, atol=1e-15)
    g, dt = O.backward(sh, f, z["user_inp0"], z["label0"], memory_reg=c["memory_reg"])
    for k in p:
        np.testing.assert_allclose(g[k], z["grad:" + k], rtol=1e-8, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(dt, G.dense_rows(z, "grad", sh.V, sh.E), rtol=1e-8, atol=1e-13)

    # the two train_step runs: clip_by_value(-1, 1) + Adam on every trainable, the table densely (hpmn.py:209-214)
    var = dict(p); var[G.TABLE] = tb
    slots = {k: (np.zeros_like(v), np.zeros_like(v)) for k, v in var.items()}
    for t in (1, 2):
        if t == 2:
            f = O.forward(sh, {k: v for k, v in var.items() if k != G.TABLE}, var[G.TABLE], z["user_inp1"], z["label1"],
                          memory_reg=c["memory_reg"], dtype=np.float64)
            g, dt = O.backward(sh, f, z["user_inp1"], z["label1"], memory_reg=c["memory_reg"])
        g = dict(g); g[G.TABLE] = dt
        for k in var:
            O.clip_adam_step(var[k], g[k], slots[k][0], slots[k][1], t, lr=c["lr"])
    for k in p:        # `after:` is stored in float32: 6e-8 relative, the two updates are ~1e-3 each
        np.testing.assert_allclose(var[k], z["after:" + k], rtol=1e-7, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(var[G.TABLE], G.dense_rows(z, "after", sh.V, sh.E, base=tb), rtol=1e-7, atol=1e-9)
    f = O.forward(sh, {k: v for k, v in var.items() if k != G.TABLE}, var[G.TABLE], z["user_inp0"], z["label0"],
                  memory_reg=c["memory_reg"], dtype=np.float64)
    np.testing.assert_allclose(f["pred"], z["prediction_after"], rtol=0, atol=1e-12)


def test_oracle_reproduces_the_reference_graph_hidden_64():
    """hidden_size = 64 (the width that selects the tensor-core recurrence on the GPU); the fixture stores its gradients in
    float32, hence the looser gradient tolerance"""
    z, c = G.load("refgraph_h64")
    sh = O.OracleShape(**G.side_kwargs(c)[0])
    p = G.trainables(z)
    tb = z["var:" + G.TABLE].astype(np.float64)
    f = O.forward(sh, p, tb, z["user_inp0"], z["label0"], memory_reg=c["memory_reg"], dtype=np.float64)
    np.testing.assert_allclose(f["pred"], z["prediction"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(f["w_hop0"], z["user_weights"], rtol=0, atol=1e-13)
    np.testing.assert_allclose([f["logloss"], f["covreg"], f["loss"]], [z["log_loss"], z["memory_loss"], z["cross_entropy"]],
                               rtol=1e-12, atol=1e-15)
    g, dt = O.backward(sh, f, z["user_inp0"], z["label0"], memory_reg=c["memory_reg"])
    for k in p:
        np.testing.assert_allclose(g[k], z["grad:" + k], rtol=2e-6, atol=1e-10, err_msg=k)
    np.testing.assert_allclose(dt, G.dense_rows(z, "grad", sh.V, sh.E), rtol=2e-6, atol=1e-10)


def test_oracle_reproduces_the_reference_graph_both_sides_with_l2():
    """user=True, item=True, l2_reg != 0: concat of the two representations, memory_loss = imloss + umloss, l2 over
    every trainable including the table (hpmn.py:204-205, 452-456)"""
    import torch
    z, c = G.load("refgraph_dual")
    ku, ki = G.side_kwargs(c)
    us, it = O.OracleShape(**ku), O.OracleShape(**ki)
    p = {k: torch.tensor(v, requires_grad=True) for k, v in G.trainables(z).items()}
    tb = torch.tensor(z["var:" + G.TABLE].astype(np.float64), requires_grad=True)
    slots = {}
    lab = [torch.tensor(z["label%d" % i], dtype=torch.float64) for i in (0, 1)]
    uid = [torch.tensor(z["user_inp%d" % i], dtype=torch.int64) for i in (0, 1)]
    iid = [torch.tensor(z["item_inp%d" % i], dtype=torch.int64) for i in (0, 1)]

    def run(i):
        out = R.forward_torch_dual(us, it, p, tb, uid[i], iid[i], lab[i], memory_reg=c["memory_reg"])
        l2 = sum((v * v).sum() for v in p.values()) + (tb * tb).sum()
        out["cross_entropy"] = out["loss"] + c["l2_reg"] * 0.5 * l2
        return out

    out = run(0)
    np.testing.assert_allclose(out["pred"].detach().numpy(), z["prediction"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["user"]["w_hop0"].detach().numpy(), z["user_weights"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(out["item"]["w_hop0"].detach().numpy(), z["item_weights"], rtol=0, atol=1e-13)
    np.testing.assert_allclose([out[k].item() for k in ("logloss", "covreg", "cross_entropy")],
                               [z["log_loss"], z["memory_loss"], z["cross_entropy"]], rtol=1e-12)
    for t in (1, 2):
        if t == 2:
            out = run(1)
        every = dict(p); every[G.TABLE] = tb
        grads = torch.autograd.grad(out["cross_entropy"], list(every.values()))
        for (k, v), g in zip(every.items(), grads):
            g = g.numpy()
            if t == 1 and k == G.TABLE:
                np.testing.assert_allclose(g, G.dense_rows(z, "grad", us.V, us.E), rtol=1e-8, atol=1e-13)
            elif t == 1:
                np.testing.assert_allclose(g, z["grad:" + k], rtol=1e-8, atol=1e-13, err_msg=k)
            m, s = slots.setdefault(k, (np.zeros_like(g), np.zeros_like(g)))
            a = v.detach().numpy().copy()
            O.clip_adam_step(a, g, m, s, t, lr=c["lr"])
            with torch.no_grad():
                v.copy_(torch.from_numpy(a))
    for k, v in p.items():
        np.testing.assert_allclose(v.detach().numpy(), z["after:" + k], rtol=1e-7, atol=1e-9, err_msg=k)
    np.testing.assert_allclose(tb.detach().numpy(), G.dense_rows(z, "after", us.V, us.E, base=z["var:" + G.TABLE]),
                               rtol=1e-7, atol=1e-9)


@pytest.mark.skipif(not os.path.exists("/root/reference/code/hpmn.py"), reason="the reference is only present in the build container")
def test_fixtures_are_what_the_reference_code_computes_today():
    """Re-runs the reference's hpmn.py on the stand-in (one subprocess, all four cases) and compares with the committed fixtures."""
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(G.GOLD)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import make_reference_graph_fixture as M; "
            "M.HERE = sys.argv[1]; ref, tf = M.load_reference(); [M.make(n, ref, tf) for n in %r]" % (G.GOLD, G.NAMES))
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([sys.executable, "-c", code, tmp], check=True, cwd=root, stdout=subprocess.DEVNULL, timeout=600)
        for name in G.NAMES:
            new = np.load(os.path.join(tmp, name + ".npz"))
            old, _ = G.load(name)
            assert sorted(new.files) == sorted(old.files), name
            for k in old.files:
                if old[k].dtype.kind in "fiu":
                    np.testing.assert_allclose(new[k], old[k], rtol=1e-12, atol=1e-15, err_msg=name + ":" + k)


@pytest.mark.skipif(not os.path.exists("/root/reference/code/data_loader.py"), reason="the reference is only present in the build container")
def test_dataloader_equals_the_reference_loader_on_the_reference_dataset_sample():
    """code/data_loader.py runs under Python 3 as far as its plain `DataLoader` goes (DataLoader_Mul is Python-2 only:
    `batchsize / 2` lines, list-returning map, fork-shared file handle).  Batches of the reference's own sample tuples
    (tests/golden/amazon_hpmn_sample.pkl) through both loaders must be identical, short last batch included."""
    import importlib.util
    import pickle
    spec = importlib.util.spec_from_file_location("reference_data_loader", "/root/reference/code/data_loader.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from hpmn_b200.data_loader import DataLoader
    with open(os.path.join(G.GOLD, "amazon_hpmn_sample.pkl"), "rb") as f:
        train = pickle.load(f)
    train = train[:100]
    a, b = list(ref.DataLoader(train, 48)), list(DataLoader(train, 48))
    assert len(a) == len(b) == 3
    for (ia, da), (ib, db) in zip(a, b):
        assert ia == ib
        assert list(da[0]) == list(db[0]) and list(da[2]) == list(db[2]) and list(da[4]) == list(db[4])
        assert np.array_equal(np.asarray(da[1]), np.asarray(db[1])) and np.array_equal(np.asarray(da[3]), np.asarray(db[3]))


@pytest.mark.skipif(not os.path.exists("/root/reference/code/util.py"), reason="the reference is only present in the build container")
def test_gru_step_equals_the_reference_in_tree_cell():
    """code/util.py:56-110 (VecAttGRUCell) is the reference's in-tree copy of the TF1.4 GRUCell arithmetic: gates =
    sigmoid(_Linear([x, h])) split r | u, candidate on [x, r * h], bias of the gates initialised to 1, h' = u h + (1 - u) c
    (its one addition, `u = (1 - att_score) * u`, is the identity for att_score = 0).  Executed as it lies on the TF1-API
    stand-in -- only `_Linear` (concat, matmul, bias) is supplied -- it must agree with the cell the stand-in's dynamic_rnn
    runs and with the oracle's GRU layer, over several steps."""
    import importlib.util
    import pickle
    import subprocess
    import sys
    code = r'''
import sys, pickle, importlib.util
import numpy as np
sys.path.insert(0, %r)
import tf1_shim as tf
tf.install(sys.modules)
sys.modules["cPickle"] = pickle
spec = importlib.util.spec_from_file_location("reference_util", "/root/reference/code/util.py")
util = importlib.util.module_from_spec(spec); spec.loader.exec_module(util)
rng = np.random.default_rng(3)
B, D, H, T = 4, 6, 5, 7
x = rng.normal(size=(B, T, D)); 
cell = util.VecAttGRUCell(H)
xs = [tf.constant(x[:, t]) for t in range(T)]
h = tf.constant(np.zeros((B, H)))
outs = []
with tf.variable_scope("rnn"):
    with tf.variable_scope("gru_cell"):
        for t in range(T):
            _, h = cell(xs[t], h, 0.0)
            outs.append(h)
g = tf._g()
init = {k: v.numpy() for k, v in g.variables.items()}
for k, v in g.variables.items():      # away from the initial values (bias 1 / 0), keeping what the cell created
    v.value.data += __import__("torch").as_tensor(rng.normal(scale=0.3, size=tuple(v.value.shape)))
sess = tf.Session()
hs = np.stack(sess.run(outs), axis=1)
var = {k: v.numpy() for k, v in g.variables.items()}
pickle.dump({"x": x, "hs": hs, "var": var, "init": init}, open(sys.argv[1], "wb"))
''' % G.GOLD
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "cell.pkl")
        subprocess.run([sys.executable, "-c", code, out], check=True, timeout=300)
        with open(out, "rb") as f:
            r = pickle.load(f)
    var = r["var"]
    assert sorted(var) == ["rnn/gru_cell/candidate/bias", "rnn/gru_cell/candidate/kernel", "rnn/gru_cell/gates/bias",
                           "rnn/gru_cell/gates/kernel"]
    Wg, bg = var["rnn/gru_cell/gates/kernel"], var["rnn/gru_cell/gates/bias"]
    Wc, bc = var["rnn/gru_cell/candidate/kernel"], var["rnn/gru_cell/candidate/bias"]
    assert Wg.shape == (6 + 5, 10) and Wc.shape == (6 + 5, 5)
    assert np.all(r["init"]["rnn/gru_cell/gates/bias"] == 1.0) and np.all(r["init"]["rnn/gru_cell/candidate/bias"] == 0.0)   # util.py:84-86
    hs_oracle = O.gru_layer_fwd(r["x"], Wg, bg, Wc, bc)[0]
    np.testing.assert_allclose(hs_oracle, r["hs"], rtol=0, atol=1e-14)
    # and the cell the stand-in's dynamic_rnn uses for the fixtures
    sys.path.insert(0, G.GOLD)
    import torch
    import tf1_shim
    h = torch.zeros(4, 5, dtype=torch.float64)
    for t in range(7):
        h = tf1_shim.GRUCell.step(torch.as_tensor(r["x"][:, t]), h, *(torch.as_tensor(a) for a in (Wg, bg, Wc, bc)))
        np.testing.assert_allclose(h.numpy(), r["hs"][:, t], rtol=0, atol=1e-14)
