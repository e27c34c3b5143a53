"""Writes tests/golden/refgraph_*.npz: what the reference's OWN, unmodified model code computes.

/root/reference/code/hpmn.py is imported as it lies (it parses under Python 3) with `tensorflow` resolved to
tests/golden/tf1_shim.py -- a lazy-graph stand-in for the ~45 TF1.4 API functions that file uses, evaluated in float64 --
and `cPickle` to `pickle`.  The reference then builds its own graph (Hpmn / Hpmn_Industry.__init__ -> define_inputs ->
build_graph -> build_memory / query_memory / attention / get_covreg / build_fc_net) and this script drives it the way
its own train() / eval() do: `sess.run(fetches, feed_dict)` with keep_prob fed (1 here: the dropout mask of a TF session
is not reproducible anyway).  Recorded per case: every variable the graph created (by its TF name, initial value), the
feeds, prediction / log_loss / memory_loss / cross_entropy / attention weights, the gradient compute_gradients()
returns for every trainable variable, and every variable after two `train_step` runs (clip_by_value + Adam).

This pins the graph wiring of hpmn.py (what DESIGN.md section 2 calls authority 1) to the reference's own code.  The
arithmetic inside the TF ops is the shim's restatement of TF1.4 -- see the header of tf1_shim.py for what that does
and does not prove.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_reference_graph_fixture.py
"""
import importlib.util
import json
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/code"


def load_reference():
    sys.path.insert(0, HERE)
    import tf1_shim
    sys.modules["tensorflow"] = tf1_shim
    sys.modules["cPickle"] = pickle
    sys.path.insert(0, REF)                                  # the reference's `from data_loader import ...`
    spec = importlib.util.spec_from_file_location("reference_hpmn", os.path.join(REF, "hpmn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, tf1_shim


def ids_batch(rng, B, T, F, V, ragged, lo=1):
    """ids like the loaders produce them: front padding with id 0 for short rows (code/util.py:152-159), ids >= lo"""
    ids = rng.integers(lo, V, size=(B, T, F)).astype(np.int32)
    if ragged:
        for b in range(B):
            n = int(rng.integers(T // 3, T + 1)) if b else T
            ids[b, : T - n] = 0
    return ids


# name -> (class, kwargs of the reference constructor in the order of hpmn.py:218-239, batch, ragged)
CASES = {
    # code/hpmn.py:577-596, the amazon configuration (BASELINE.json configs[0]) on a small vocabulary
    "refgraph_amazon": dict(cls="Hpmn", V=300, user_dim=3, item_dim=2, user_maxlen=100, item_maxlen=100, lr=0.003, H=32, E=16,
                            hop=3, user_layers=[2, 2, 5, 5, 1], item_layers=[2, 2, 5, 5, 1], user_num_layers=3,
                            item_num_layers=3, user=True, item=False, l2_reg=0.0, memory_reg=1e-5, B=4, ragged=True,
                            emb_init=False),
    # code/hpmn.py:626-662, the XLong configuration (BASELINE.json configs[1]): 1001 steps + 23 pad, 5 layers of period 2
    "refgraph_xlong": dict(cls="Hpmn_Industry", V=600, user_dim=2, item_dim=1, user_maxlen=1001, item_maxlen=184, lr=0.001,
                           H=32, E=16, hop=3, user_layers=[2] * 10 + [1], item_layers=[3, 2, 2, 2, 2, 2, 2, 1],
                           user_num_layers=5, item_num_layers=8, user=True, item=False, l2_reg=0.0, memory_reg=5e-5, B=2,
                           ragged=False, emb_init=True),
    # both towers and an l2 term: the taobao shape of code/hpmn.py:604-623 with item=True
    "refgraph_dual": dict(cls="Hpmn", V=200, user_dim=4, item_dim=3, user_maxlen=300, item_maxlen=36, lr=0.001, H=32, E=16,
                          hop=3, user_layers=[2, 2, 3, 5, 5, 1], item_layers=[2, 2, 3, 3, 1], user_num_layers=4,
                          item_num_layers=5, user=True, item=True, l2_reg=1e-5, memory_reg=1e-5, B=3, ragged=True,
                          emb_init=False),
    # hidden_size is a free constructor argument (hpmn.py:218-239): 64 selects the tensor-core recurrence of the CUDA path.
    # compact: gradients stored in float32, no optimiser steps (clip + Adam are pinned by the cases above)
    "refgraph_h64": dict(cls="Hpmn", V=150, user_dim=2, item_dim=1, user_maxlen=24, item_maxlen=4, lr=0.001, H=64, E=16,
                         hop=3, user_layers=[2, 3, 1], item_layers=[2, 1], user_num_layers=3, item_num_layers=2, user=True,
                         item=False, l2_reg=0.0, memory_reg=1e-4, B=5, ragged=True, emb_init=False, compact=True),
}


def build(ref, tf, c, seed):
    tf.reset_default_graph()
    rng = np.random.default_rng(seed)
    emb = None
    if c["emb_init"]:                                        # hpmn.py:633-639: graph embeddings, then zero rows
        emb = np.concatenate([rng.normal(0, 0.3, size=(c["V"] - 100, c["E"])), np.zeros((100, c["E"]))], 0).astype(np.float32)
    path = tempfile.mkdtemp(prefix="refgraph_")
    model = getattr(ref, c["cls"])(path, None, None, c["V"], c["user_dim"], c["item_dim"], c["user_maxlen"], c["item_maxlen"],
                                   c["lr"], c["H"], c["E"], c["hop"], c["user_layers"], c["item_layers"],
                                   c["user_num_layers"], c["item_num_layers"], c["user"], c["item"], emb, c["l2_reg"],
                                   c["memory_reg"])
    return model, rng


def feed(model, batch, keep_prob):
    lab, uinp, iinp = batch
    B = len(lab)
    return {model.label: lab, model.user_inp: uinp, model.user_len: [uinp.shape[1]] * B, model.item_inp: iinp,
            model.item_len: [iinp.shape[1]] * B, model.keep_prob: keep_prob}


def sparse_rows(dense):
    rows = np.flatnonzero(np.any(dense != 0, axis=1))
    return rows.astype(np.int64), dense[rows]


def make(name, ref, tf):
    c = CASES[name]
    model, rng = build(ref, tf, c, seed=sum(map(ord, name)))
    lo = 0 if c["cls"] == "Hpmn_Industry" else 1
    batches = []
    for _ in range(2):
        lab = rng.integers(0, 2, size=c["B"]).astype(np.int32)
        uinp = ids_batch(rng, c["B"], c["user_maxlen"], c["user_dim"], c["V"], c["ragged"], lo)
        iinp = ids_batch(rng, c["B"], c["item_maxlen"], c["item_dim"], c["V"], c["ragged"], lo)
        batches.append((lab, uinp, iinp))
    g = model.graph
    out = {"cfg": np.array(json.dumps(c))}
    table_name = "Embedding/emb_mtx"
    for k, v in g.variables.items():
        out["var:" + k] = v.numpy().astype(np.float32)
        assert np.array_equal(out["var:" + k].astype(np.float64), v.numpy()), k      # initial values are fp32-exact
    initial = {k[4:]: v for k, v in out.items() if k.startswith("var:")}
    out["trainable"] = np.array([v.name for v in g.variables.values() if v.trainable])
    for i, (lab, uinp, iinp) in enumerate(batches):
        out["label%d" % i], out["user_inp%d" % i], out["item_inp%d" % i] = lab, uinp, iinp

    fd = feed(model, batches[0], 1.0)
    raw = model.optimizer.raw_gvs
    fetches = [model.prediction, model.log_loss, model.memory_loss, model.cross_entropy, model.user_weights,
               model.item_weights] + [gr for gr, _ in raw]
    vals = model.sess.run(fetches=fetches, feed_dict=fd)
    for k, v in zip(["prediction", "log_loss", "memory_loss", "cross_entropy", "user_weights", "item_weights"], vals[:6]):
        out[k] = np.asarray(v, np.float64)
    unused = set()                                           # the tower hpmn.py builds but does not feed to the head
    for (_, var), gval in zip(raw, vals[6:]):
        if var.name == table_name:
            out["grad_rows"], out["grad_vals"] = sparse_rows(gval)
        elif not c["item"] and var.name.lower().startswith("item/"):
            assert c["l2_reg"] == 0.0 and not np.any(gval), var.name      # connected only through 0 * l2_loss(v)
            unused.add(var.name)
        else:
            out["grad:" + var.name] = gval.astype(np.float32) if c.get("compact") else gval
    for k in unused:
        del out["var:" + k]
    out["unused"] = np.array(sorted(unused))
    if c.get("compact"):
        dst = os.path.join(HERE, name + ".npz")
        np.savez_compressed(dst, **out)
        print(name, os.path.getsize(dst), "bytes;", len(g.variables), "variables; pred", out["prediction"], "loss", out["log_loss"])
        return out
    # two optimiser steps, driven like Hpmn.train() (hpmn.py:473-483) but with keep_prob 1
    before = g.variables[table_name].numpy()
    for b in batches:
        model.sess.run(fetches=[model.train_step], feed_dict=feed(model, b, 1.0))
    for k, v in g.variables.items():
        if k == table_name:
            out["after_rows"], out["after_vals"] = sparse_rows(v.numpy() - before)
            out["after_vals"] = v.numpy()[out["after_rows"]].astype(np.float32)
        elif k in unused:
            assert np.array_equal(v.numpy().astype(np.float32), initial[k]), k    # zero gradient: Adam leaves it
        elif v.trainable:
            out["after:" + k] = v.numpy().astype(np.float32)
    out["prediction_after"] = np.asarray(model.sess.run(model.prediction, feed_dict=fd), np.float64)
    dst = os.path.join(HERE, name + ".npz")
    np.savez_compressed(dst, **out)
    print(name, os.path.getsize(dst), "bytes;", len(g.variables), "variables; pred", out["prediction"], "loss", out["log_loss"],
          "mem", out["memory_loss"])
    return out


def main():
    ref, tf = load_reference()
    for name in (sys.argv[1:] or sorted(CASES)):
        make(name, ref, tf)


if __name__ == "__main__":
    main()
