"""Generate the golden fixtures under tests/golden/ from the fp64 oracle (python -m oracle.make_golden).

These vectors pin the CUDA path (and future edits of the oracle) to THIS restatement
of the reference graph, not to TensorFlow 1.4 -- the reference ships no outputs to compare with (the fixtures made from the
reference's own graph code are tests/golden/refgraph_*.npz, see tests/golden/make_reference_graph_fixture.py).
Inputs are regenerated from the seeds stored in each file; only outputs / gradients are stored."""
from __future__ import annotations

import os

import numpy as np

from . import hpmn_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "..", "tests", "golden")

CASES = {
    # name: (OracleShape kwargs, memory_reg, init mode, ragged)
    "amazon_ref_small": (dict(B=6, T=20, F=3, E=16, H=32, periods=[2, 5], L=3, hops=3, V=997), 1e-5, "tf_default", True),
    "amazon_synth_h18": (dict(B=6, T=20, F=2, E=16, H=18, periods=[2, 2], L=3, hops=3, V=997), 1e-3, "stress", True),
    "xlong_industry_small": (dict(B=4, T=41, F=2, E=16, H=32, periods=[2, 2, 2, 2], L=5, hops=3, V=499, front_pad=7,
                                  mask_id0=False, last_offset=2), 5e-5, "stress", False),
    "taobao_ref_small": (dict(B=4, T=36, F=4, E=16, H=32, periods=[2, 2, 3], L=4, hops=3, V=499), 1e-5, "stress", True),
}


def compute(name):
    kw, mreg, mode, ragged = CASES[name]
    sh = O.OracleShape(**kw)
    params, table = O.init_params(sh, seed=4321, mode=mode, dtype=np.float32)
    ids, labels = O.synthetic_batch(sh, seed=1234, ragged=ragged)
    fwd = O.forward(sh, params, table, ids, labels, memory_reg=mreg, dtype=np.float64)
    grads, dtable = O.backward(sh, fwd, ids, labels, memory_reg=mreg)
    out = dict(memory=fwd["memory"], pred=fwd["pred"], logit=fwd["logit"], w_hop0=fwd["w_hop0"],
               covreg=np.float64(fwd["covreg"]), logloss=np.float64(fwd["logloss"]), loss=np.float64(fwd["loss"]),
               dtable_rows=dtable[np.unique(ids)], ids_checksum=np.int64(ids.astype(np.int64).sum()))
    for k, v in grads.items():
        out["grad:" + k] = v
    return sh, mreg, mode, ragged, out


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name in CASES:
        sh, mreg, mode, ragged, out = compute(name)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "loss %.12f" % out["loss"], "pred[0] %.12f" % out["pred"][0])


if __name__ == "__main__":
    main()
