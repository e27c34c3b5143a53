"""Loader for tests/golden/refgraph_*.npz -- what the reference's own, unmodified code/hpmn.py computes when its graph is
built on the TF1-API stand-in (tests/golden/make_reference_graph_fixture.py, tests/golden/tf1_shim.py)."""
from __future__ import annotations

import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TABLE = "Embedding/emb_mtx"
NAMES = ["refgraph_amazon", "refgraph_xlong", "refgraph_dual", "refgraph_h64"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, json.loads(str(z["cfg"]))


def side_kwargs(c, B=None):
    """(user, item) shape keyword dicts (same field names in oracle.OracleShape and hpmn_b200.layout.HpmnShape) for the
    class the fixture was made with: Hpmn (hpmn.py:432-450) or Hpmn_Industry (hpmn.py:284-305)."""
    industry = c["cls"] == "Hpmn_Industry"
    common = dict(B=B or c["B"], E=c["E"], H=c["H"], hops=c["hop"], V=c["V"], mask_id0=not industry)
    user = dict(common, T=c["user_maxlen"], F=c["user_dim"], periods=list(c["user_layers"]), L=c["user_num_layers"],
                front_pad=23 if industry else 0, last_offset=2 if industry else 1, scope="User")
    item = dict(common, T=c["item_maxlen"], F=c["item_dim"], periods=list(c["item_layers"]), L=c["item_num_layers"],
                front_pad=8 if industry else 0, last_offset=1, scope="Item" if industry else "item")
    return user, item


def trainables(z):
    """name -> initial value (float64) of every trainable the fixture keeps, the table excluded"""
    return {str(k): z["var:" + str(k)].astype(np.float64) for k in z["trainable"]
            if str(k) != TABLE and "var:" + str(k) in z.files}


def dense_rows(z, prefix, V, E, base=None):
    out = np.zeros((V, E)) if base is None else np.array(base, dtype=np.float64)
    out[z[prefix + "_rows"]] = z[prefix + "_vals"]
    return out
