"""CPU tests of the host side: C-ABI library loads and exports every symbol the header declares, layouts
agree between Python and C, shapes are validated, loaders reproduce the reference's batching, and the product
fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest

from hpmn_b200 import _lib, layout
from hpmn_b200.data_loader import DataLoader, DataLoader_Mul, parse_xlong_lines, synthetic_dataset, write_synthetic_xlong
from oracle import hpmn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

XLONG = layout.HpmnShape(B=256, T=1001, F=2, E=16, H=32, periods=[2] * 10 + [1], L=5, hops=3, V=3308019, front_pad=23,
                         mask_id0=False, last_offset=2)
AMAZON = layout.HpmnShape(B=128, T=100, F=3, E=16, H=32, periods=[2, 2, 5, 5, 1], L=3, hops=3, V=256205)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hpmn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hpmn_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()                      # raises if the .so is missing or lacks a symbol
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.hpmn_abi_version() == _lib.HPMN_ABI_VERSION


def test_c_structs_match_header_sizes():
    assert C.sizeof(_lib.hpmn_shape) == 10 * 4 + 16 * 4 + 8
    assert C.sizeof(_lib.hpmn_outputs) == 5 * 8


@pytest.mark.parametrize("sh", [XLONG, AMAZON, layout.HpmnShape(B=4, T=20, F=2, E=16, H=18, periods=[2, 2], L=3, hops=2, V=100)])
def test_param_layout_python_equals_c(sh):
    lib = _lib.lib()
    c = sh.to_c()
    lay, total = layout.param_layout(sh)
    n = lib.hpmn_param_tensors(C.byref(c))
    assert n == len(lay)
    offs = (C.c_int64 * n)(); sizes = (C.c_int64 * n)()
    assert lib.hpmn_param_offsets(C.byref(c), offs, sizes, n) == n
    assert lib.hpmn_param_count(C.byref(c)) == total
    for i, (name, (off, shp)) in enumerate(lay.items()):
        assert offs[i] == off and sizes[i] == int(np.prod(shp)), name
        assert off % 4 == 0
    # and the oracle's independent naming agrees
    osh = O.OracleShape(B=sh.B, T=sh.T, F=sh.F, E=sh.E, H=sh.H, periods=list(sh.periods), L=sh.L, hops=sh.hops, V=sh.V)
    assert list(O.param_names(osh).items()) == [(k, v[1]) for k, v in lay.items()]


def test_workspace_and_shape_validation():
    lib = _lib.lib()
    assert lib.hpmn_workspace_bytes(C.byref(XLONG.to_c()), 1) > 500e6
    small = XLONG.with_batch(8)
    assert lib.hpmn_workspace_bytes(C.byref(small.to_c()), 1) < lib.hpmn_workspace_bytes(C.byref(XLONG.to_c()), 1)
    bad = layout.HpmnShape(B=4, T=21, F=2, E=16, H=32, periods=[2, 2], L=3, hops=2, V=100)     # 21 % 2 != 0
    assert lib.hpmn_workspace_bytes(C.byref(bad.to_c()), 1) == 0
    assert lib.hpmn_param_count(C.byref(bad.to_c())) < 0
    with pytest.raises(ValueError):
        bad.steps()
    wide = layout.HpmnShape(B=4, T=20, F=2, E=16, H=64, periods=[2, 2], L=3, hops=2, V=100)    # H = 64: tensor-core recurrence
    assert lib.hpmn_param_count(C.byref(wide.to_c())) == layout.param_layout(wide)[1]
    assert lib.hpmn_workspace_bytes(C.byref(wide.to_c()), 1) > 0
    odd = layout.HpmnShape(B=4, T=20, F=2, E=16, H=48, periods=[2, 2], L=3, hops=2, V=100)     # neither <= 32 nor 64
    assert lib.hpmn_param_count(C.byref(odd.to_c())) < 0
    assert XLONG.steps() == [1024, 512, 256, 128, 64] and AMAZON.steps() == [100, 50, 25]
    assert XLONG.gru_flops_fwd_per_sample() == 24379392      # 24.38 MFLOP


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.lib()
    ctx = C.c_void_p()
    rc = lib.hpmn_create(C.byref(ctx), 0)
    assert rc in (_lib.HPMN_ECUDA, _lib.HPMN_EARCH) and not ctx.value
    assert b"no CPU fallback" in lib.hpmn_last_error(None) or rc == _lib.HPMN_EARCH
    from hpmn_b200.engine import HpmnEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HpmnEngine(AMAZON.with_batch(2))


def test_product_does_not_import_oracle():
    import hpmn_b200
    pkg = os.path.dirname(hpmn_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M), fn
    assert "oracle" not in open(os.path.join(ROOT, "hpmn.py")).read()


def test_dataloader_matches_reference_batching():
    ds = synthetic_dataset(10, 6, 3, 50, user_T=4, user_F=2)
    batches = list(DataLoader(ds, 4))
    assert [b[0] for b in batches] == [1, 2, 3]                      # data_loader.py:297 returns the 1-based step
    assert [len(b[1][0]) for b in batches] == [4, 4, 2]              # last batch is short
    lab, item, ilen, user, ulen = batches[0][1]
    assert item.shape == (4, 6, 3) and user.shape == (4, 4, 2) and item.dtype == np.int32
    assert np.all(item[0, : 6 - ilen[0]] == 0) and np.all(item[0, 6 - ilen[0]:, 0] == item[0, -1, 0])


def test_xlong_parser_and_loader(tmp_path):
    line = "7\t12\t5,6,7\t8\t9\t1,2\t3,4\n"
    lab, item, ilen, user, ulen = parse_xlong_lines([line])
    assert lab == [1, 0] and item.shape == (2, 4, 2) and user.shape == (2, 2, 1)
    assert item[0, 0].tolist() == [12 + 3269017, 5] and item[0, -1, 1] == 8 and item[1, -1, 1] == 9   # data_loader.py:66-70
    p = tmp_path / "x.txt"
    write_synthetic_xlong(str(p), 5, hist=10, user_len=3)
    got = list(DataLoader_Mul(str(p), 4))                            # 2 lines per batch -> 3 batches, last one short
    assert [len(b[1][0]) for b in got] == [4, 4, 2]
    assert got[0][1][1].shape == (4, 11, 2)


def test_model_signatures_match_reference():
    from hpmn_b200 import model
    ref = ["path", "trainset", "testset", "feature_size", "user_dim", "item_dim", "user_maxlen", "item_maxlen",
           "learning_rate", "hidden_size", "embedding_size", "hop", "user_layers", "item_layers", "user_num_layers",
           "item_num_layers", "user", "item", "emb_initializer", "l2_reg", "memory_reg"]      # hpmn.py:218-239
    sig = list(inspect.signature(model.Hpmn_Industry.__init__).parameters)[1:]
    assert sig[: len(ref)] == ref
    for m in ("train", "eval", "save_model", "load_model", "get_weights", "log"):
        assert callable(getattr(model.Hpmn, m))
    assert (model.Hpmn.mask_id0, model.Hpmn.last_offset, model.Hpmn.eval_every) == (True, 1, 100)
    assert (model.Hpmn_Industry.mask_id0, model.Hpmn_Industry.last_offset, model.Hpmn_Industry.eval_every) == (False, 2, 10)


def test_xlong_loader_forwards_a_parse_error_and_can_be_closed(tmp_path):
    """A malformed line must fail the consumer (not hang it on an empty queue); close() releases a producer that is
    blocked on a full queue (train() leaving early on the early-stop rule)."""
    import time
    from hpmn_b200.data_loader import DataLoader_Mul, write_synthetic_xlong
    good = tmp_path / "good.txt"
    write_synthetic_xlong(str(good), 6, hist=4, user_len=3)
    bad = tmp_path / "bad.txt"
    bad.write_text(good.read_text() + "7\tnot-a-uid\t1,2\t3\t4\t5\t6\n")
    with pytest.raises(ValueError):
        for _ in DataLoader_Mul(str(bad), 4):
            pass
    ld = DataLoader_Mul(str(good), 2, max_q_size=1)     # 6 batches of one line, queue of 1: the producer blocks
    next(ld)
    ld.close()
    ld.thread.join(timeout=5)
    assert not ld.thread.is_alive()


def test_reference_arm_reports_the_product_arms_config():
    """bench.py: both arms must describe the same workload (the driver compares `config`)."""
    import bench
    cfg = bench.CONFIGS["xlong"]
    a = bench.make_config("xlong", cfg, 256, 1, 0.5)
    assert a["global_batch"] == 256 and "V=3308019" in a["workload"] and "T=1001->1024" in a["workload"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("\"config\": make_config(cfg_name, cfg, B,") == 2          # one call per arm
    assert "V_cap=" not in src.split("def run_reference")[1].split("def run_product")[0]
