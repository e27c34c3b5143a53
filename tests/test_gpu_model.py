"""-m gpu: the reference-facing classes (hpmn_b200/model.py mirrors /root/reference/code/hpmn.py:16-560) driven end to end on
the CUDA engine: train loop with clip + Adam, eval with AUC / log-loss, result.log format, checkpoint round trip,
get_weights dumps, the XLong TSV loader path, and the `python hpmn.py DATASET` entry point."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _learnable_dataset(n, T, F, V, seed):
    """Front-padded tuples (util.py:152-159) whose label depends only on the target item id (40 distinct items), so a few
    dozen Adam steps must lower the loss."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ln = int(rng.integers(3, T + 1))
        uid = int(rng.integers(41, V))
        seq = [[uid, int(rng.integers(1, 41)), int(rng.integers(1, 20))] for _ in range(ln)]
        label = int(seq[-1][1] <= 20)
        item_part = [[0] * F for _ in range(T - ln)] + [s[:F] for s in seq]
        out.append((label, item_part, ln, [[0, 0]] * 4, 0))
    return out


def test_hpmn_class_train_eval_checkpoint(tmp_path):
    from hpmn_b200.model import Hpmn
    V = 300
    train = _learnable_dataset(512, 20, 3, V, 1)
    test = _learnable_dataset(128, 20, 3, V, 2)
    path = str(tmp_path / "amazon" / "hpmn") + "/"
    m = Hpmn(path, train, test, V, 3, 2, 20, 4, 0.01, 32, 16, 3, [2, 2, 5, 5, 1], [2, 2, 5, 5, 1], 3, 3, True, False,
             l2_reg=0., memory_reg=1e-5, max_batch=256)
    m.eval_every = 8                                    # the reference evaluates every 100 steps (hpmn.py:483)
    auc0, loss0, mem0 = m.eval(test, 64)
    best = m.train(6, 64)                               # 6 epochs x 8 steps
    auc1, loss1, mem1 = m.eval(train, 256)
    assert np.isfinite([auc0, loss0, mem0, auc1, loss1, mem1]).all()
    assert loss1 < loss0 - 0.02 and auc1 > 0.65, (loss0, loss1, auc1)     # it learns the rule on the target id
    assert best > 0.5
    # result.log: step \t 6 x %.5f (hpmn.py:100-103)
    lines = open(path + "result.log").read().strip().split("\n")
    assert len(lines) >= 1 and all(len(l.split("\t")) == 7 for l in lines)
    assert all(len(f.split(".")[1]) == 5 for f in lines[0].split("\t")[1:])
    # checkpoint round trip (hpmn.py:91-92, 105-111)
    m.save_model()
    assert os.path.exists(path + "ckpt/model.ckpt")
    before = m.eval(test, 64)
    m.engine.params.zero_()
    m.load_model()
    after = m.eval(test, 64)
    np.testing.assert_allclose(before, after, rtol=1e-6)
    # get_weights dumps (hpmn.py:521-560)
    m.get_weights()
    w = np.load(path + "weights.npy"); ln = np.load(path + "lengths.npy"); lb = np.load(path + "labels.npy")
    assert w.shape == (640, 3) and ln.shape == (640,) and lb.shape == (640,)
    np.testing.assert_allclose(w.sum(axis=1), 1.0, atol=1e-5)       # softmax weights of hop 0 (hpmn.py:182)


def test_hpmn_industry_xlong_loader_path(tmp_path):
    """Hpmn_Industry: TSV loader (data_loader.py:7-107), no id-0 mask, 23 zero steps in front, target = step -2."""
    from hpmn_b200.data_loader import write_synthetic_xlong
    from hpmn_b200.model import Hpmn_Industry
    tr, te = str(tmp_path / "train.txt"), str(tmp_path / "test.txt")
    write_synthetic_xlong(tr, 24, seed=1, n_items=5000, n_users=50)
    write_synthetic_xlong(te, 8, seed=2, n_items=5000, n_users=50)
    feature_size = 3269017 + 50 + 1                      # uid column is offset by 3269017 (data_loader.py:49)
    m = Hpmn_Industry(str(tmp_path / "xlong") + "/", tr, te, feature_size, 2, 1, 1001, 184, 0.001, 32, 16, 3, [2] * 10 + [1],
                      [3, 2, 2, 2, 2, 2, 2, 1], 5, 8, True, False, None, l2_reg=0, memory_reg=5e-5, max_batch=64)
    assert (m.shape.Tpad, m.shape.front_pad, m.shape.last_offset, m.shape.mask_id0) == (1024, 23, 2, False)
    assert m.shape.steps() == [1024, 512, 256, 128, 64]
    m.eval_every = 2
    best = m.train(epochs=1, batchsize=16)               # 8 lines -> 16 samples per batch, 3 steps
    auc, loss, mem = m.eval(te, 16)
    assert np.isfinite([best, auc, loss, mem]).all() and 0.0 < loss < 2.0
    m.get_weights()
    assert np.load(str(tmp_path / "xlong") + "/weights_new.npy").shape == (64, 5)
    assert np.load(str(tmp_path / "xlong") + "/ids.npy").shape == (64, 1001)


def test_cli_entry_point(tmp_path):
    """`python hpmn.py amazon` (hpmn.py:563-596) on synthetic tuples; usage message like the reference."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "hpmn.py"), "amazon", "--synthetic", "256", "--epochs", "1",
                        "--out", str(tmp_path)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "best test AUC" in r.stdout
    assert os.path.exists(str(tmp_path / "amazon" / "hpmn" / "ckpt" / "model.ckpt"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "hpmn.py")], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 1 and "Useage: python hpmn.py [dataset]" in r.stdout
