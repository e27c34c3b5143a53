"""CPU tests of the oracle: the NumPy restatement against (a) torch autograd of an independent
restatement, (b) finite differences, (c) the committed golden vectors, (d) the TF1.4 / reference
semantics it claims to follow (each test names the reference lines)."""
import os
import pickle

import numpy as np
import pytest

from oracle import hpmn_oracle as O
from oracle import make_golden
from oracle import tf1_restatement as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mk(**kw):
    base = dict(B=5, T=12, F=2, E=4, H=8, periods=[2, 3], L=3, hops=2, V=60)
    base.update(kw)
    return O.OracleShape(**base)


@pytest.mark.parametrize("kw", [
    dict(),
    dict(F=3, periods=[2, 2], T=16),
    dict(front_pad=4, mask_id0=False, last_offset=2, periods=[2, 2, 2], L=4, T=12, hops=3),
    dict(L=1, periods=[], T=7),
])
def test_adjoint_matches_autograd(kw):
    sh = _mk(**kw)
    p, tb = O.init_params(sh, mode="stress", dtype=np.float64)
    ids, labels = O.synthetic_batch(sh, ragged=sh.mask_id0)
    mreg = 1e-2 if sh.L > 1 else 0.0
    f = O.forward(sh, p, tb, ids, labels, memory_reg=mreg)
    g, dtb = O.backward(sh, f, ids, labels, memory_reg=mreg)     # L == 1: NaN on both sides, like TF
    res, g2, dtb2 = R.forward_backward_numpy(sh, p, tb, ids, labels, memory_reg=mreg)
    assert abs(f["loss"] - res["loss"]) < 1e-12
    np.testing.assert_allclose(f["pred"], res["pred"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(f["memory"], res["memory"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(f["w_hop0"], res["w_hop0"], rtol=0, atol=1e-13)
    for k in g:
        np.testing.assert_allclose(g[k], g2[k], rtol=1e-9, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(dtb, dtb2, rtol=1e-9, atol=1e-12)


def test_adjoint_matches_finite_differences():
    sh = _mk(B=3, T=8, periods=[2, 2])
    p, tb = O.init_params(sh, mode="stress", dtype=np.float64)
    ids, labels = O.synthetic_batch(sh, ragged=True)
    f = O.forward(sh, p, tb, ids, labels, memory_reg=1e-2)
    g, dtb = O.backward(sh, f, ids, labels, memory_reg=1e-2)
    rng = np.random.default_rng(0)
    eps = 1e-6
    for name in ["User/GRU0/rnn/gru_cell/gates/kernel", "User/GRU2/rnn/gru_cell/candidate/bias", "User/map",
                 "User/dense_4/kernel", "output/bn1/gamma", "output/fc2/kernel"]:
        for _ in range(3):
            idx = tuple(rng.integers(0, s) for s in p[name].shape)
            q = {k: v.copy() for k, v in p.items()}
            q[name][idx] += eps
            lp = O.forward(sh, q, tb, ids, labels, memory_reg=1e-2)["loss"]
            q[name][idx] -= 2 * eps
            lm = O.forward(sh, q, tb, ids, labels, memory_reg=1e-2)["loss"]
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - g[name][idx]) < 1e-7 + 1e-5 * abs(fd), (name, idx, fd, g[name][idx])
    row = int(ids[0, -1, 1])
    t2 = tb.copy(); t2[row, 1] += eps
    lp = O.forward(sh, p, t2, ids, labels, memory_reg=1e-2)["loss"]
    t2[row, 1] -= 2 * eps
    lm = O.forward(sh, p, t2, ids, labels, memory_reg=1e-2)["loss"]
    assert abs((lp - lm) / (2 * eps) - dtb[row, 1]) < 1e-7


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_golden_vectors_reproduce(name):
    """The committed fixtures are regenerated bit-for-bit (up to BLAS summation order) by the oracle."""
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    sh, mreg, mode, ragged, out = make_golden.compute(name)
    for k in gold.files:
        np.testing.assert_allclose(out[k], gold[k], rtol=1e-10, atol=1e-13, err_msg=k)


def test_fp32_oracle_close_to_fp64():
    """Round-off budget: the same restatement in fp32 stays inside the 1e-4 parity tolerance."""
    sh, mreg, mode, ragged, out = make_golden.compute("xlong_industry_small")
    params, table = O.init_params(sh, seed=4321, mode=mode, dtype=np.float32)
    ids, labels = O.synthetic_batch(sh, seed=1234, ragged=ragged)
    f32 = O.forward(sh, params, table, ids, labels, memory_reg=mreg, dtype=np.float32)
    np.testing.assert_allclose(f32["pred"], out["pred"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(f32["logit"], out["logit"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(f32["memory"], out["memory"], rtol=1e-4, atol=1e-6)


# ---- reference / TF1.4 semantics ---------------------------------------------------------------

def test_gru_gate_order_and_reset_before_matmul():
    """util.py:95-109: split gives r first then u; the reset gate multiplies h BEFORE the candidate matmul."""
    rng = np.random.default_rng(1)
    H, D = 3, 2
    Wg, bg = rng.normal(size=(D + H, 2 * H)), rng.normal(size=2 * H)
    Wc, bc = rng.normal(size=(D + H, H)), rng.normal(size=H)
    x = rng.normal(size=(1, 2, D))
    hs, rs, us, cs = O.gru_layer_fwd(x, Wg, bg, Wc, bc)
    h0 = np.zeros(H)
    g = 1 / (1 + np.exp(-(np.concatenate([x[0, 0], h0]) @ Wg + bg)))
    r, u = g[:H], g[H:]
    c = np.tanh(np.concatenate([x[0, 0], r * h0]) @ Wc + bc)
    h1 = u * h0 + (1 - u) * c
    np.testing.assert_allclose(hs[0, 0], h1, atol=1e-14)
    g = 1 / (1 + np.exp(-(np.concatenate([x[0, 1], h1]) @ Wg + bg)))
    r, u = g[:H], g[H:]
    c = np.tanh(np.concatenate([x[0, 1], r * h1]) @ Wc + bc)          # reset inside the matmul operand
    np.testing.assert_allclose(hs[0, 1], u * h1 + (1 - u) * c, atol=1e-14)
    c_cudnn = np.tanh(x[0, 1] @ Wc[:D] + r * (h1 @ Wc[D:]) + bc)      # cuDNN / torch.nn.GRU variant differs
    assert np.abs(c_cudnn - c).max() > 1e-3


def test_default_init_gate_bias_one():
    """util.py:84-86: gate bias initialised to 1.0, every other bias 0, BN gamma 1."""
    sh = _mk()
    p, _ = O.init_params(sh, mode="tf_default")
    assert np.all(p["User/GRU0/rnn/gru_cell/gates/bias"] == 1.0)
    assert np.all(p["User/GRU0/rnn/gru_cell/candidate/bias"] == 0.0)
    assert np.all(p["output/bn1/gamma"] == 1.0) and np.all(p["output/fc1/bias"] == 0.0)
    lim = np.sqrt(6.0 / (sh.D + sh.H + 2 * sh.H))
    assert np.abs(p["User/GRU0/rnn/gru_cell/gates/kernel"]).max() <= lim


def test_zero_padding_still_drives_the_gru():
    """hpmn.py:119-120 passes no sequence_length, so rnn.py:767-768 runs every step: an all-zero (id 0)
    prefix changes the state through the candidate bias (h stays 0 only while that bias is still at its 0 init)."""
    sh = _mk(B=2, T=8, periods=[2, 2])
    p, tb = O.init_params(sh, mode="stress", dtype=np.float64)
    ids, labels = O.synthetic_batch(sh, ragged=False)
    ids2 = ids.copy(); ids2[:, :4] = 0
    x = O.embed(sh, tb, ids2)
    assert np.all(x[:, :4] == 0)
    mem, saved = O.build_memory_fwd(sh, {k: v for k, v in p.items()}, x)
    assert np.abs(saved[0][1][:, 3]).max() > 1e-3      # h after 4 zero steps is not zero


def test_periodic_subsample_and_wavefront_equivalence():
    """hpmn.py:124-128 keeps outputs p-1, 2p-1, ...; running the layers one after another equals the
    online formulation where layer k fires when (t+1) % prod(p[:k]) == 0 (srnn.py:725-748)."""
    sh = _mk(B=3, T=12, periods=[2, 3], L=3)
    p, tb = O.init_params(sh, mode="stress", dtype=np.float64)
    ids, _ = O.synthetic_batch(sh, ragged=False)
    x = O.embed(sh, tb, ids)
    mem, saved = O.build_memory_fwd(sh, p, x)
    np.testing.assert_array_equal(saved[1][0], saved[0][1][:, 1::2])
    np.testing.assert_array_equal(saved[2][0], saved[1][1][:, 2::3])
    H = sh.H
    h = [np.zeros((sh.B, H)) for _ in range(sh.L)]
    prod = [1, 2, 6]
    for t in range(sh.Tpad):
        inp = x[:, t]
        for k in range(sh.L):
            if (t + 1) % prod[k]:
                break
            base = "User/GRU%d/rnn/gru_cell/" % k
            g = 1 / (1 + np.exp(-(np.concatenate([inp, h[k]], 1) @ p[base + "gates/kernel"] + p[base + "gates/bias"])))
            r, u = g[:, :H], g[:, H:]
            c = np.tanh(np.concatenate([inp, r * h[k]], 1) @ p[base + "candidate/kernel"] + p[base + "candidate/bias"])
            h[k] = u * h[k] + (1 - u) * c
            inp = h[k]
    np.testing.assert_allclose(np.stack(h, 1), mem, atol=1e-14)


def test_covreg_is_sum_over_batch_of_offdiag_frobenius():
    """hpmn.py:161-170."""
    rng = np.random.default_rng(2)
    M = rng.normal(size=(4, 3, 8))
    val, _ = O.covreg_fwd(M)
    tot = 0.0
    for b in range(4):
        mc = M[b] - M[b].mean(axis=1, keepdims=True)
        Cm = mc @ mc.T / 8
        Cm = Cm - np.diag(np.diag(Cm))
        tot += np.sqrt((Cm ** 2).sum())
    assert abs(val - tot) < 1e-12


def test_covreg_gradient_is_nan_for_single_slot_like_tf():
    """tf.norm's gradient at an all-zero matrix is 0/0; with L == 1 the off-diagonal part is empty."""
    sh = _mk(L=1, periods=[], T=6)
    p, tb = O.init_params(sh, dtype=np.float64)
    ids, labels = O.synthetic_batch(sh)
    f = O.forward(sh, p, tb, ids, labels)
    g, _ = O.backward(sh, f, ids, labels)
    assert np.isnan(g["User/GRU0/rnn/gru_cell/gates/kernel"]).any()
    g2, _ = O.backward(sh, f, ids, labels, guard_zero_norm=True)
    assert np.isfinite(g2["User/GRU0/rnn/gru_cell/gates/kernel"]).all()


def test_head_bn_inference_and_logloss():
    """hpmn.py:190: BN never in training mode -> gamma*x/sqrt(1+1e-3)+beta; hpmn.py:202: log_loss eps 1e-7, mean."""
    sh = _mk()
    p, _ = O.init_params(sh, mode="stress", dtype=np.float64)
    rng = np.random.default_rng(3)
    repre = rng.normal(size=(sh.B, sh.H + sh.D))
    labels = np.array([0, 1, 1, 0, 1], dtype=np.int32)
    logit, pred, ll, saved = O.head_fwd(p, repre, labels)
    np.testing.assert_allclose(saved[0], repre / np.sqrt(1.001) * p["output/bn1/gamma"] + p["output/bn1/beta"], atol=1e-14)
    ref = np.mean(-labels * np.log(pred + 1e-7) - (1 - labels) * np.log(1 - pred + 1e-7))
    assert abs(ll - ref) < 1e-14
    # dropout scales kept units by 1/keep_prob (tf.nn.dropout)
    m1 = (rng.random((sh.B, 200)) < 0.5).astype(np.float64); m2 = (rng.random((sh.B, 80)) < 0.5).astype(np.float64)
    _, _, _, sv = O.head_fwd(p, repre, labels, keep_prob=0.5, masks=(m1, m2))
    np.testing.assert_allclose(sv[3], sv[2] * m1 * 2.0, atol=1e-14)


def test_embedding_mask_and_industry_variant():
    """hpmn.py:417-423 (id 0 -> zeros) vs hpmn.py:266-282 + 288-292 (no mask, 23 zero steps in front, last = -2)."""
    sh = _mk(B=2, T=6, periods=[2, 2])
    _, tb = O.init_params(sh, mode="stress", dtype=np.float64)
    ids = np.array([[[0, 5], [3, 0], [1, 2], [4, 4], [0, 0], [7, 8]]] * 2, dtype=np.int32)
    x = O.embed(sh, tb, ids)
    assert np.all(x[0, 0, :4] == 0) and np.all(x[0, 0, 4:] == tb[5]) and np.all(x[0, 4] == 0)
    shi = _mk(B=2, T=6, periods=[2, 2], front_pad=2, mask_id0=False, last_offset=2)
    xi = O.embed(shi, tb, ids)
    assert xi.shape == (2, 8, 8) and np.all(xi[:, :2] == 0) and np.all(xi[0, 2, :4] == tb[0])


def test_clip_adam_matches_tf_formula():
    """hpmn.py:209-214 [TF1.4 AdamOptimizer]: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); eps outside the sqrt."""
    rng = np.random.default_rng(4)
    var = rng.normal(size=10); grad = rng.normal(size=10) * 3
    m = np.zeros(10); v = np.zeros(10)
    v0 = var.copy()
    O.clip_adam_step(var, grad, m, v, t=1, lr=0.01)
    g = np.clip(grad, -1, 1)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    np.testing.assert_allclose(var, v0 - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8), atol=1e-15)


@pytest.mark.skipif(not os.path.exists("/root/reference/data/amazon/dataset_hpmn.pkl"), reason="reference data not mounted")
def test_amazon_sample_fixture_layout():
    """data/amazon/dataset_hpmn.pkl: (label, item_part[100][3], len, user_part[100][2], len), front padded with
    id 0 (util.py:152-159), target last, feature_size 256205."""
    with open("/root/reference/data/amazon/dataset_hpmn.pkl", "rb") as f:
        train = pickle.load(f, encoding="latin1"); test = pickle.load(f, encoding="latin1"); fs = pickle.load(f, encoding="latin1")
    assert (len(train), len(test), fs) == (2048, 1024, 256205)
    lab, ip, il, up, ul = train[0]
    ip = np.array(ip)
    assert ip.shape == (100, 3) and np.array(up).shape == (100, 2) and lab in (0, 1)
    assert np.all(ip[: 100 - il] == 0) and np.all(ip[100 - il:, 0] == ip[-1, 0]) and ip.max() < fs
