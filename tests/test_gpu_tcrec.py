"""-m gpu: the tensor-core recurrence (hpmn_b200/csrc/tcrec.cu: tcgen05 MMAs fed by TMA, recurrent state in tensor
memory) against the fp64 oracle, through the K2 / K4 entry points of the C ABI (hpmn_memory_fwd / hpmn_memory_bwd).
The library selects this path for H = 64 and for large batches; HPMN_TCREC=1 forces it so that small cases test it too.

Tolerances: memory |d| <= 1e-4*|ref| + 1e-5 (GRU states are bounded by 1; the 3xTF32 products carry ~2^-21 per term and the
tensor core truncates when it accumulates, so an element that cancels to ~1e-5 keeps ~4e-6 of round-off after 1024 dependent
steps -- 1e-4 relative on such an element is not meaningful; the whole-path tests check the logits at 1e-4 relative);
gradients 1e-3 relative L2 per tensor."""
import ctypes as C

import numpy as np
import pytest

from hpmn_b200.layout import HpmnShape, param_layout
from oracle import hpmn_oracle as O
from tests._parity import oracle_shape

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def _flat_params(sh, params):
    lay, n = param_layout(sh)
    flat = np.zeros(n, np.float32)
    for name, (off, shp) in lay.items():
        flat[off: off + int(np.prod(shp))] = np.asarray(params[name], np.float32).reshape(-1)
    return flat, lay


def run_memory_fwd(sh, mode="stress", seed=4321, x_scale=0.5, force=True, monkeypatch=None):
    import torch
    from hpmn_b200 import _lib
    if monkeypatch is not None and force:
        monkeypatch.setenv("HPMN_TCREC", "1")
    lib = _lib.lib()
    osh = oracle_shape(sh)
    params, _ = O.init_params(osh, seed=seed, mode=mode, dtype=np.float32)
    rng = np.random.default_rng(seed + 1)
    x = (rng.standard_normal((sh.B, sh.Tpad, sh.D)) * x_scale).astype(np.float32)
    mem_ref, saved = O.build_memory_fwd(osh, {k: v.astype(np.float64) for k, v in params.items()}, x.astype(np.float64))
    flat, _ = _flat_params(sh, params)
    ctx = C.c_void_p(); _lib.check(lib.hpmn_create(C.byref(ctx), 0))
    c = sh.to_c()
    ws = torch.empty(lib.hpmn_workspace_bytes(C.byref(c), 1), dtype=torch.uint8, device="cuda")
    d_x = torch.as_tensor(x, device="cuda"); d_p = torch.as_tensor(flat, device="cuda")
    mem = torch.zeros(sh.B, sh.L, sh.H, device="cuda")
    l0 = lib.hpmn_launch_count(ctx)
    _lib.check(lib.hpmn_memory_fwd(ctx, C.byref(c), d_x.data_ptr(), d_p.data_ptr(), mem.data_ptr(), ws.data_ptr(), None), ctx)
    torch.cuda.synchronize()
    launches = lib.hpmn_launch_count(ctx) - l0
    lib.hpmn_destroy(ctx)
    return mem.cpu().numpy(), mem_ref, launches


FWD_CASES = {
    "H32_L1_T1": HpmnShape(B=5, T=1, F=2, E=16, H=32, periods=[], L=1, hops=1, V=10),
    "H32_L1_T2": HpmnShape(B=5, T=2, F=2, E=16, H=32, periods=[], L=1, hops=1, V=10),
    "H32_L1_T33_B130": HpmnShape(B=130, T=33, F=2, E=16, H=32, periods=[], L=1, hops=1, V=10),
    "H32_L3_p22_T64": HpmnShape(B=200, T=64, F=2, E=16, H=32, periods=[2, 2], L=3, hops=1, V=10),
    "H32_F3_D48_L3_p25": HpmnShape(B=70, T=100, F=3, E=16, H=32, periods=[2, 5], L=3, hops=1, V=10),
    "H32_F4_D64_L2": HpmnShape(B=9, T=12, F=4, E=16, H=32, periods=[3], L=2, hops=1, V=10),
    "H64_L1_T1": HpmnShape(B=7, T=1, F=2, E=16, H=64, periods=[], L=1, hops=1, V=10),
    "H64_L1_T2": HpmnShape(B=7, T=2, F=2, E=16, H=64, periods=[], L=1, hops=1, V=10),
    "H64_L1_T3": HpmnShape(B=7, T=3, F=2, E=16, H=64, periods=[], L=1, hops=1, V=10),
    "H64_L4_p222_T64": HpmnShape(B=131, T=64, F=2, E=16, H=64, periods=[2, 2, 2], L=4, hops=1, V=10),
    "H64_F3_L2_p3": HpmnShape(B=20, T=27, F=3, E=16, H=64, periods=[3], L=2, hops=1, V=10),
    "H32_xlong_T1024_L5": HpmnShape(B=140, T=1001, F=2, E=16, H=32, periods=[2, 2, 2, 2], L=5, hops=1, V=10, front_pad=23),
}


@pytest.mark.parametrize("name", sorted(FWD_CASES))
def test_tcrec_forward_matches_oracle(name, monkeypatch):
    sh = FWD_CASES[name]
    mem, ref, launches = run_memory_fwd(sh, monkeypatch=monkeypatch)
    assert launches >= sh.L + 2          # pack + split + one tcgen05 launch per layer
    bad = np.abs(mem - ref) - (RTOL * np.abs(ref) + ATOL)
    assert bad.max() <= 0, "memory: max |d| %.3e" % np.abs(mem - ref).max()


def _torch_memory(sh, params, x):
    """fp64 autograd restatement of build_memory (hpmn.py:113-131, util.py:81-110) for the K4 check."""
    import torch
    H = sh.H
    p = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=True) for k, v in params.items() if "/GRU" in k}
    xt = torch.tensor(x.astype(np.float64), requires_grad=True)
    inp, finals = xt, []
    for k in range(sh.L):
        base = "%s/GRU%d/rnn/gru_cell/" % (sh.scope, k)
        Wg, bg, Wc, bc = p[base + "gates/kernel"], p[base + "gates/bias"], p[base + "candidate/kernel"], p[base + "candidate/bias"]
        h = torch.zeros(x.shape[0], H, dtype=torch.float64)
        outs = []
        for xs in inp.unbind(1):
            g = torch.sigmoid(torch.cat([xs, h], 1) @ Wg + bg)
            r, u = g[:, :H], g[:, H:]
            c = torch.tanh(torch.cat([xs, r * h], 1) @ Wc + bc)
            h = u * h + (1 - u) * c
            outs.append(h)
        finals.append(h)
        if k < sh.L - 1:
            pk = sh.periods[k]
            inp = torch.stack(outs[pk - 1::pk], 1)
    return torch.stack(finals, 1), p, xt


def run_memory_bwd(sh, monkeypatch=None, mode="stress", seed=4321, x_scale=0.5):
    import torch
    from hpmn_b200 import _lib
    if monkeypatch is not None:
        monkeypatch.setenv("HPMN_TCREC", "1")
    lib = _lib.lib()
    osh = oracle_shape(sh)
    params, _ = O.init_params(osh, seed=seed, mode=mode, dtype=np.float32)
    rng = np.random.default_rng(seed + 1)
    x = (rng.standard_normal((sh.B, sh.Tpad, sh.D)) * x_scale).astype(np.float32)
    dmem = (rng.standard_normal((sh.B, sh.L, sh.H)) * 0.3).astype(np.float32)
    mem_t, p_t, x_t = _torch_memory(sh, params, x)
    (mem_t * torch.tensor(dmem.astype(np.float64))).sum().backward()
    flat, lay = _flat_params(sh, params)
    ctx = C.c_void_p(); _lib.check(lib.hpmn_create(C.byref(ctx), 0))
    c = sh.to_c()
    ws = torch.empty(lib.hpmn_workspace_bytes(C.byref(c), 1), dtype=torch.uint8, device="cuda")
    d_x = torch.as_tensor(x, device="cuda"); d_p = torch.as_tensor(flat, device="cuda")
    mem = torch.zeros(sh.B, sh.L, sh.H, device="cuda")
    d_dm = torch.as_tensor(dmem, device="cuda")
    dx = torch.full((sh.B, sh.Tpad, sh.D), 7.0, device="cuda")
    grads = torch.zeros(len(flat), device="cuda")
    _lib.check(lib.hpmn_memory_fwd(ctx, C.byref(c), d_x.data_ptr(), d_p.data_ptr(), mem.data_ptr(), ws.data_ptr(), None), ctx)
    _lib.check(lib.hpmn_memory_bwd(ctx, C.byref(c), d_x.data_ptr(), d_p.data_ptr(), d_dm.data_ptr(), dx.data_ptr(), grads.data_ptr(),
                                   ws.data_ptr(), None), ctx)
    torch.cuda.synchronize()
    lib.hpmn_destroy(ctx)
    g = grads.cpu().numpy()
    got = {name: g[off: off + int(np.prod(shp))].reshape(shp) for name, (off, shp) in lay.items() if name in p_t}
    ref = {name: v.grad.numpy() for name, v in p_t.items()}
    got["dx"] = dx.cpu().numpy(); ref["dx"] = x_t.grad.numpy()
    return got, ref


BWD_CASES = {
    "H32_L1_T1": HpmnShape(B=5, T=1, F=2, E=16, H=32, periods=[], L=1, hops=1, V=10),
    "H32_L1_T5": HpmnShape(B=5, T=5, F=2, E=16, H=32, periods=[], L=1, hops=1, V=10),
    "H32_L3_p22_T64_B200": HpmnShape(B=200, T=64, F=2, E=16, H=32, periods=[2, 2], L=3, hops=1, V=10),
    "H32_F3_L3_p25_T100": HpmnShape(B=70, T=100, F=3, E=16, H=32, periods=[2, 5], L=3, hops=1, V=10),
    "H32_F4_L2_p3": HpmnShape(B=9, T=12, F=4, E=16, H=32, periods=[3], L=2, hops=1, V=10),
    "H64_L1_T4": HpmnShape(B=7, T=4, F=2, E=16, H=64, periods=[], L=1, hops=1, V=10),
    "H64_L3_p22_T32_B131": HpmnShape(B=131, T=32, F=2, E=16, H=64, periods=[2, 2], L=3, hops=1, V=10),
    "H64_F3_L2_p3": HpmnShape(B=20, T=27, F=3, E=16, H=64, periods=[3], L=2, hops=1, V=10),
    "H32_T512_L4": HpmnShape(B=130, T=512, F=2, E=16, H=32, periods=[2, 2, 2], L=4, hops=1, V=10),
}


@pytest.mark.parametrize("name", sorted(BWD_CASES))
def test_tcrec_backward_matches_autograd(name, monkeypatch):
    got, ref = run_memory_bwd(BWD_CASES[name], monkeypatch)
    gmax = max(np.abs(v).max() for v in ref.values())
    for k, v in ref.items():
        err = np.linalg.norm(got[k].astype(np.float64) - v)
        assert err <= 1e-3 * (np.linalg.norm(v) + 1e-4 * gmax), "%s: %.3e vs ||ref|| %.3e" % (k, err, np.linalg.norm(v))
