"""World-size-2 gloo test (CPU) of the data-parallel host logic in hpmn_b200/dist.py: row sharding, the
loss_batch convention and the single flat all-reduce must reproduce the single-process gradient.  The
per-rank compute is the oracle here (no GPU in this container); on the B200 box the same helpers run over
NCCL with the CUDA engine (tests/test_gpu_parity.py::test_two_gpu_gradient_equals_single)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hpmn_b200 import dist as hd
from oracle import hpmn_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _shape(B):
    return O.OracleShape(B=B, T=12, F=2, E=4, H=8, periods=[2, 3], L=3, hops=2, V=40)


def _flat(g, dtable):
    return np.concatenate([g[k].reshape(-1) for k in g] + [dtable.reshape(-1)])


def _worker(rank, world, port, B, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    r, lr, w = hd.init_process_group("gloo")
    assert (r, w) == (rank, world)
    sh = _shape(B)
    p, tb = O.init_params(sh, mode="stress", dtype=np.float64)
    ids, labels = O.synthetic_batch(sh, ragged=True)
    lo, hi = hd.shard_range(B, rank, world)
    shl = _shape(hi - lo)
    f = O.forward(shl, p, tb, ids[lo:hi], labels[lo:hi], memory_reg=1e-2)
    g, dtb = O.backward(shl, f, ids[lo:hi], labels[lo:hi], memory_reg=1e-2, loss_scale_B=B)
    flat = torch.from_numpy(_flat(g, dtb))
    hd.allreduce_flat(flat)

    class _Eng:                   # stand-in for HpmnEngine without a comm stream: allreduce_grads == the flat all-reduce
        pass
    eng = _Eng(); eng.flat_grad = torch.from_numpy(_flat(g, dtb)).clone(); eng.comm_stream = None
    hd.allreduce_grads(eng)
    assert torch.equal(eng.flat_grad, flat)
    # GradExchange on an engine that is not in symmetric memory (CPU / gloo here) must fall back to the flat all-reduce
    eng2 = _Eng(); eng2.flat_grad = torch.from_numpy(_flat(g, dtb)).clone(); eng2.comm_stream = None
    eng2.symmetric = False; eng2.device = torch.device("cpu")
    ex = hd.GradExchange(eng2, mode="auto").attach()
    assert ex.mode == "nccl" and (ex.rank, ex.world) == (rank, world)
    hd.exchange_grads(eng2)
    assert torch.equal(eng2.flat_grad, flat)
    scal = torch.tensor([f["logloss"] * (hi - lo) / B, f["covreg"]], dtype=torch.float64)
    hd.allreduce_scalars(scal)
    hd.barrier()
    if rank == 0:
        q.put((flat.numpy(), scal.numpy()))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_allreduce_equals_single_process():
    B, world = 7, 2          # odd batch: shards of 4 and 3 rows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat, scal = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sh = _shape(B)
    p, tb = O.init_params(sh, mode="stress", dtype=np.float64)
    ids, labels = O.synthetic_batch(sh, ragged=True)
    f = O.forward(sh, p, tb, ids, labels, memory_reg=1e-2)
    g, dtb = O.backward(sh, f, ids, labels, memory_reg=1e-2)
    ref = _flat(g, dtb)
    assert np.linalg.norm(flat - ref) / np.linalg.norm(ref) < 1e-12      # <= 1e-6 relative
    assert abs(scal[0] - f["logloss"]) < 1e-12 and abs(scal[1] - f["covreg"]) < 1e-12


def test_shard_range_covers_batch():
    for n, w in [(256, 8), (7, 2), (5, 8), (1000, 3)]:
        spans = [hd.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
