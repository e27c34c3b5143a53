"""Time the update step (hpmn_clip_adam: clip_by_value + dense Adam over the flat [params | table] buffer, hpmn.py:209-214)
at the XLong size, and the full train step fwd + bwd + update.  python -m tools.adam_bench"""
import numpy as np
import torch

import bench
from hpmn_b200.data_loader import synthetic_ids
from hpmn_b200.engine import HpmnEngine
from hpmn_b200.layout import HpmnShape


def main():
    cfg = bench.CONFIGS["xlong"]
    B = cfg["batch"]
    sh = HpmnShape(B=B, T=cfg["T"], F=cfg["F"], E=cfg["E"], H=cfg["H"], periods=cfg["periods"], L=cfg["L"], hops=cfg["hops"],
                   V=cfg["V"], front_pad=cfg["front_pad"], mask_id0=cfg["mask_id0"], last_offset=cfg["last_offset"])
    eng = HpmnEngine(sh, device=0, memory_reg=cfg["memory_reg"], seed=4321)
    dev = eng.device
    ids = [torch.from_numpy(synthetic_ids(B, sh.T, sh.F, sh.V, seed=1234 + i)).to(dev) for i in range(4)]
    lab = [torch.from_numpy(np.random.default_rng(9 + i).integers(0, 2, size=B).astype(np.int32)).to(dev) for i in range(4)]

    def timed(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    eng.forward_backward(ids[0], lab[0], keep_prob=0.5, seed=1)
    t_adam = timed(lambda i: eng.apply_gradients(0.001), 20)
    n_bytes = 7 * 4 * (eng.n_params + sh.V * sh.E)

    def train(i):
        eng.forward_backward(ids[i % 4], lab[i % 4], keep_prob=0.5, seed=i)
        eng.apply_gradients(0.001)
    t_train = timed(train, 20)
    t_fb = timed(lambda i: eng.forward_backward(ids[i % 4], lab[i % 4], keep_prob=0.5, seed=i), 20)
    print("clip+Adam: %.3f ms (%.0f GB/s over 7 x %.0f MB); fwd+bwd %.3f ms; fwd+bwd+update %.3f ms = %.0f samples/s"
          % (t_adam, n_bytes / t_adam / 1e6, n_bytes / 7e6, t_fb, t_train, B / t_train * 1e3))


if __name__ == "__main__":
    main()
