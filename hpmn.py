"""`python hpmn.py [DATASET]` -- the reference's entry point (/root/reference/code/hpmn.py:563-667) on the
B200-native engine.  DATASET is amazon | taobao | xlong with the reference's hard-coded hyper-parameters.

Extra, optional flags (the reference has none): --data-root DIR (default ../data, the reference's relative
layout), --synthetic N (train on N synthetic samples of the dataset's shape instead of reading files),
--epochs / --batchsize / --eval-every overrides, --out DIR for checkpoints and result.log."""
from __future__ import annotations

import argparse
import os
import random
import sys

import numpy as np

random.seed(42)   # hpmn.py:13


def main(argv=None):
    ap = argparse.ArgumentParser(usage="python hpmn.py [dataset]")
    ap.add_argument("dataset")
    ap.add_argument("--data-root", default=os.environ.get("HPMN_DATA", "../data"))
    ap.add_argument("--out", default="model")
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--epochs", type=int, default=0)
    ap.add_argument("--batchsize", type=int, default=0)
    ap.add_argument("--eval-every", type=int, default=0,
                    help="evaluate / log every N steps (reference: 100 for amazon/taobao, 10 for xlong -- on the 2048-tuple "
                         "sample file that is never reached, hpmn.py:483)")
    if argv is None and len(sys.argv) < 2:
        print("Useage: python hpmn.py [dataset]")   # sic, hpmn.py:565
        return 1
    a = ap.parse_args(argv)

    from hpmn_b200 import dist as hd
    from hpmn_b200.data_loader import load_hpmn_pickle, synthetic_dataset
    from hpmn_b200.model import Hpmn, Hpmn_Industry

    # `torchrun --nproc-per-node N hpmn.py DATASET`: one process per GPU, every batch sharded by rows, gradients exchanged once
    # per step (the reference is single-process; without torchrun this is rank 0 of 1)
    rank, local_rank, world = hd.env_world()
    if world > 1:
        hd.init_process_group("nccl")
    dev = dict(device=local_rank)

    if a.dataset == "amazon":                                   # hpmn.py:570-596
        if a.synthetic:
            feature_size = 256205
            trainset = synthetic_dataset(a.synthetic, 100, 3, feature_size, 100, 2, seed=1)
            testset = synthetic_dataset(max(a.synthetic // 2, 1), 100, 3, feature_size, 100, 2, seed=2)
        else:
            trainset, testset, feature_size = load_hpmn_pickle(os.path.join(a.data_root, "amazon/dataset_hpmn.pkl"))
        model = Hpmn(os.path.join(a.out, "amazon/hpmn/"), trainset, testset, feature_size, 3, 2, 100, 100, 0.003, 32, 16,
                     3, [2, 2, 5, 5, 1], [2, 2, 5, 5, 1], 3, 3, True, False, l2_reg=0., memory_reg=1e-5, max_batch=512, **dev)
        if a.eval_every:
            model.eval_every = a.eval_every
        best = model.train(a.epochs or 2, a.batchsize or 128)
        if rank == 0:
            model.save_model()
    elif a.dataset == "taobao":                                 # hpmn.py:598-624
        if a.synthetic:
            feature_size = 4000000
            trainset = synthetic_dataset(a.synthetic, 300, 4, feature_size, 36, 3, seed=1)
            testset = synthetic_dataset(max(a.synthetic // 2, 1), 300, 4, feature_size, 36, 3, seed=2)
        else:
            trainset, testset, feature_size = load_hpmn_pickle(os.path.join(a.data_root, "taobao/dataset_hpmn.pkl"))
            feature_size += 1   # the target btag id equals feature_size (preprocess_taobao.py:48,130,148)
        model = Hpmn(os.path.join(a.out, "taobao/hpmn/"), trainset, testset, feature_size, 4, 3, 300, 36, 0.001, 32, 16,
                     3, [2, 2, 3, 5, 5, 1], [2, 2, 3, 3, 1], 4, 5, True, False, l2_reg=0, memory_reg=1e-5, max_batch=512, **dev)
        if a.eval_every:
            model.eval_every = a.eval_every
        best = model.train(a.epochs or 2, a.batchsize or 128)
        if rank == 0:
            model.save_model()
    elif a.dataset == "xlong":                                  # hpmn.py:627-664
        pv_cnt = 19002
        if a.synthetic:
            from hpmn_b200.data_loader import write_synthetic_xlong
            os.makedirs(a.out, exist_ok=True)
            train_set = os.path.join(a.out, "synthetic_xlong_train.txt")
            test_set = os.path.join(a.out, "synthetic_xlong_test.txt")
            write_synthetic_xlong(train_set, a.synthetic // 2, seed=1)
            write_synthetic_xlong(test_set, max(a.synthetic // 4, 1), seed=2)
            graph_rows, emb_initializer = 3269017, None
        else:
            train_set = os.path.join(a.data_root, "xlong/train_corpus_total_dual.txt")
            test_set = os.path.join(a.data_root, "xlong/test_corpus_total_dual.txt")
            graph = np.load(os.path.join(a.data_root, "xlong/graph_emb.npy"))
            graph_rows = graph.shape[0]
            emb_initializer = np.concatenate((graph, np.zeros([20000, 16]), np.zeros([pv_cnt, 16])), 0).astype(np.float32)
        feature_size = pv_cnt + graph_rows + 20000
        model = Hpmn_Industry(os.path.join(a.out, "xlong/hpmn/"), train_set, test_set, feature_size, 2, 1, 1000 + 1, 184,
                              0.001, 32, 16, 3, [2] * 10 + [1], [3, 2, 2, 2, 2, 2, 2, 1], 5, 8, True, False,
                              emb_initializer, l2_reg=0, memory_reg=5e-5, max_batch=2048, **dev)
        if a.eval_every:
            model.eval_every = a.eval_every
        best = model.train(epochs=a.epochs or 3, batchsize=a.batchsize or 500)
        if rank == 0:
            model.get_weights()
    else:
        print("Dataset must be one of taobao or amazon.")       # sic, hpmn.py:666
        return 1
    if rank == 0:
        print("best test AUC: %.5f" % best)
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
