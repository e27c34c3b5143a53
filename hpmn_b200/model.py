"""Python-3 mirror of the reference's model classes (/root/reference/code/hpmn.py:16-560): same class
names, constructor signatures, methods and log / dump formats, with the TF1 session replaced by
HpmnEngine (libhpmn_b200.so).  `sess.run(train_step)` becomes engine.step_host + apply_gradients;
`sess.run([memory_loss, prediction])` becomes engine.step_host(with_backward=False).

`user=True, item=False` -- the graph every reference configuration runs (hpmn.py:591-592, 619-620, 658-659) -- is the fused
step of HpmnEngine.  `item=True` (second memory over `item_inp`, hpmn.py:444-462) runs on HpmnDualEngine, which composes both
sides and the wider head from the K1-K5 entry points; `user=False, item=True` is the single engine over the item side."""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch

from .data_loader import DataLoader, DataLoader_Mul
from .engine import HpmnEngine
from .layout import HpmnShape


def _metrics(labels, preds):
    from sklearn.metrics import log_loss, roc_auc_score   # same host-side metrics as hpmn.py:371-372
    return roc_auc_score(labels, preds), log_loss(labels, preds)


class Hpmn_Basic(object):
    """hpmn.py:16-214.  Subclasses define the variant (front padding, id-0 mask, target position, loader)."""

    item_front_pad = 0     # zero steps prepended to the item-side sequence
    item_scope = "item"    # variable scope of the item side (hpmn.py:444; Hpmn_Industry: "Item", hpmn.py:297)
    front_pad = 0          # zero steps prepended to the user sequence
    mask_id0 = True        # embedding of id 0 forced to zeros
    last_offset = 1        # target step counted from the end
    eval_every = 100       # hpmn.py:483 (Hpmn) / :338 (Hpmn_Industry)
    loader_cls = DataLoader

    def __init__(self, path, trainset, testset, feature_size, user_dim, item_dim, learning_rate, hidden_size,
                 embedding_size, hop, user_layers, item_layers, user_num_layers, item_num_layers, user, item,
                 emb_initializer=None, l2_reg=0, memory_reg=1e-5, max_batch=2048, device=0, seed=4321):
        if not user and not item:
            raise ValueError("at least one of user / item must be on (hpmn.py:452-462)")
        self._path = path
        self.trainset, self.testset = trainset, testset
        self._save_path = None
        self.feature_size = feature_size
        self.learning_rate = learning_rate
        self.l2_reg, self.memory_reg = l2_reg, memory_reg
        self.hidden_size, self.embedding_size, self.hop = hidden_size, embedding_size, hop
        self.emb_initializer = emb_initializer
        self.user_layers, self.item_layers = user_layers, item_layers
        self.user_num_layers, self.item_num_layers = user_num_layers, item_num_layers
        self.user_dim, self.item_dim = user_dim, item_dim
        self.user, self.item = user, item
        self.max_batch = max_batch
        self._device, self._seed = device, seed
        self._step_seed = 0
        self.define_inputs()
        self.build_graph()

    # ---- hpmn.py:64-72
    @property
    def save_path(self):
        if self._save_path is None:
            save_path = "%s/ckpt" % self._path
            os.makedirs(save_path, exist_ok=True)
            self._save_path = os.path.join(save_path, "model.ckpt")
        return self._save_path

    def define_inputs(self):
        raise NotImplementedError

    def build_graph(self):
        """hpmn.py:432-465 / 284-320: here the "graph" is the shape handed to the engine."""
        user_shape = HpmnShape(B=self.max_batch, T=self.user_maxlen, F=self.user_dim, E=self.embedding_size,
                               H=self.hidden_size, periods=list(self.user_layers), L=self.user_num_layers, hops=self.hop,
                               V=self.feature_size, front_pad=self.front_pad, mask_id0=self.mask_id0,
                               last_offset=self.last_offset)
        # item side: its own lengths / periods, target = last step, no id-0 mask change (hpmn.py:298-304, 444-450)
        item_shape = HpmnShape(B=self.max_batch, T=self.item_maxlen, F=self.item_dim, E=self.embedding_size,
                               H=self.hidden_size, periods=list(self.item_layers), L=self.item_num_layers, hops=self.hop,
                               V=self.feature_size, front_pad=self.item_front_pad, mask_id0=self.mask_id0, last_offset=1,
                               scope=self.item_scope)
        self.dual = bool(self.user and self.item)
        self.shape = user_shape if self.user else item_shape
        self.shape.steps()   # raises like TF's reshape would when a length is not divisible by its period
        if self.dual:
            from .dual import HpmnDualEngine
            item_shape.steps()
            self.item_shape = item_shape
            self.engine = HpmnDualEngine(user_shape, item_shape, device=self._device, memory_reg=self.memory_reg,
                                         l2_reg=self.l2_reg, table=self.emb_initializer, seed=self._seed)
            return
        from . import dist as hd
        world = hd.rank_world()[1]
        self.engine = HpmnEngine(self.shape, device=self._device, memory_reg=self.memory_reg, l2_reg=self.l2_reg,
                                 table=self.emb_initializer, seed=self._seed, symmetric=world > 1)
        if world > 1:          # torchrun: gradients are exchanged once per step over NVLink (hpmn_b200.dist.GradExchange)
            hd.GradExchange(self.engine).attach()

    # ---- hpmn.py:91-111
    def save_model(self, global_step=None):
        """tf.train.Saver saves every variable INCLUDING the Adam slots and beta powers (hpmn.py:61-62,91-99), so a
        resumed run continues the bias correction where it stopped.  Plain tensors only (loadable with weights_only=True)."""
        eng = self.engine
        state = {"step": torch.tensor(eng.adam_t), "table": eng.table.cpu()}
        for name, arr in eng.named_parameters().items():
            state["param:" + name] = torch.from_numpy(arr)
        if getattr(eng, "adam_m", None) is not None:
            state["adam_m"], state["adam_v"] = eng.adam_m.cpu(), eng.adam_v.cpu()
        path = self.save_path if global_step is None else "%s-%d" % (self.save_path, global_step)
        torch.save(state, path)

    def load_model(self):
        try:
            state = torch.load(self.save_path, weights_only=True)
            eng = self.engine
            eng.load_named({k[6:]: v.numpy() for k, v in state.items() if k.startswith("param:")})
            eng.table.copy_(state["table"])
            eng.adam_t = int(state["step"])
            if "adam_m" in state and hasattr(eng, "adam_m"):
                eng.adam_m = state["adam_m"].to(eng.device)
                eng.adam_v = state["adam_v"].to(eng.device)
        except Exception:
            raise IOError("Failed to load model from save path: %s" % self.save_path)
        print("Successfully load model from save path: %s" % self.save_path)

    def log(self, step, result):
        print("Step: %s\tTrain AUC: %.5f\tTrain Loss: %.5f\tTrain Mem_loss: %.5f"
              "\tTest AUC: %.5f\tTest Loss: %.5f\tTest Mem_loss: %.5f" % ((str(step),) + tuple(result[:6])))
        os.makedirs(self._path, exist_ok=True)
        with open(self._path + "/result.log", "a") as fout:
            fout.write("%s\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\n" % ((str(step),) + tuple(result[:6])))

    # ---- the two sess.run calls
    def _feed(self, data):
        """feed_dict of hpmn.py:474-481: `user_inp` is fed data[1] (the item_part of the tuple), `item_inp` data[3]."""
        ids = data[1] if self.user else data[3]
        return np.asarray(ids, dtype=np.int32), np.asarray(data[0], dtype=np.int32)

    def _dual_feed(self, data, lo, hi):
        dev = self.engine.device
        return (torch.as_tensor(np.ascontiguousarray(np.asarray(data[1], dtype=np.int32)[lo:hi]), device=dev),
                torch.as_tensor(np.ascontiguousarray(np.asarray(data[3], dtype=np.int32)[lo:hi]), device=dev),
                torch.as_tensor(np.asarray(data[0], dtype=np.int32)[lo:hi], device=dev))

    def train_on_batch(self, data):
        """One `sess.run(train_step)` (hpmn.py:482).  Under torchrun (torch.distributed initialised, world > 1) the batch
        rows are sharded across the ranks, every rank back-propagates its share of the GLOBAL-batch loss, the gradients
        are exchanged once (hpmn_b200.dist.exchange_grads) and every rank applies the identical clip + Adam update."""
        from . import dist as hd
        if self.dual:
            if len(data[0]) > self.max_batch:
                raise ValueError("train batch %d exceeds max_batch=%d" % (len(data[0]), self.max_batch))
            self._step_seed += 1
            u, i, y = self._dual_feed(data, 0, len(data[0]))
            self.engine.forward_backward(u, i, y, keep_prob=0.5, seed=self._step_seed)
            self.engine.apply_gradients(self.learning_rate)
            return
        ids, labels = self._feed(data)
        n = len(labels)
        rank, world = hd.rank_world()
        lo, hi = hd.shard_range(n, rank, world)
        if hi - lo > self.max_batch:
            raise ValueError("train batch %d exceeds max_batch=%d (pass a larger max_batch)" % (hi - lo, self.max_batch))
        self._step_seed += 1
        # dropout masks are keyed on (seed, local row): a distinct seed per rank keeps the shards' masks independent
        self.engine.step_host(ids[lo:hi], labels[lo:hi], with_backward=True, keep_prob=0.5,                    # hpmn.py:480
                              seed=self._step_seed * world + rank, loss_batch=n)
        if world > 1:
            hd.exchange_grads(self.engine)
        self.engine.apply_gradients(self.learning_rate)                                                        # hpmn.py:209-214

    def predict_on_batch(self, data):
        """eval fetch; batches larger than the engine capacity are evaluated in chunks (rows are independent; the
        memory loss is a sum over rows, hpmn.py:170)."""
        if self.dual:
            mem, preds, weights = 0.0, [], []
            for lo in range(0, len(data[0]), self.max_batch):
                hi = min(len(data[0]), lo + self.max_batch)
                u, i, y = self._dual_feed(data, lo, hi)
                self.engine.forward(u, i, y)
                sc = self.engine.scalars.cpu().numpy()
                mem += float(sc[1])
                preds.append(self.engine.pred[: hi - lo].cpu().numpy())
                weights.append(self.engine.user.w_hop0[: hi - lo].cpu().numpy())
            return mem, np.concatenate(preds), np.concatenate(weights)
        ids, labels = self._feed(data)
        mem, preds, weights = 0.0, [], []
        for lo in range(0, len(labels), self.max_batch):
            hi = min(len(labels), lo + self.max_batch)
            scalars, pred = self.engine.step_host(ids[lo:hi], labels[lo:hi], with_backward=False, keep_prob=1.0)   # hpmn.py:509
            mem += float(scalars[1])
            preds.append(pred.copy())
            weights.append(self.engine.h_w_hop0.numpy().reshape(-1)[: (hi - lo) * self.shape.L].reshape(hi - lo, self.shape.L).copy())
        return mem, np.concatenate(preds), np.concatenate(weights)

    # ---- hpmn.py:467-495 / 322-349
    def train(self, epochs, batchsize):
        from . import dist as hd
        is_main = hd.rank_world()[0] == 0          # replicas are identical: every rank evaluates, rank 0 logs
        step, count, best = 0, 0, 0.0
        for _ in range(epochs):
            loader = self.loader_cls(self.trainset, batchsize)
            try:
                for _, data in loader:
                    self.train_on_batch(data)
                    step += 1
                    if step % self.eval_every == 0:
                        result = list(self.eval(self.trainset, 4 * batchsize))
                        result += list(self.eval(self.testset, 4 * batchsize))
                        if is_main:
                            self.log(step, result)
                        if result[3] <= best:
                            count += 1
                            if count > 3:
                                return best
                        else:
                            count = 0
                            best = result[3]
            finally:
                loader.close()
        return best

    # ---- hpmn.py:497-519 / 351-373
    def eval(self, dataset, batchsize):
        labels, preds, mem_losses = [], [], []
        loader = self.loader_cls(dataset, batchsize)
        try:
            for _, data in loader:
                labels += list(data[0])
                mem_loss, pred, _ = self.predict_on_batch(data)
                mem_losses.append(mem_loss)
                preds += pred.tolist()
        finally:
            loader.close()
        auc, loss = _metrics(labels, preds)
        return auc, loss, float(np.average(mem_losses))


class Hpmn_Industry(Hpmn_Basic):
    """hpmn.py:217-410: no id-0 mask, 23 zero steps in front (1001 -> 1024), target = second to last step,
    XLong TSV loader, eval every 10 steps."""

    front_pad = 23
    item_front_pad = 8     # 192 - 184 zero steps in front of the item side (hpmn.py:298-299)
    item_scope = "Item"    # hpmn.py:297
    mask_id0 = False
    last_offset = 2
    eval_every = 10
    loader_cls = DataLoader_Mul

    def __init__(self, path, trainset, testset, feature_size, user_dim, item_dim, user_maxlen, item_maxlen,
                 learning_rate, hidden_size, embedding_size, hop, user_layers, item_layers, user_num_layers,
                 item_num_layers, user, item, emb_initializer=None, l2_reg=0, memory_reg=1e-5, **kw):
        self.user_maxlen, self.item_maxlen = user_maxlen, item_maxlen
        super(Hpmn_Industry, self).__init__(path, trainset, testset, feature_size, user_dim, item_dim, learning_rate,
                                            hidden_size, embedding_size, hop, user_layers, item_layers, user_num_layers,
                                            item_num_layers, user, item, emb_initializer, l2_reg, memory_reg, **kw)

    def define_inputs(self):
        """hpmn.py:247-264: placeholders -> pinned host staging inside the engine."""
        if self.front_pad:
            self.front_pad = 1024 - self.user_maxlen if self.user_maxlen <= 1024 else 0   # hpmn.py:288-290 pads to 1024

    # ---- hpmn.py:375-410
    def get_weights(self):
        weights, ids = [], []
        for ds in (self.trainset, self.testset):
            for _, data in self.loader_cls(ds, 512):
                _, _, w = self.predict_on_batch(data)
                weights += w.tolist()
                ids += np.asarray(data[1])[:, :, 1].tolist()
        np.save(self._path + "/weights_new.npy", np.array(weights))
        np.save(self._path + "/ids.npy", np.array(ids))


class Hpmn(Hpmn_Industry):
    """hpmn.py:413-560: id-0 mask, no front padding, target = last step, in-memory loader, eval every 100 steps."""

    front_pad = 0
    item_front_pad = 0
    item_scope = "item"    # hpmn.py:444
    mask_id0 = True
    last_offset = 1
    eval_every = 100
    loader_cls = DataLoader

    def get_weights(self):
        weights, lengths, labels = [], [], []
        for ds in (self.trainset, self.testset):
            for _, data in self.loader_cls(ds, 512):
                _, _, w = self.predict_on_batch(data)
                weights += w.tolist()
                lengths += list(data[2])
                labels += list(data[0])
        np.save(self._path + "/weights.npy", np.array(weights))
        np.save(self._path + "/lengths.npy", np.array(lengths))
        np.save(self._path + "/labels.npy", labels)
