"""Python-3 restatement of the two loaders that feed HPMN in the reference
(/root/reference/code/data_loader.py): `DataLoader` (in-memory tuples, :267-301) and `DataLoader_Mul`
(XLong TSV, :7-107), plus synthetic datasets of the BASELINE.json shapes.  Batches keep the reference's
5-tuple `(label, item_part, item_part_len, user_part, user_part_len)` (data_loader.py:298)."""
from __future__ import annotations

import pickle
import queue
import threading
from typing import Iterator, List, Sequence, Tuple

import numpy as np


def load_hpmn_pickle(path: str):
    """dataset_hpmn.pkl = three consecutive py2 protocol-0 pickles: train list, test list, feature_size
    (hpmn.py:571-575; tuples laid out by util.py:152-159)."""
    with open(path, "rb") as f:
        train = pickle.load(f, encoding="latin1")
        test = pickle.load(f, encoding="latin1")
        feature_size = pickle.load(f, encoding="latin1")
    return train, test, int(feature_size)


class DataLoader:
    """data_loader.py:267-301: consecutive slices of `batch_size` tuples, last batch may be short."""

    def __init__(self, dataset: Sequence, batch_size: int):
        self.batch_size = batch_size
        self.dataset = dataset
        self.num_of_step = len(dataset) // batch_size
        if batch_size * self.num_of_step < len(dataset):
            self.num_of_step += 1
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.i == self.num_of_step:
            raise StopIteration
        ts = self.dataset[self.i * self.batch_size: min(len(self.dataset), (self.i + 1) * self.batch_size)]
        label = [t[0] for t in ts]
        item_part = np.array([t[1] for t in ts], dtype=np.int32)
        item_part_len = [t[2] for t in ts]
        user_part = np.array([t[3] for t in ts], dtype=np.int32)
        user_part_len = [t[4] for t in ts]
        self.i += 1
        return self.i, (label, item_part, item_part_len, user_part, user_part_len)

    next = __next__

    def close(self):
        pass


XLONG_ITEM_CNT = 3269017   # data_loader.py:49
XLONG_ITEM_LEN = 1000 + 1  # data_loader.py:56
XLONG_USER_LEN = 184


def parse_xlong_lines(lines: List[str]):
    """data_loader.py:56-85: `index \\t uid \\t hist,... \\t pos \\t neg \\t userseq_pos \\t userseq_neg`; every line
    yields a positive and a negative sample sharing the history; the uid column is uid + 3269017."""
    label, item_part, user_part = [], [], []
    for line in lines:
        it = line.rstrip("\n").split("\t")
        uid = int(it[1]) + XLONG_ITEM_CNT
        hist = [[uid, int(i)] for i in it[2].split(",")]
        item_part.append(hist + [[uid, int(it[3])]])
        item_part.append(hist + [[uid, int(it[4])]])
        user_part.append([int(v) for v in it[5].split(",")])
        user_part.append([int(v) for v in it[6].split(",")])
        label += [1, 0]
    n = len(label)
    return (label, np.array(item_part, dtype=np.int32), [XLONG_ITEM_LEN] * n,
            np.expand_dims(np.array(user_part, dtype=np.int32), 2), [XLONG_USER_LEN] * n)


class DataLoader_Mul:
    """data_loader.py:7-107 reads `batchsize/2` lines per batch with 1 producer + 8 worker processes; here one
    background thread parses ahead (order preserved, which the reference's worker pool does not guarantee)."""

    def __init__(self, dataset: str, batchsize: int, max_q_size: int = 10, wait_time: float = 0.1, worker_n: int = 8):
        self.batch_size = batchsize // 2
        self.path = dataset
        self.q: "queue.Queue" = queue.Queue(maxsize=max_q_size)
        self._stop = threading.Event()
        self.thread = threading.Thread(target=self._produce, daemon=True)
        self.thread.start()

    def _put(self, item) -> bool:
        while not self._stop.is_set():
            try:
                self.q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _produce(self):
        """Always ends with a sentinel: None after the last batch, or the exception a malformed line raised (re-raised by
        __next__, so a bad file fails the job instead of hanging it)."""
        try:
            with open(self.path) as f:
                while not self._stop.is_set():
                    lines = []
                    for _ in range(self.batch_size):
                        line = f.readline()
                        if not line:
                            break
                        lines.append(line)
                    if lines and not self._put((None, parse_xlong_lines(lines))):
                        return
                    if len(lines) < self.batch_size:
                        break
            self._put(None)
        except BaseException as e:  # noqa: BLE001 -- forwarded to the consumer
            self._put(e)

    def close(self):
        """Stop the producer and drop what it parsed ahead (train() leaving early on the early-stop rule, eval() done)."""
        self._stop.set()
        try:
            while True:
                self.q.get_nowait()
        except queue.Empty:
            pass

    def __iter__(self):
        return self

    def __next__(self):
        item = self.q.get()
        if item is None:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        return item

    next = __next__


# ---- synthetic data of the BASELINE.json shapes -------------------------------------------------

def synthetic_ids(B: int, T: int, F: int, V: int, seed: int = 1234, ragged: bool = False, zipf: float = 0.0,
                  uid_col: bool = False) -> np.ndarray:
    """ids uniform on [1,V) (worst case for caches) or Zipf(zipf); for F >= 3 -- or uid_col=True, the XLong feed, whose
    every step is (uid + offset, item) with ONE uid per sample (data_loader.py:58-67) -- column 0 is constant along t
    (T colliding atomics per sample in the scatter-add); ragged: random-length suffix kept, prefix id 0."""
    rng = np.random.default_rng(seed)
    if zipf > 0:
        ids = (rng.zipf(zipf, size=(B, T, F)) % (V - 1) + 1).astype(np.int64)
    else:
        ids = rng.integers(1, V, size=(B, T, F), dtype=np.int64)
    if F >= 3 or uid_col:
        ids[:, :, 0] = rng.integers(1, V, size=(B, 1), dtype=np.int64) if zipf > 0 else ids[:, :1, 0]
    if ragged:
        lens = rng.integers(min(5, T), T + 1, size=B)
        for b in range(B):
            ids[b, : T - lens[b]] = 0
    return ids.astype(np.int32)


def synthetic_dataset(n: int, T: int, F: int, V: int, user_T: int = 4, user_F: int = 2, seed: int = 1234):
    """A list of reference-layout tuples (label, item_part, item_len, user_part, user_len) for the in-memory loader."""
    ids = synthetic_ids(n, T, F, V, seed=seed, ragged=True)
    rng = np.random.default_rng(seed + 1)
    labels = rng.integers(0, 2, size=n)
    out = []
    for i in range(n):
        il = int((ids[i, :, -1] != 0).sum())
        up = np.zeros((user_T, user_F), dtype=np.int32)
        out.append((int(labels[i]), ids[i].tolist(), il, up.tolist(), 0))
    return out


def write_synthetic_xlong(path: str, n_lines: int, seed: int = 1, hist: int = 1000, user_len: int = XLONG_USER_LEN,
                          n_items: int = XLONG_ITEM_CNT, n_users: int = 19002 + 20000):
    """A synthetic file in the XLong TSV layout parse_xlong_lines reads (item ids < 3269017, uid offset added
    by the parser, so uid < pv_cnt + 20000 keeps every id inside feature_size)."""
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for i in range(n_lines):
            uid = int(rng.integers(0, n_users))
            h = ",".join(map(str, rng.integers(1, n_items, size=hist)))
            pos, neg = int(rng.integers(1, n_items)), int(rng.integers(1, n_items))
            up = ",".join(map(str, rng.integers(1, n_items, size=user_len)))
            un = ",".join(map(str, rng.integers(1, n_items, size=user_len)))
            f.write("%d\t%d\t%s\t%d\t%d\t%s\t%s\n" % (i, uid, h, pos, neg, up, un))
