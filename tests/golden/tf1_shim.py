"""A small stand-in for the TensorFlow 1.4 graph API  --  TEST INFRASTRUCTURE ONLY (fixture generation).

TensorFlow 1.4 / Python 2 cannot be installed in this image, so the reference cannot be run as published.  What CAN be
run is the reference's own, UNMODIFIED model file: /root/reference/code/hpmn.py parses under Python 3 and touches
TensorFlow through ~45 API functions.  This module implements exactly those functions as a lazy graph evaluated with
torch (CPU, float64), so that `tests/golden/make_reference_graph_fixture.py` can import the reference's hpmn.py with
`sys.modules['tensorflow']` pointing here, let the reference build its own graph (variable scopes, slices, reshapes,
gathers, attention hops, loss, clipped Adam) and record what it computes.

What this pins and what it does not:
  * pinned: everything hpmn.py itself decides -- the wiring of the graph, every shape / axis / slice / constant / scope
    name, the order of operations, which variables exist and which receive gradients;
  * not pinned: the arithmetic inside the TF1.4 ops, which is restated here from TF1.4's published behaviour
    (each function below says which); the in-tree copy of the GRUCell arithmetic, code/util.py:56-110, and the in-tree
    dynamic_rnn copy, code/rnn.py:588-807, are the anchors for the two ops that matter most.  The first of the two is
    also EXECUTED: `install()` provides the tensorflow.python.ops submodules util.py imports, and
    tests/test_reference_graph.py::test_gru_step_equals_the_reference_in_tree_cell runs the reference's VecAttGRUCell
    as it lies against `GRUCell.step` below and the oracle's GRU layer.

Nothing in the product, the oracle or the tests imports this file; only the fixture generator does, in the build
container.  Graph tensors declared `tf.float32` are evaluated in float64 so the fixtures are good to ~1e-15.
"""
from __future__ import annotations

import contextlib
import math
import pickle
import types
from collections import OrderedDict

import numpy as np
import torch

DT = torch.float64


class _DType:
    def __init__(self, name, torch_dtype):
        self.name, self.torch = name, torch_dtype

    def __repr__(self):
        return "tf." + self.name


float32 = _DType("float32", DT)       # evaluated in float64, see the header
float64 = _DType("float64", DT)
int32 = _DType("int32", torch.int64)


# ----------------------------------------------------------------------------------------------------------------------
# graph, scopes, variables
# ----------------------------------------------------------------------------------------------------------------------
class Graph:
    def __init__(self):
        self.variables = OrderedDict()       # full name -> Variable, creation order (= tf.global_variables())
        self.scope = []                      # variable-scope stack
        self.scope_count = {}                # full scope name -> times opened (default_name uniquification)
        self.rng = torch.Generator().manual_seed(20171101)

    @contextlib.contextmanager
    def as_default(self):
        _STACK.append(self)
        try:
            yield self
        finally:
            _STACK.pop()


_STACK = [Graph()]


def _g() -> Graph:
    return _STACK[-1]


def reset_default_graph():
    _STACK[-1] = Graph()


def set_random_seed(seed):
    _g().rng.manual_seed(int(seed))


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None):
    """tf.variable_scope: a named scope is entered as is; `None` + default_name is made unique among the scopes already
    opened under the current one: dense, dense_1, dense_2 ...  [TF1.4 variable_scope._get_unique_variable_scope]"""
    g = _g()
    prefix = "/".join(g.scope)
    if name_or_scope is None:
        base, idx = default_name, 0
        full = (prefix + "/" + base) if prefix else base
        name = base
        while g.scope_count.get(full, 0) > 0:
            idx += 1
            name = "%s_%d" % (base, idx)
            full = (prefix + "/" + name) if prefix else name
    else:
        name = name_or_scope
        full = (prefix + "/" + name) if prefix else name
    g.scope_count[full] = g.scope_count.get(full, 0) + 1
    g.scope.append(name)
    try:
        yield full
    finally:
        g.scope.pop()


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name


class TensorShape:
    def __init__(self, dims):
        self.dims = list(dims)

    def as_list(self):
        return list(self.dims)

    def __getitem__(self, i):
        return self.dims[i]

    def __len__(self):
        return len(self.dims)


class Tensor:
    """A lazy graph node: `fn(*values of inputs)` evaluated by Session.run.  At construction the node is evaluated once on
    probe values (placeholders: zeros with batch 1) so that static shapes are known, as they are in a TF graph."""

    def __init__(self, fn, inputs=(), name=None, probe=True, dyn0=None):
        self.fn, self.inputs, self.name = fn, list(inputs), name
        self.dyn0 = any(getattr(i, "dyn0", False) for i in self.inputs) if dyn0 is None else dyn0
        self.probe = None
        if probe:
            with torch.no_grad():
                self.probe = fn(*[_probe(i) for i in self.inputs])

    # static shape ------------------------------------------------------------------------------------------------
    def get_shape(self):
        dims = list(self.probe.shape)
        if self.dyn0 and dims:
            dims[0] = None
        return TensorShape(dims)

    @property
    def shape(self):
        return self.get_shape()

    # operators ---------------------------------------------------------------------------------------------------
    def __add__(self, o): return Tensor(lambda a, b: a + b, [self, o])
    def __radd__(self, o): return Tensor(lambda a, b: b + a, [self, o])
    def __sub__(self, o): return Tensor(lambda a, b: a - b, [self, o])
    def __rsub__(self, o): return Tensor(lambda a, b: b - a, [self, o])
    def __mul__(self, o): return Tensor(lambda a, b: a * b, [self, o])
    def __rmul__(self, o): return Tensor(lambda a, b: b * a, [self, o])
    def __truediv__(self, o): return Tensor(lambda a, b: a / b, [self, o])
    def __rtruediv__(self, o): return Tensor(lambda a, b: b / a, [self, o])
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __neg__(self): return Tensor(lambda a: -a, [self])
    def __getitem__(self, idx): return Tensor(lambda a: a[idx], [self])
    __hash__ = object.__hash__


class Variable(Tensor):
    def __init__(self, name, value: torch.Tensor, trainable=True):
        self.value = value.detach().clone().requires_grad_(bool(trainable) and value.dtype.is_floating_point)
        self.trainable = trainable
        super().__init__(lambda: self.value, [], name=name, probe=False, dyn0=False)
        self.probe = self.value.detach()
        self.op = types.SimpleNamespace(name=name)

    def numpy(self):
        return self.value.detach().numpy().copy()


class _Placeholder(Tensor):
    def __init__(self, dtype, shape):
        self.dtype, self.decl = dtype, shape
        dims = [1 if d is None else int(d) for d in (shape or [])]
        super().__init__(None, [], probe=False, dyn0=bool(shape) and shape[0] is None)
        self.probe = torch.ones(dims, dtype=dtype.torch) if dtype.torch.is_floating_point else torch.zeros(dims, dtype=dtype.torch)


def _probe(x):
    return x.probe if isinstance(x, Tensor) else _const(x)


def _const(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (bool, int, float)):
        return x
    a = np.asarray(x)
    return torch.as_tensor(a, dtype=DT if a.dtype.kind == "f" else torch.int64)


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, shape)


def _fp32_exact(t: torch.Tensor) -> torch.Tensor:
    """initial values are rounded to float32 so that a float32 copy of the fixture is exact"""
    return t.to(torch.float32).to(DT)


class _Init:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, shape):
        return self.fn(shape)


def glorot_uniform_initializer():
    """U(-l, l), l = sqrt(6 / (fan_in + fan_out)); the default of tf.get_variable [TF1.4 variable_scope.py: `initializer
    = init_ops.glorot_uniform_initializer()` for floating dtypes]"""
    def fn(shape):
        fan_in = shape[0] if len(shape) == 2 else (int(np.prod(shape[:-1])) if len(shape) > 1 else shape[0])
        fan_out = shape[-1]
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=_g().rng, dtype=DT) * 2 - 1) * lim
    return _Init(fn)


def constant_initializer(v=0.0):
    return _Init(lambda shape: torch.full(shape, float(v), dtype=DT))


def zeros_initializer():
    return constant_initializer(0.0)


def ones_initializer():
    return constant_initializer(1.0)


def random_normal_initializer(mean=0.0, stddev=1.0, **k):
    return _Init(lambda shape: mean + stddev * torch.randn(shape, generator=_g().rng, dtype=DT))


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    g = _g()
    full = "/".join(g.scope + [name])
    if full in g.variables:
        raise ValueError("Variable %s already exists, disallowed (reuse is not set)" % full)
    if initializer is None:
        initializer = glorot_uniform_initializer()
    if isinstance(initializer, _Init):
        val = initializer([int(s) for s in shape])
    else:                                               # a constant: numpy array or nested list (hpmn.py:268, 418)
        val = torch.as_tensor(np.asarray(initializer, dtype=np.float64))
    v = Variable(full, _fp32_exact(val), trainable=trainable)
    g.variables[full] = v
    return v


def trainable_variables():
    return [v for v in _g().variables.values() if v.trainable]


def global_variables():
    return list(_g().variables.values())


def global_variables_initializer():
    return Tensor(lambda: None, [], probe=False)        # variables hold their initial value from creation


# ----------------------------------------------------------------------------------------------------------------------
# array / math ops (semantics of the TF1.4 ops of the same name)
# ----------------------------------------------------------------------------------------------------------------------
def _op(fn, *inputs):
    return Tensor(fn, inputs)


def _ints(shape):
    return [int(s) for s in shape]


def reshape(x, shape, name=None):
    shape = _ints(shape)
    return _op(lambda a: a.reshape(shape), x)


def concat(values, axis, name=None):
    return Tensor(lambda *v: torch.cat(list(v), dim=axis), list(values))


def expand_dims(x, axis=None, name=None, dim=None):
    axis = dim if axis is None else axis
    return _op(lambda a: a.unsqueeze(axis), x)


def squeeze(x, axis=None, name=None):
    return _op(lambda a: a.squeeze() if axis is None else a.squeeze(axis), x)


def tile(x, multiples, name=None):
    multiples = _ints(multiples)
    return _op(lambda a: a.repeat(*multiples), x)


def gather(params, indices, axis=0, name=None):
    """tf.gather(params, indices, axis): output keeps the rank of `indices` in place of `axis`"""
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.int64)

    def fn(a):
        out = a.index_select(axis, idx.reshape(-1))
        return out.reshape(list(a.shape[:axis]) + list(idx.shape) + list(a.shape[axis + 1:]))
    return _op(fn, params)


def split(value, num_or_size_splits, axis=0, name=None):
    n = int(num_or_size_splits)
    return [_op(lambda a, i=i: torch.chunk(a, n, dim=axis)[i], value) for i in range(n)]


def transpose(x, perm=None, name=None):
    return _op(lambda a: a.permute(*perm) if perm is not None else a.t(), x)


def shape(x, name=None):
    return _op(lambda a: torch.as_tensor(list(a.shape), dtype=torch.int64), x)


def cast(x, dtype, name=None):
    return _op(lambda a: (a if isinstance(a, torch.Tensor) else torch.as_tensor(a)).to(dtype.torch), x)


def to_float(x, name=None):
    return cast(x, float32)


def zeros_like(x, dtype=None, name=None):
    return _op(lambda a: torch.zeros_like(a, dtype=dtype.torch if dtype is not None else a.dtype), x)


def ones_like(x, dtype=None, name=None):
    return _op(lambda a: torch.ones_like(a, dtype=dtype.torch if dtype is not None else a.dtype), x)


def multiply(a, b, name=None):
    return _op(lambda x, y: x * y, a, b)


def add(a, b, name=None):
    return _op(lambda x, y: x + y, a, b)


def subtract(a, b, name=None):
    return _op(lambda x, y: x - y, a, b)


def matmul(a, b, name=None):
    return _op(lambda x, y: torch.matmul(x, y), a, b)


def reduce_sum(x, axis=None, keep_dims=False, name=None):
    return _op(lambda a: a.sum() if axis is None else a.sum(dim=axis, keepdim=keep_dims), x)


def reduce_mean(x, axis=None, keep_dims=False, name=None):
    return _op(lambda a: a.mean() if axis is None else a.mean(dim=axis, keepdim=keep_dims), x)


def norm(x, ord="euclidean", axis=None, name=None):
    """tf.norm(ord='fro', axis=[i, j]) = sqrt(reduce_sum(x * x, axis)) [TF1.4 linalg_ops.norm]"""
    assert ord in ("fro", "euclidean", 2)
    return _op(lambda a: torch.sqrt((a * a).sum(dim=tuple(axis) if axis is not None else None)), x)


def sqrt(x, name=None):
    return _op(torch.sqrt, x)


def square(x, name=None):
    return _op(lambda a: a * a, x)


def exp(x, name=None):
    return _op(torch.exp, x)


def tanh(x, name=None):
    return _op(torch.tanh, x)


def sigmoid(x, name=None):
    return _op(torch.sigmoid, x)


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    """minimum(maximum(t, lo), hi); a sparse gradient is densified first [TF1.4 clip_ops.clip_by_value]"""
    return _op(lambda a: torch.clamp(a, clip_value_min, clip_value_max), t)


def constant(v, dtype=None, shape=None, name=None):
    return _op(lambda: _const(v))


def _diag_part(x, name=None):
    return _op(lambda a: torch.diagonal(a, dim1=-2, dim2=-1), x)


def _diag(x, name=None):
    return _op(lambda a: torch.diag_embed(a), x)


linalg = types.SimpleNamespace(diag_part=_diag_part, diag=_diag)


# ----------------------------------------------------------------------------------------------------------------------
# tf.nn
# ----------------------------------------------------------------------------------------------------------------------
def _embedding_lookup(params, ids, name=None):
    return _op(lambda p, i: p[i], params, ids)


def _dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
    """x / keep_prob * floor(keep_prob + U[0,1))  [TF1.4 nn_ops.dropout]; keep_prob is a fed tensor here, so the random
    mask is always drawn; with keep_prob = 1 it is all ones"""
    def fn(a, kp):
        u = torch.rand(a.shape, generator=_g_rng[0], dtype=DT)
        return a / kp * torch.floor(kp + u)
    return _op(fn, x, keep_prob)


_g_rng = [torch.Generator().manual_seed(7)]


def _l2_loss(t, name=None):
    return _op(lambda a: (a * a).sum() / 2, t)           # sum(t ** 2) / 2


def _elu(x, name=None):
    return _op(lambda a: torch.where(a > 0, a, torch.expm1(a)), x)


def _softmax(x, dim=-1, name=None):
    return _op(lambda a: torch.softmax(a, dim=dim), x)


class GRUCell:
    """tf.nn.rnn_cell.GRUCell [TF1.4 rnn_cell_impl.GRUCell.call; the reference keeps a copy of the same arithmetic in
    code/util.py:81-110]:
        [r, u] = split(sigmoid([x, h] @ gates/kernel + gates/bias), 2)        gates/bias initialised to 1
        c      = tanh([x, r * h] @ candidate/kernel + candidate/bias)          candidate/bias initialised to 0
        h'     = u * h + (1 - u) * c
    Variables are created on first call under <current scope>/gru_cell/."""

    def __init__(self, num_units, activation=None, reuse=None, kernel_initializer=None, bias_initializer=None):
        self.num_units = int(num_units)
        self.vars = None

    def build(self, input_depth):
        H = self.num_units
        with variable_scope("gru_cell"):
            with variable_scope("gates"):
                wg = get_variable("kernel", [input_depth + H, 2 * H])
                bg = get_variable("bias", [2 * H], initializer=constant_initializer(1.0))
            with variable_scope("candidate"):
                wc = get_variable("kernel", [input_depth + H, H])
                bc = get_variable("bias", [H], initializer=constant_initializer(0.0))
        self.vars = (wg, bg, wc, bc)

    @staticmethod
    def step(x, h, wg, bg, wc, bc):
        H = h.shape[1]
        value = torch.sigmoid(torch.cat([x, h], dim=1) @ wg + bg)
        r, u = value[:, :H], value[:, H:]
        c = torch.tanh(torch.cat([x, r * h], dim=1) @ wc + bc)
        return u * h + (1 - u) * c


def _dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, time_major=False, scope=None):
    """tf.nn.dynamic_rnn without sequence_length: zero initial state, every step of every row is run, `outputs` stacks the
    states, the final state is the last one [TF1.4 rnn.dynamic_rnn / _dynamic_rnn_loop; in-tree copy code/rnn.py:588-807].
    Variables live under <scope>/rnn/."""
    assert sequence_length is None and initial_state is None and not time_major
    with variable_scope(scope or "rnn"):
        if cell.vars is None:
            cell.build(int(inputs.probe.shape[-1]))
    wg, bg, wc, bc = cell.vars

    def fn(x, wg, bg, wc, bc):
        B, T, _ = x.shape
        h = torch.zeros(B, cell.num_units, dtype=x.dtype)
        outs = []
        for t in range(T):
            h = GRUCell.step(x[:, t], h, wg, bg, wc, bc)
            outs.append(h)
        return torch.stack(outs, dim=1), h
    both = Tensor(fn, [inputs, wg, bg, wc, bc])
    return _op(lambda p: p[0], both), _op(lambda p: p[1], both)


nn = types.SimpleNamespace(
    embedding_lookup=_embedding_lookup, dropout=_dropout, l2_loss=_l2_loss, softmax=_softmax,
    relu=lambda x, name=None: _op(torch.relu, x), elu=_elu, sigmoid=sigmoid, tanh=tanh,
    dynamic_rnn=_dynamic_rnn, rnn_cell=types.SimpleNamespace(GRUCell=GRUCell))


# ----------------------------------------------------------------------------------------------------------------------
# tf.layers, tf.losses
# ----------------------------------------------------------------------------------------------------------------------
def _dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, name=None, **kw):
    """tf.layers.dense: act(inputs . kernel + bias) over the last axis; kernel [in, units] (get_variable default
    initializer), bias zeros; the layer's scope is `name` or the unique default `dense`, `dense_1`, ...
    [TF1.4 layers/core.py Dense, layers/base.py Layer._set_scope]"""
    with variable_scope(name, default_name="dense"):
        k = get_variable("kernel", [int(inputs.probe.shape[-1]), int(units)], initializer=kernel_initializer)
        b = get_variable("bias", [int(units)], initializer=zeros_initializer())
    out = _op(lambda x, k, b: torch.matmul(x, k) + b, inputs, k, b)
    return activation(out) if activation is not None else out


def _batch_normalization(inputs, axis=-1, momentum=0.99, epsilon=1e-3, training=False, name=None, **kw):
    """tf.layers.batch_normalization with the default training=False: the moving statistics (initialised to mean 0,
    variance 1 and never updated by this graph) normalise the input:
        inv = rsqrt(moving_variance + eps) * gamma ;  y = x * inv + (beta - moving_mean * inv)
    [TF1.4 layers/normalization.py BatchNormalization.call -> nn.batch_normalization]"""
    assert training is False and axis == -1
    C = int(inputs.probe.shape[-1])
    with variable_scope(name, default_name="batch_normalization"):
        gamma = get_variable("gamma", [C], initializer=ones_initializer())
        beta = get_variable("beta", [C], initializer=zeros_initializer())
        mean = get_variable("moving_mean", [C], initializer=zeros_initializer(), trainable=False)
        var = get_variable("moving_variance", [C], initializer=ones_initializer(), trainable=False)

    def fn(x, gamma, beta, mean, var):
        inv = torch.rsqrt(var + epsilon) * gamma
        return x * inv + (beta - mean * inv)
    return _op(fn, inputs, gamma, beta, mean, var)


layers = types.SimpleNamespace(dense=_dense, batch_normalization=_batch_normalization)


def _log_loss(labels, predictions, weights=1.0, epsilon=1e-7, scope=None, **kw):
    """tf.losses.log_loss: both sides cast to float, -y log(p + eps) - (1 - y) log(1 - p + eps), reduced with
    SUM_BY_NONZERO_WEIGHTS = mean over the elements for weights = 1 [TF1.4 losses/losses_impl.py log_loss]"""
    def fn(y, p):
        y = y.to(DT)
        losses = -y * torch.log(p + epsilon) - (1 - y) * torch.log(1 - p + epsilon)
        return losses.sum() / losses.numel()
    return _op(fn, labels, predictions)


losses = types.SimpleNamespace(log_loss=_log_loss)


# ----------------------------------------------------------------------------------------------------------------------
# tf.train
# ----------------------------------------------------------------------------------------------------------------------
class _Gradients(Tensor):
    def __init__(self, loss, variables):
        self.loss_node, self.variables = loss, variables
        super().__init__(None, [loss] + list(variables), probe=False, dyn0=False)


class AdamOptimizer:
    """tf.train.AdamOptimizer [TF1.4 training/adam.py]:  lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);
    m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  var -= lr_t * m / (sqrt(v) + eps)   (dense update of every row:
    the clipped embedding gradient is a dense tensor, see clip_by_value)"""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **kw):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0
        self.slots = {}
        self.raw_gvs = None       # what compute_gradients returned           } kept for the fixture generator: the
        self.applied_gvs = None   # what apply_gradients was given (clipped)  } reference keeps neither

    def compute_gradients(self, loss, var_list=None):
        variables = var_list or trainable_variables()
        bundle = _Gradients(loss, variables)
        out = []
        for i, v in enumerate(variables):
            gnode = Tensor(lambda b, i=i: b[i], [bundle], probe=False, dyn0=False)
            gnode.probe = torch.zeros_like(v.value)
            out.append((gnode, v))
        self.raw_gvs = out
        return out

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gvs = list(grads_and_vars)
        self.applied_gvs = gvs

        def fn(*grads):
            self.t += 1
            lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
            with torch.no_grad():
                for (_, var), g in zip(gvs, grads):
                    m, v = self.slots.setdefault(var.name, (torch.zeros_like(var.value), torch.zeros_like(var.value)))
                    m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                    v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                    var.value.sub_(lr_t * m / (torch.sqrt(v) + self.eps))
            return None
        return Tensor(fn, [g for g, _ in gvs], probe=False, dyn0=False)

    def minimize(self, loss, **kw):
        return self.apply_gradients(self.compute_gradients(loss))


class Saver:
    def __init__(self, *a, **k):
        self.graph = _g()

    def save(self, sess, path, global_step=None):
        with open(path, "wb") as f:
            pickle.dump({k: v.numpy() for k, v in self.graph.variables.items()}, f)
        return path

    def restore(self, sess, path):
        with open(path, "rb") as f:
            for k, a in pickle.load(f).items():
                with torch.no_grad():
                    self.graph.variables[k].value.copy_(torch.as_tensor(a))


train = types.SimpleNamespace(AdamOptimizer=AdamOptimizer, Saver=Saver)


# ----------------------------------------------------------------------------------------------------------------------
# session
# ----------------------------------------------------------------------------------------------------------------------
class ConfigProto:
    def __init__(self, *a, **k):
        self.gpu_options = types.SimpleNamespace(allow_growth=False)


class GPUOptions:
    def __init__(self, *a, **k):
        pass


class Session:
    def __init__(self, graph=None, config=None):
        self.graph = graph or _g()

    def run(self, fetches, feed_dict=None):
        feed = {}
        for k, v in (feed_dict or {}).items():
            a = np.asarray(v)
            feed[k] = torch.as_tensor(a, dtype=k.dtype.torch) if isinstance(k, _Placeholder) else torch.as_tensor(a)
        cache = {}
        single = not isinstance(fetches, (list, tuple))
        with _as_default(self.graph):
            vals = [_eval(f, cache, feed) for f in ([fetches] if single else fetches)]
        out = [v.detach().numpy().copy() if isinstance(v, torch.Tensor) else v for v in vals]
        return out[0] if single else out

    def close(self):
        pass


@contextlib.contextmanager
def _as_default(g):
    _STACK.append(g)
    try:
        yield
    finally:
        _STACK.pop()


def _eval(node, cache, feed):
    """iterative post-order evaluation (the graphs are shallow: dynamic_rnn is one node)"""
    if not isinstance(node, Tensor):
        return _const(node)
    key = id(node)
    if key in cache:
        return cache[key]
    if isinstance(node, _Placeholder):
        if node not in feed:
            raise ValueError("placeholder was not fed")
        val = feed[node]
    elif isinstance(node, Variable):
        val = node.value
    elif isinstance(node, _Gradients):
        loss = _eval(node.loss_node, cache, feed)
        vs = [v.value for v in node.variables]
        gs = torch.autograd.grad(loss, vs, allow_unused=True, retain_graph=True)
        val = [torch.zeros_like(v) if g is None else g for g, v in zip(gs, vs)]
    else:
        val = node.fn(*[_eval(i, cache, feed) for i in node.inputs])
    cache[key] = val
    return val


class _FileWriter:
    def __init__(self, *a, **k):
        pass

    def add_summary(self, *a, **k):
        pass


summary = types.SimpleNamespace(FileWriter=_FileWriter, scalar=lambda *a, **k: None, histogram=lambda *a, **k: None,
                                merge_all=lambda: None)


# ----------------------------------------------------------------------------------------------------------------------
# tensorflow.python.ops.* -- only what code/util.py imports, so that the reference's in-tree copy of the GRU cell
# arithmetic (VecAttGRUCell, util.py:56-110) can be executed as it lies (tests/test_reference_graph.py)
# ----------------------------------------------------------------------------------------------------------------------
class RNNCell:
    def __init__(self, _reuse=None, **kw):
        pass


class _Linear:
    """tensorflow.python.ops.rnn_cell_impl._Linear [TF1.4]: concat(args, 1) @ kernel (+ bias); `kernel` [sum of the args' widths,
    output_size] with the scope's default initializer unless one is given, `bias` [output_size] initialised to 0 unless
    a bias_initializer is given; both created in the variable scope that is current at construction."""

    def __init__(self, args, output_size, build_bias, bias_initializer=None, kernel_initializer=None):
        args = list(args) if isinstance(args, (list, tuple)) else [args]
        total = sum(int(a.probe.shape[1]) for a in args)
        self.kernel = get_variable("kernel", [total, int(output_size)], initializer=kernel_initializer)
        self.bias = None
        if build_bias:
            self.bias = get_variable("bias", [int(output_size)],
                                     initializer=bias_initializer if bias_initializer is not None else constant_initializer(0.0))

    def __call__(self, args):
        args = list(args) if isinstance(args, (list, tuple)) else [args]
        res = matmul(concat(args, 1), self.kernel)
        return res + self.bias if self.bias is not None else res


Tensor.dtype = float32      # `inputs.dtype` (util.py:84)


def install(sys_modules):
    """sys.modules entries for `import tensorflow` and the tensorflow.python.ops submodules code/util.py imports"""
    import sys as _sys
    me = _sys.modules[__name__]
    ns = types.SimpleNamespace
    mods = {
        "tensorflow": me,
        "tensorflow.python": types.ModuleType("tensorflow.python"),
        "tensorflow.python.ops": types.ModuleType("tensorflow.python.ops"),
        "tensorflow.python.ops.array_ops": types.ModuleType("array_ops"),
        "tensorflow.python.ops.init_ops": types.ModuleType("init_ops"),
        "tensorflow.python.ops.math_ops": types.ModuleType("math_ops"),
        "tensorflow.python.ops.variable_scope": types.ModuleType("variable_scope"),
        "tensorflow.python.ops.rnn_cell": types.ModuleType("rnn_cell"),
        "tensorflow.python.ops.rnn_cell_impl": types.ModuleType("rnn_cell_impl"),
    }
    a = mods["tensorflow.python.ops.array_ops"]; a.split, a.concat = split, concat
    i = mods["tensorflow.python.ops.init_ops"]
    i.constant_initializer = lambda value=0.0, dtype=None: constant_initializer(value)
    m = mods["tensorflow.python.ops.math_ops"]; m.sigmoid, m.tanh, m.matmul = sigmoid, tanh, matmul
    v = mods["tensorflow.python.ops.variable_scope"]; v.variable_scope, v.get_variable = variable_scope, get_variable
    r = mods["tensorflow.python.ops.rnn_cell"]; r.RNNCell, r.GRUCell = RNNCell, GRUCell
    r.__all__ = ["RNNCell", "GRUCell"]
    mods["tensorflow.python.ops.rnn_cell_impl"]._Linear = _Linear
    sys_modules.update(mods)
    return me
