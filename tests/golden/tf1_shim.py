"""Minimal TensorFlow-1.x API shim (lazy graph, evaluated with PyTorch-CPU) -- TEST INFRASTRUCTURE.

Purpose: TensorFlow 1.13 cannot be installed in this image, so the reference's ``model.py`` /
``aggregators.py`` cannot run as shipped.  This shim implements exactly the TF1 surface those two files touch
(placeholders, variables, a dozen math ops, ``Session.run`` with a feed dict, ``AdamOptimizer.minimize``) so
that ``tests/golden/make_golden.py`` can import the UNMODIFIED reference model from ``/root/reference`` and
execute the graph it builds.  The outputs are committed as ``tests/golden/*.npz`` and pin the oracle
(``oracle/mvin_oracle.py``) against the reference's own wiring.

Op semantics follow the TF1 documentation: ``matmul`` batches over leading dims, ``softmax`` acts on the last
axis, ``l2_loss = sum(x^2)/2``, ``reduce_mean`` over all elements when ``axis`` is None, ``gather`` /
``embedding_lookup`` index axis 0, ``sigmoid_cross_entropy_with_logits = max(x,0) - x z + log(1+exp(-|x|))``,
``AdamOptimizer``: lr_t = lr sqrt(1-b2^t)/(1-b1^t), var -= lr_t m/(sqrt(v)+eps).

Static shapes (the reference reads ``tensor.shape[1]`` while building the graph) are obtained by evaluating
every node once at construction on zero-filled placeholders whose unknown dimension is ``DEFAULT_BATCH``.
"""
from __future__ import annotations

import contextlib
import math
import sys
import types

import numpy as np
import torch

DEFAULT_BATCH = 4          # value substituted for ``None`` dims when tracing static shapes
_VARIABLES = []            # creation order, like tf.global_variables()
_SCOPE = []

int32, int64, float32, float64 = "int32", "int64", "float32", "float64"
_TORCH_DT = {"int32": torch.int32, "int64": torch.int64, "float32": torch.float32, "float64": torch.float64}
COMPUTE_DTYPE = torch.float32


def _proto(x):
    return x.proto if isinstance(x, Node) else x


class Node:
    def __init__(self, fn, inputs, name=None, proto=None):
        self.fn, self.inputs, self.name = fn, list(inputs), name
        self.proto = fn(*[_map(_proto, i) for i in self.inputs]) if proto is None else proto

    # static shape API used by the reference
    @property
    def shape(self):
        return tuple(int(s) for s in self.proto.shape)

    def get_shape(self):
        return self.shape

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __add__(self, o):  return Node(lambda a, b: a + b, [self, o])
    def __radd__(self, o): return Node(lambda a, b: b + a, [self, o])
    def __sub__(self, o):  return Node(lambda a, b: a - b, [self, o])
    def __rsub__(self, o): return Node(lambda a, b: b - a, [self, o])
    def __mul__(self, o):  return Node(lambda a, b: a * b, [self, o])
    def __rmul__(self, o): return Node(lambda a, b: b * a, [self, o])
    def __neg__(self):     return Node(lambda a: -a, [self])


def _map(f, x):
    if isinstance(x, (list, tuple)):
        return type(x)(_map(f, i) for i in x)
    return f(x)


class Placeholder(Node):
    def __init__(self, dtype, shape, name):
        self.dtype = dtype
        shp = [DEFAULT_BATCH if s is None else int(s) for s in shape]
        super().__init__(None, [], name=name, proto=torch.zeros(shp, dtype=_TORCH_DT[dtype]))


class Variable(Node):
    def __init__(self, value, name):
        self.value = value
        super().__init__(None, [], name=name, proto=value)


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name)


@contextlib.contextmanager
def variable_scope(name):
    _SCOPE.append(str(name))
    try:
        yield
    finally:
        _SCOPE.pop()


def _xavier(shape, gen):
    if len(shape) == 1:
        fi = fo = shape[0]
    elif len(shape) == 2:
        fi, fo = shape
    else:
        rec = int(np.prod(shape[:-2]))
        fi, fo = shape[-2] * rec, shape[-1] * rec
    lim = math.sqrt(6.0 / (fi + fo))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(COMPUTE_DTYPE)


class _Init:
    def __init__(self, kind, seed=None):
        self.kind, self.seed = kind, seed


_GEN = torch.Generator().manual_seed(1234)


def get_variable(name=None, shape=None, dtype=None, initializer=None):
    shape = [int(s) for s in shape]
    if initializer is not None and initializer.kind == "zeros":
        val = torch.zeros(shape, dtype=COMPUTE_DTYPE)
    else:
        val = _xavier(shape, _GEN)
    v = Variable(val, "/".join(_SCOPE + [name]) + ":0")
    _VARIABLES.append(v)
    return v


def zeros_initializer():
    return _Init("zeros")


def global_variables():
    return list(_VARIABLES)


def reset_default_graph():
    _VARIABLES.clear()


def global_variables_initializer():
    return None


def _idx(i):
    return torch.as_tensor(i).long()


def gather(params, indices):
    if isinstance(params, np.ndarray):
        const = torch.as_tensor(params)
        return Node(lambda i: const[_idx(i)], [indices])
    return Node(lambda p, i: p[_idx(i)], [params, indices])


def expand_dims(x, axis):
    return Node(lambda a: a.unsqueeze(axis), [x])


def tile(x, multiples):
    mult = [int(m) for m in multiples]
    return Node(lambda a: a.repeat(*mult), [x])


def concat(values, axis):
    return Node(lambda *a: torch.cat(list(a), dim=axis), list(values))


def reshape(x, shape):
    shp = [int(s) for s in shape]
    return Node(lambda a: a.reshape(shp), [x])


def squeeze(x, axis=None):
    return Node(lambda a: a.squeeze(axis) if axis is not None else a.squeeze(), [x])


def matmul(a, b):
    return Node(lambda x, y: torch.matmul(x, y), [a, b])


def reduce_sum(x, axis=None):
    return Node(lambda a: a.sum() if axis is None else a.sum(dim=axis), [x])


def reduce_mean(x, axis=None):
    return Node(lambda a: a.mean() if axis is None else a.mean(dim=axis), [x])


def sigmoid(x):
    return Node(torch.sigmoid, [x])


class _NN:
    @staticmethod
    def relu(x):
        return Node(torch.relu, [x])

    @staticmethod
    def embedding_lookup(params, ids):
        return gather(params, ids)

    @staticmethod
    def softmax(x, dim=-1):
        return Node(lambda a: torch.softmax(a, dim=dim), [x])

    @staticmethod
    def dropout(x, keep_prob=1.0):
        assert keep_prob == 1.0, "the reference only uses dropout with keep_prob = 1 (aggregators.py:80,109)"
        return x

    @staticmethod
    def l2_loss(x):
        return Node(lambda a: (a * a).sum() / 2, [x])

    @staticmethod
    def sigmoid_cross_entropy_with_logits(labels=None, logits=None):
        return Node(lambda z, x: torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x))),
                    [labels, logits])


nn = _NN()


class _TrainOp:
    def __init__(self, opt, loss):
        self.opt, self.loss = opt, loss


class AdamOptimizer:
    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps, self.t = learning_rate, beta1, beta2, epsilon, 0
        self.m, self.v = {}, {}
        self.last_grads = {}

    def minimize(self, loss):
        return _TrainOp(self, loss)

    def apply(self, loss_value, variables):
        grads = torch.autograd.grad(loss_value, [v.value for v in variables], allow_unused=True)
        self.t += 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        with torch.no_grad():
            for var, g in zip(variables, grads):
                g = torch.zeros_like(var.value) if g is None else g
                self.last_grads[var.name] = g.clone()
                m = self.m.setdefault(var.name, torch.zeros_like(var.value))
                v = self.v.setdefault(var.name, torch.zeros_like(var.value))
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                var.value.sub_(lr_t * m / (v.sqrt() + self.eps))


class _Train:
    AdamOptimizer = AdamOptimizer

    class Saver:
        def __init__(self, var_list=None):
            self.var_list = var_list

        def save(self, sess, path):
            np.savez(path, **{v.name: v.value.detach().numpy() for v in self.var_list})

        def restore(self, sess, path):
            z = np.load(path + ".npz")
            for v in self.var_list:
                v.value = torch.as_tensor(z[v.name])


train = _Train()


class Session:
    def __init__(self, config=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        if fetches is None:
            return None
        feed_dict = feed_dict or {}
        memo = {}
        want_grad = any(isinstance(f, _TrainOp) for f in _flatten(fetches))
        for v in _VARIABLES:
            v.value = v.value.detach().requires_grad_(want_grad)
        for ph, val in feed_dict.items():
            memo[id(ph)] = torch.as_tensor(np.asarray(val)).to(_TORCH_DT[ph.dtype])

        def ev(x):
            if isinstance(x, (list, tuple)):
                return type(x)(ev(i) for i in x)
            if not isinstance(x, Node):
                return x
            if id(x) in memo:
                return memo[id(x)]
            if isinstance(x, Variable):
                out = x.value
            elif isinstance(x, Placeholder):
                raise KeyError(f"placeholder {x.name} was not fed")
            else:
                out = x.fn(*[ev(i) for i in x.inputs])
            memo[id(x)] = out
            return out

        def fetch(f):
            if isinstance(f, (list, tuple)):
                return [fetch(i) for i in f]
            if isinstance(f, _TrainOp):
                return ("__train__", f)
            r = ev(f)
            return r.detach().numpy().copy() if isinstance(r, torch.Tensor) else r

        with torch.set_grad_enabled(want_grad):
            res = fetch(fetches)
            # apply optimizer ops last so every other fetch sees pre-update variables
            def finish(r):
                if isinstance(r, list):
                    return [finish(i) for i in r]
                if isinstance(r, tuple) and len(r) == 2 and r[0] == "__train__":
                    op = r[1]
                    op.opt.apply(ev(op.loss), _VARIABLES)
                    return None
                return r
            return finish(res)


def _flatten(x):
    if isinstance(x, (list, tuple)):
        for i in x:
            yield from _flatten(i)
    else:
        yield x


class ConfigProto:
    def __init__(self):
        self.gpu_options = types.SimpleNamespace(allow_growth=False)


def install():
    """Register this module as ``tensorflow`` (with ``tensorflow.contrib.layers``) in sys.modules."""
    me = sys.modules[__name__]
    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.xavier_initializer = lambda seed=None, uniform=True: _Init("xavier", seed)
    contrib.layers = layers
    me.contrib = contrib
    sys.modules["tensorflow"] = me
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.layers"] = layers
    return me
