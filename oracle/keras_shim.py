"""A minimal stand-in for the Keras-2 functional API and backend, evaluated with torch float64 on the CPU.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Purpose: the reference (febrianrachmadi/dep-gan-im) is Keras-2 / TF-1 code and
neither library exists here, so its scripts cannot be imported.  With this shim the reference's OWN source text -- the
layer helpers, Dis_C2D_FCN1, Gen_UNet2D and the WGAN-GP / generator-loss graph construction, read from /root/reference at
fixture-generation time and exec'ed unmodified (tests/golden/make_reference_vectors.py) -- runs and produces numbers.
The golden vectors made that way pin everything the reference's code decides: layer names, creation order and weight
shapes, the wiring of the U-ResNet and of the noise / FiLM path, concatenation order, the composition and signs of the
three loss graphs, the gradient-penalty axes, which weights each optimizer updates and with which hyper-parameters.
What the shim itself restates (from the Keras 2.x documentation / source, third-party dependency absent from
/root/reference; the reference pins no version, its README says Keras 2 on TensorFlow 1):
  Conv2D / Conv2DTranspose ('same' / 'valid', channels_last, HWIO and HWOI kernels), Dense on the last axis,
  BatchNormalization(axis=-1, epsilon=1e-3) and Dropout in learning phase 0, MaxPooling2D(2, 2), Flatten, Reshape, Lambda,
  the merge layers' rank broadcasting (a lower-rank input gets axes inserted at position 1), K.gradients = gradient of the
  sum, K.function = outputs evaluated on the pre-update weights followed by the updates, and optimizers.Adam.get_updates
  (lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); p -= lr_t * m / (sqrt(v) + epsilon), epsilon = K.epsilon() = 1e-7).
Only tests/ and the fixture generator import this module.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64
_counters: dict = {}


def reset_names():
    _counters.clear()


def _auto_name(prefix):
    _counters[prefix] = _counters.get(prefix, 0) + 1
    return "%s_%d" % (prefix, _counters[prefix])


# --------------------------------------------------------------------------------------------------------------
# lazy tensors
# --------------------------------------------------------------------------------------------------------------
class Sym:
    """A node of the symbolic graph: fn(*evaluated args).  fn is None for placeholders (fed at evaluation time)."""

    def __init__(self, fn, args=(), shape=None, name=None, layer=None):
        self.fn, self.args, self.shape, self.name, self.layer = fn, tuple(args), shape, name, layer

    # arithmetic the reference's graph code uses on tensors
    def __add__(self, o): return _binop(torch.add, self, o)
    def __radd__(self, o): return _binop(torch.add, o, self)
    def __sub__(self, o): return _binop(torch.sub, self, o)
    def __rsub__(self, o): return _binop(torch.sub, o, self)
    def __mul__(self, o): return _binop(torch.mul, self, o)
    def __rmul__(self, o): return _binop(torch.mul, o, self)
    def __truediv__(self, o): return _binop(torch.div, self, o)
    __div__ = __truediv__
    def __neg__(self): return Sym(lambda a: -a, (self,), self.shape)
    def __getitem__(self, idx): return Sym(lambda a: a[idx], (self,))


def _lift(v):
    return v if isinstance(v, Sym) else Sym(lambda: torch.as_tensor(v, dtype=DT), ())


def _binop(op, a, b):
    a, b = _lift(a), _lift(b)
    return Sym(lambda x, y: op(x, y), (a, b), a.shape or b.shape)


def evaluate(nodes, feed):
    """Values of `nodes` given {placeholder Sym: tensor}.  One memo per call: every node is computed once."""
    memo = {}
    for k, v in feed.items():  # fed tensors are differentiable leaves, so K.gradients may name a placeholder
        if torch.is_grad_enabled() and v.is_floating_point() and v.is_leaf and not v.requires_grad:
            v = v.detach().clone().requires_grad_(True)
        memo[id(k)] = v

    def ev(n):
        k = id(n)
        if k in memo:
            return memo[k]
        if n.fn is None:
            raise KeyError("placeholder %r was not fed" % (n.name,))
        stack = [n]
        while stack:  # iterative post-order: the generator graph is a few hundred nodes deep
            cur = stack[-1]
            if id(cur) in memo:
                stack.pop()
                continue
            if cur.fn is None:
                raise KeyError("placeholder %r was not fed" % (cur.name,))
            pending = [a for a in cur.args if isinstance(a, Sym) and id(a) not in memo]
            if pending:
                stack.extend(pending)
                continue
            memo[id(cur)] = cur.fn(*[memo[id(a)] if isinstance(a, Sym) else a for a in cur.args])
            stack.pop()
        return memo[k]

    return [ev(n) for n in nodes]


# --------------------------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------------------------
class Layer:
    prefix = "layer"

    def __init__(self, name=None, **_):
        self.name = name if name is not None else _auto_name(self.prefix)
        self.weights = OrderedDict()      # weight name -> tensor (created by build on the first call)
        self.non_trainable = set()
        self.built = False

    def build(self, in_shape):
        pass

    def out_shape(self, in_shape):
        return in_shape

    def forward(self, x):
        raise NotImplementedError

    def __call__(self, x):
        if not self.built:
            self.build(x.shape)
            self.built = True
        return Sym(self.forward, (x,), self.out_shape(x.shape), layer=self)

    def _add(self, wname, shape, value, trainable=True):
        self.weights[wname] = torch.full(tuple(shape), float(value), dtype=DT, requires_grad=trainable)
        if not trainable:
            self.non_trainable.add(wname)


class Dense(Layer):
    prefix = "dense"

    def __init__(self, units, kernel_initializer=None, name=None, **kw):
        super().__init__(name)
        self.units = int(units)

    def build(self, s):
        self._add("kernel", (s[-1], self.units), 0.0)
        self._add("bias", (self.units,), 0.0)

    def out_shape(self, s):
        return tuple(s[:-1]) + (self.units,)

    def forward(self, x):
        return x @ self.weights["kernel"] + self.weights["bias"]


class Conv2D(Layer):
    prefix = "conv2d"

    def __init__(self, filters, kernel_size, padding="valid", strides=(1, 1), kernel_initializer=None, name=None, **kw):
        super().__init__(name)
        self.filters = int(filters)
        self.ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.padding, self.strides = padding, tuple(strides)
        assert self.strides == (1, 1) and self.ks[0] == self.ks[1] and self.ks[0] % 2 == 1

    def build(self, s):
        self._add("kernel", self.ks + (s[-1], self.filters), 0.0)   # HWIO
        self._add("bias", (self.filters,), 0.0)

    def out_shape(self, s):
        assert self.padding == "same" or self.ks == (1, 1)
        return (s[0], s[1], s[2], self.filters)

    def forward(self, x):  # x NHWC
        w = self.weights["kernel"].permute(3, 2, 0, 1)
        pad = self.ks[0] // 2 if self.padding == "same" else 0
        y = F.conv2d(x.permute(0, 3, 1, 2), w, self.weights["bias"], padding=pad)
        return y.permute(0, 2, 3, 1)


class Conv1D(Layer):  # defined by the reference's helpers, never instantiated by the graphs it builds
    prefix = "conv1d"

    def __init__(self, *a, **kw):
        raise NotImplementedError("Conv1D is not used by the reference's networks")


class Conv2DTranspose(Layer):
    prefix = "conv2d_transpose"

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", name=None, **kw):
        super().__init__(name)
        self.filters, self.ks, self.strides, self.padding = int(filters), tuple(kernel_size), tuple(strides), padding
        assert self.padding == "valid" and self.ks == self.strides, "only the k == stride 'valid' case is restated"

    def build(self, s):
        self._add("kernel", self.ks + (self.filters, s[-1]), 0.0)   # (kh, kw, out, in)
        self._add("bias", (self.filters,), 0.0)

    def out_shape(self, s):
        return (s[0], s[1] * self.strides[0], s[2] * self.strides[1], self.filters)

    def forward(self, x):
        w = self.weights["kernel"].permute(3, 2, 0, 1)   # torch: (in, out, kh, kw)
        y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, self.weights["bias"], stride=self.strides)
        return y.permute(0, 2, 3, 1)


class BatchNormalization(Layer):
    prefix = "batch_normalization"

    def __init__(self, axis=-1, epsilon=1e-3, name=None, **kw):
        super().__init__(name)
        assert axis == -1
        self.eps = epsilon

    def build(self, s):
        c = s[-1]
        self._add("gamma", (c,), 1.0)
        self._add("beta", (c,), 0.0)
        self._add("moving_mean", (c,), 0.0, trainable=False)
        self._add("moving_variance", (c,), 1.0, trainable=False)

    def forward(self, x):  # learning phase 0: the moving statistics
        w = self.weights
        return (x - w["moving_mean"]) / torch.sqrt(w["moving_variance"] + self.eps) * w["gamma"] + w["beta"]


class Activation(Layer):
    prefix = "activation"

    def __init__(self, activation, name=None, **kw):
        super().__init__(name)
        self.fn = {"relu": torch.relu, "tanh": torch.tanh, "softmax": lambda t: torch.softmax(t, dim=-1),
                   "linear": lambda t: t}[activation]

    def forward(self, x):
        return self.fn(x)


class Dropout(Layer):
    prefix = "dropout"

    def __init__(self, rate, name=None, **kw):
        super().__init__(name)

    def forward(self, x):  # learning phase 0
        return x


class MaxPooling2D(Layer):
    prefix = "max_pooling2d"

    def __init__(self, pool_size=(2, 2), name=None, **kw):
        super().__init__(name)
        assert tuple(pool_size) == (2, 2)

    def out_shape(self, s):
        return (s[0], s[1] // 2, s[2] // 2, s[3])

    def forward(self, x):
        return F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)


class Flatten(Layer):
    prefix = "flatten"

    def out_shape(self, s):
        return (s[0], int(np.prod(s[1:])))

    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class Reshape(Layer):
    prefix = "reshape"

    def __init__(self, target_shape, name=None, **kw):
        super().__init__(name)
        self.target = tuple(target_shape)

    def out_shape(self, s):
        return (None,) + self.target

    def forward(self, x):
        return x.reshape((x.shape[0],) + self.target)


class Lambda(Layer):
    prefix = "lambda"

    def __init__(self, function, name=None, **kw):
        super().__init__(name)
        self.function = function

    def out_shape(self, s):
        return None

    def forward(self, x):
        return self.function(x)


class _Merge(Layer):
    def __call__(self, xs):
        return Sym(self.forward, tuple(xs), max((x.shape for x in xs if x.shape), key=len, default=None), layer=self)

    @staticmethod
    def _align(vals):  # keras.layers.merge._Merge.call: lower-rank inputs get axes inserted at position 1
        r = max(v.dim() for v in vals)
        out = []
        for v in vals:
            while v.dim() < r:
                v = v.unsqueeze(1)
            out.append(v)
        return out


class _Multiply(_Merge):
    prefix = "multiply"

    def forward(self, *vals):
        vals = self._align(vals)
        y = vals[0]
        for v in vals[1:]:
            y = y * v
        return y


class _Add(_Merge):
    prefix = "add"

    def forward(self, *vals):
        vals = self._align(vals)
        y = vals[0]
        for v in vals[1:]:
            y = y + v
        return y


class _Concatenate(_Merge):
    prefix = "concatenate"

    def __init__(self, axis=-1, name=None):
        super().__init__(name)
        self.axis = axis

    def __call__(self, xs):
        s = list(xs[0].shape)
        s[self.axis] = sum(x.shape[self.axis] for x in xs)
        return Sym(self.forward, tuple(xs), tuple(s), layer=self)

    def forward(self, *vals):
        return torch.cat(vals, dim=self.axis)


def multiply(xs, name=None): return _Multiply(name)(xs)
def add(xs, name=None): return _Add(name)(xs)
def concatenate(xs, axis=-1, name=None): return _Concatenate(axis, name)(xs)


def Input(shape=None, name=None, tensor=None):
    if tensor is not None:  # an existing tensor wrapped as a model input: the graph node itself
        return tensor
    return Sym(None, (), (None,) + tuple(shape), name=name if name is not None else _auto_name("input"))


class Model:
    def __init__(self, inputs, outputs, name=None):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs, self.name = outputs, name
        self.layers = []  # in creation (= first use) order, like keras' topological layer list for these graphs
        seen, order = set(), []

        def walk(n):
            stack, post = [n], []
            while stack:
                cur = stack.pop()
                if id(cur) in seen:
                    continue
                seen.add(id(cur))
                post.append(cur)
                stack.extend(a for a in cur.args if isinstance(a, Sym))
            return post

        for n in walk(outputs):
            if n.layer is not None and n.layer not in order:
                order.append(n.layer)
        self.layers = sorted(order, key=lambda l: l._serial)

    # weights ---------------------------------------------------------------------------------------------
    def named_weights(self):
        """[(layer name, weight name, tensor)] in layer order."""
        return [(l.name, w, t) for l in self.layers for w, t in l.weights.items()]

    @property
    def trainable_weights(self):
        return [t for l in self.layers for w, t in l.weights.items() if w not in l.non_trainable]

    def set_named_weights(self, values):
        """values: {(layer, weight) or 'layer/weight': array}."""
        for l in self.layers:
            for w, t in l.weights.items():
                v = values[l.name + "/" + w] if (l.name + "/" + w) in values else values[(l.name, w)]
                v = torch.as_tensor(np.asarray(v), dtype=DT)
                assert tuple(v.shape) == tuple(t.shape), (l.name, w, tuple(v.shape), tuple(t.shape))
                with torch.no_grad():
                    t.copy_(v)

    # evaluation ------------------------------------------------------------------------------------------
    def __call__(self, xs):
        xs = list(xs) if isinstance(xs, (list, tuple)) else [xs]

        def run(*vals):
            return evaluate([self.outputs], dict(zip(self.inputs, vals)))[0]

        return Sym(run, tuple(xs), self.outputs.shape)

    def predict(self, xs, batch_size=None):
        xs = xs if isinstance(xs, (list, tuple)) else [xs]
        with torch.no_grad():
            y = evaluate([self.outputs], {p: torch.as_tensor(np.asarray(x), dtype=DT) for p, x in zip(self.inputs, xs)})[0]
        return y.numpy()

    def summary(self):
        pass


# every layer gets a creation serial so Model can list layers in creation order
_serial = [0]
_orig_init = Layer.__init__


def _init_with_serial(self, *a, **kw):
    _orig_init(self, *a, **kw)
    _serial[0] += 1
    self._serial = _serial[0]


Layer.__init__ = _init_with_serial


# --------------------------------------------------------------------------------------------------------------
# backend (`import keras.backend as K`), optimizers, the `tf` names the graph code touches
# --------------------------------------------------------------------------------------------------------------
class _Updates:
    def __init__(self, opt, params, loss):
        self.opt, self.params, self.loss = opt, list(params), loss


class Adam:
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0.0, **kw):
        self.lr, self.beta_1, self.beta_2 = lr, beta_1, beta_2
        self.epsilon = 1e-7 if epsilon is None else epsilon   # K.epsilon()
        assert decay == 0.0
        self.iterations = 0
        self.m = self.v = None

    def get_updates(self, *args, **kw):
        # keras 2.0 / 2.1: get_updates(params, constraints, loss); keras >= 2.1.3: get_updates(loss, params)
        if len(args) == 3:
            params, _, loss = args
        else:
            loss, params = args
        return _Updates(self, params, loss)

    def apply(self, params, grads):
        if self.m is None:
            self.m = [torch.zeros_like(p) for p in params]
            self.v = [torch.zeros_like(p) for p in params]
        self.iterations += 1
        t = self.iterations
        lr_t = self.lr * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        with torch.no_grad():
            for p, g, m, v in zip(params, grads, self.m, self.v):
                m.mul_(self.beta_1).add_((1.0 - self.beta_1) * g)
                v.mul_(self.beta_2).add_((1.0 - self.beta_2) * g * g)
                p.sub_(lr_t * m / (torch.sqrt(v) + self.epsilon))


class _Function:
    def __init__(self, inputs, outputs, updates=None):
        self.inputs, self.outputs, self.updates = list(inputs), list(outputs), updates
        self.last_grads = None

    def __call__(self, values):
        feed = {p: torch.as_tensor(np.asarray(v), dtype=DT) for p, v in zip(self.inputs, values)}
        nodes = list(self.outputs) + ([self.updates.loss] if self.updates is not None else [])
        vals = evaluate(nodes, feed)
        if self.updates is not None:
            params = self.updates.params
            grads = torch.autograd.grad(vals[-1], params, allow_unused=True)
            grads = [torch.zeros_like(p) if g is None else g for p, g in zip(params, grads)]
            self.last_grads = [g.detach().clone() for g in grads]
            self.updates.opt.apply(params, grads)
            vals = vals[:-1]
        return [np.asarray(v.detach().numpy()) for v in vals]


class _Backend:
    @staticmethod
    def epsilon(): return 1e-7

    @staticmethod
    def placeholder(shape=None, **kw): return Sym(None, (), tuple(shape) if shape else None, name=_auto_name("placeholder"))

    @staticmethod
    def mean(x, axis=None): return Sym(lambda a: a.mean() if axis is None else a.mean(dim=axis), (x,))

    @staticmethod
    def sum(x, axis=None):
        return Sym(lambda a: a.sum() if axis is None else a.sum(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis), (x,))

    @staticmethod
    def square(x): return Sym(lambda a: a * a, (x,))

    @staticmethod
    def sqrt(x): return Sym(lambda a: torch.sqrt(torch.clamp(a, min=0.0)), (x,))   # K.sqrt clips at zero

    @staticmethod
    def abs(x): return Sym(torch.abs, (x,))

    @staticmethod
    def flatten(x): return Sym(lambda a: a.reshape(-1), (x,))

    @staticmethod
    def greater_equal(x, y): return Sym(lambda a: a >= y, (x,))

    @staticmethod
    def cast(x, dtype): return Sym(lambda a: a.to(DT), (x,))

    @staticmethod
    def gradients(y, xs):
        def grad_of(x):
            def g(yv, xv):
                return torch.autograd.grad(yv.sum(), xv, create_graph=True)[0]
            return Sym(g, (y, x))
        return [grad_of(x) for x in xs]

    @staticmethod
    def function(inputs, outputs, updates=None): return _Function(inputs, outputs, updates)


K = _Backend()


class _TF:
    float32 = "float32"


tf = _TF()


def namespace():
    """The names the reference's network / graph code expects at module level."""
    reset_names()
    return dict(K=K, tf=tf, np=np, Model=Model, Input=Input, Dense=Dense, Conv1D=Conv1D, Conv2D=Conv2D,
                Conv2DTranspose=Conv2DTranspose, BatchNormalization=BatchNormalization, Activation=Activation,
                Dropout=Dropout, MaxPooling2D=MaxPooling2D, Flatten=Flatten, Reshape=Reshape, Lambda=Lambda,
                multiply=multiply, add=add, concatenate=concatenate, Adam=Adam)
