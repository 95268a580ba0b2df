"""CPU oracle: a numpy restatement of tinynn-autograd's forward/backward tensor-op path.

TEST INFRASTRUCTURE ONLY.  Nothing under core/, utils/ or tinynn-autograd_b200/ imports this
module; it is used by tests/ (as the checker), by __graft_entry__.smoke() (as the checker) and by
bench.py's cpu_baseline / --impl reference legs (as the CPU implementation being timed).

What it restates (all citations are /root/reference/...):
  core/tensor.py:13-171    Tensor, float64 gradients, the per-path recursive backward
  core/ops.py:32-344       every primitive's forward value and gradient closure, including the
                           un-broadcast loops each binary op carries
  core/layers.py:25-98     Dense (lazy Xavier-uniform init, x@w+b), ReLU = clip(x, 0)
  core/losses.py:24-32     SoftmaxCrossEntropyLoss with its batch-global max / normaliser
  core/optimizer.py:12-164 flatten -> _compute_step -> unflatten, SGD/Adam/RMSProp/Momentum/
                           Adagrad/Adadelta
  core/model.py:45-68      step (param += step, which promotes float32 params to float64) and
                           zero_grad

Parity status: PINNED.  tests/test_oracle_golden.py checks this module against
  (a) the known-answer vectors of the reference's own test/test_autograd.py, and
  (b) tests/golden/*.npz, produced by oracle/make_golden.py by importing and running the real
      reference from /root/reference in this container (script committed next to this file).
The arithmetic backend of the reference is numpy (pinned numpy==1.22.0 in requirements.txt,
2.3.x installed here); this restatement calls the same numpy ufuncs in the same order, so (b)
holds bit for bit on this machine.
"""
import numpy as np


# --------------------------------------------------------------------------------------------
# tensor + backward (tensor.py:13-171)
# --------------------------------------------------------------------------------------------
class RefTensor(object):

    def __init__(self, values, requires_grad=False, parents=None, dtype=None):
        self.values = np.asarray(values, dtype)
        self.requires_grad = bool(requires_grad)
        self.grad = np.zeros(self.values.shape) if self.requires_grad else None  # always float64
        self.parents = parents or []   # [(RefTensor, vjp)] for the inputs that require grad

    @property
    def shape(self):
        return self.values.shape

    def zero_grad(self):
        self.grad = np.zeros(self.values.shape)

    def assign(self, new_values):
        """the `values` setter (tensor.py:35-38): new storage, gradient dropped"""
        self.values = np.asarray(new_values)
        self.grad = None

    def backward(self, grad=None):
        """tensor.py:157-168 -- one recursive call per PATH through the graph (no visited set),
        accumulating into .grad of every node on the way, leaves and intermediates alike."""
        assert self.requires_grad, "Call backward() on a non-requires-grad tensor."
        g = np.array(1.0 if grad is None else grad)
        self.grad += g
        for parent, vjp in self.parents:
            parent.backward(vjp(g))

    # operator sugar so test code reads like the reference's
    def __add__(self, o): return add(self, lift(o))
    def __radd__(self, o): return add(lift(o), self)
    def __sub__(self, o): return sub(self, lift(o))
    def __rsub__(self, o): return sub(lift(o), self)
    def __mul__(self, o): return mul(self, lift(o))
    def __rmul__(self, o): return mul(lift(o), self)
    def __truediv__(self, o): return div(self, lift(o))
    def __rtruediv__(self, o): return div(lift(o), self)
    def __pow__(self, o): return power(self, lift(o))
    def __matmul__(self, o): return matmul(self, lift(o))
    def __neg__(self): return neg(self)
    def __getitem__(self, k): return getitem(self, k)
    def sum(self, axis=None): return reduce_sum(self, axis)
    def max(self, axis=None): return reduce_max(self, axis)
    def min(self, axis=None): return reduce_min(self, axis)
    @property
    def T(self): return transpose(self, None)


def lift(x):
    return x if isinstance(x, RefTensor) else RefTensor(x)


def _node(values, inputs_and_vjps):
    parents = [(t, f) for t, f in inputs_and_vjps if t.requires_grad]
    return RefTensor(values, requires_grad=any(t.requires_grad for t, _ in inputs_and_vjps),
                     parents=parents)


def unbroadcast(grad, shape):
    """the loop pair every binary op repeats (ops.py:41-46): sum away the leading axes the
    operand did not have, then sum (keepdims) over its size-1 axes"""
    for _ in range(grad.ndim - len(shape)):
        grad = grad.sum(axis=0)
    for axis, extent in enumerate(shape):
        if extent == 1:
            grad = grad.sum(axis=axis, keepdims=True)
    return grad


# --------------------------------------------------------------------------------------------
# primitives (ops.py:32-344)
# --------------------------------------------------------------------------------------------
def add(a, b):                                                      # ops.py:32-58
    return _node(a.values + b.values,
                 [(a, lambda g: unbroadcast(g, a.shape)), (b, lambda g: unbroadcast(g, b.shape))])


def neg(a):                                                         # ops.py:293-299
    return _node(-a.values, [(a, lambda g: -g)])


def sub(a, b):                                                      # ops.py:61-62
    return add(a, neg(b))


def mul(a, b):                                                      # ops.py:65-90
    return _node(a.values * b.values,
                 [(a, lambda g: unbroadcast(g * b.values, a.shape)),
                  (b, lambda g: unbroadcast(g * a.values, b.shape))])


def div(a, b):                                                      # ops.py:93-118
    return _node(a.values / b.values,
                 [(a, lambda g: unbroadcast(g / b.values, a.shape)),
                  (b, lambda g: unbroadcast(-g * a.values / b.values ** 2, b.shape))])


def power(a, b):                                                    # ops.py:121-147
    out = a.values ** b.values
    return _node(out,
                 [(a, lambda g: unbroadcast(g * b.values * a.values ** (b.values - 1), a.shape)),
                  (b, lambda g: unbroadcast(g * (np.log(a.values) * out), b.shape))])


def matmul(a, b):                                                   # ops.py:150-163
    return _node(a.values @ b.values,
                 [(a, lambda g: g @ b.values.T), (b, lambda g: a.values.T @ g)])


def maximum(a, b):                                                  # ops.py:166-188
    return _node(np.maximum(a.values, b.values),
                 [(a, lambda g: unbroadcast(g * (a.values >= b.values), a.shape)),
                  (b, lambda g: unbroadcast(g * (b.values > a.values), b.shape))])


def minimum(a, b):                                                  # ops.py:191-213
    return _node(np.minimum(a.values, b.values),
                 [(a, lambda g: unbroadcast(g * (a.values <= b.values), a.shape)),
                  (b, lambda g: unbroadcast(g * (b.values < a.values), b.shape))])


def exp(a):                                                         # ops.py:216-222
    out = np.exp(a.values)
    return _node(out, [(a, lambda g: out * g)])


def log(a):                                                         # ops.py:243-249
    return _node(np.log(a.values), [(a, lambda g: g / a.values)])


def reduce_max(a, axis=None):                                       # ops.py:225-231
    return _node(np.max(a.values, axis=axis),
                 [(a, lambda g: g * (a.values.max(axis=axis, keepdims=1) == a.values))])


def reduce_min(a, axis=None):                                       # ops.py:234-240
    return _node(np.min(a.values, axis=axis),
                 [(a, lambda g: g * (a.values.min(axis=axis, keepdims=1) == a.values))])


def reduce_sum(a, axis=None):                                       # ops.py:252-265
    def vjp(g):
        if axis is None:
            return g * np.ones_like(a.values)
        return np.repeat(np.expand_dims(g, axis), a.values.shape[axis], axis)
    return _node(a.values.sum(axis=axis), [(a, vjp)])


def transpose(a, axes=None):                                        # ops.py:268-279
    order = list(reversed(range(a.values.ndim))) if axes is None else list(axes)
    return _node(a.values.transpose(axes), [(a, lambda g: g.transpose(np.argsort(order)))])


def getitem(a, key):                                                # ops.py:282-290
    def vjp(g):
        full = np.zeros_like(a.values)
        full[key] = g          # assignment: duplicate indices are not accumulated
        return full
    return _node(a.values[key], [(a, vjp)])


def reshape(a, newshape):                                           # ops.py:302-309
    old = a.values.shape
    return _node(a.values.reshape(newshape), [(a, lambda g: g.reshape(old))])


def flatten(a):                                                     # ops.py:324-330
    old = a.values.shape
    return _node(a.values.ravel(), [(a, lambda g: g.reshape(old))])


def pad(a, pad_width, mode="constant"):                             # ops.py:312-321
    out = np.pad(a.values, pad_width=pad_width, mode=mode)
    window = tuple(slice(before, size - after)
                   for size, (before, after) in zip(out.shape, pad_width))
    return _node(out, [(a, lambda g: g[window])])


def clip(a, lo=None, hi=None, keep_override=None):                  # ops.py:333-344
    """keep_override is NOT part of the reference: a parity test may hand in the mask the device
    computed when a pre-activation sits within rounding distance of the kink (the two sides then
    differ by the discontinuity of ReLU', not by arithmetic); the test bounds how many entries and
    how close to zero they are (tests/test_gpu_wide_fullsize.py)."""
    keep = np.ones(a.values.shape, dtype=bool)   # built eagerly in the forward pass
    if lo is not None:
        keep &= a.values >= lo
    if hi is not None:
        keep &= a.values <= hi
    if keep_override is not None:
        keep = np.asarray(keep_override, dtype=bool)
    return _node(a.values.clip(lo, hi), [(a, lambda g: g * keep)])


# --------------------------------------------------------------------------------------------
# layers / loss (layers.py, losses.py)
# --------------------------------------------------------------------------------------------
def xavier_uniform(shape):
    """initializer.py:83-86 + 17-19: np.random.uniform in float64, stored as float32"""
    bound = 1.0 * np.sqrt(6.0 / (shape[0] + shape[1]))
    return RefTensor(np.random.uniform(low=-bound, high=bound, size=shape), True, dtype=np.float32)


class RefDense(object):
    """layers.py:25-57"""

    def __init__(self, num_out, num_in=None):
        self.num_out, self.w, self.b = num_out, None, None
        if num_in is not None:
            self._init(num_in)

    def _init(self, num_in):
        self.w = xavier_uniform([num_in, self.num_out])
        self.b = RefTensor(np.full(shape=[1, self.num_out], fill_value=0.0), True, dtype=np.float32)

    def params(self):
        return [self.w, self.b]

    def forward(self, x):
        if self.w is None:
            self._init(x.shape[1])
        return matmul(x, self.w) + self.b


class RefReLU(object):
    """layers.py:92-98: ops.clip(x, 0.0)"""

    keep_override = None   # see clip()

    def params(self):
        return []

    def forward(self, x):
        return clip(x, 0.0, keep_override=self.keep_override)


def softmax_cross_entropy(logits, labels):
    """losses.py:24-32, expression by expression (note: .max() and .sum() have no axis)"""
    m = logits.shape[0]
    exps = exp(logits - logits.max())
    p = exps / exps.sum()
    nll = -log((p * lift(labels)).sum(1))
    return nll.sum() / m


def cross_entropy_from_class_indices(logits, class_idx):
    """The same loss and dL/dz for targets given as class indices (what the engine's fused kernels
    read when a batch's one-hot rows were never written, tnn_ce_loss / tnn_ce_bwd `labels_dev`),
    in plain numpy at the logits' precision: with y_i = one_hot(c_i) the row sum of losses.py:28,
    sum_j p_ij y_ij, has the single non-zero term p_{i,c_i}, and adding exact zeros to it changes
    no bit.  Returns (loss, dz); an index outside [0, C) behaves like an all-zero row (q_i = 0)."""
    z = np.asarray(logits)
    m, C = z.shape
    e = np.exp(z - z.max())
    p = e / e.sum()
    idx = np.asarray(class_idx)
    ok = (idx >= 0) & (idx < C)
    q = np.zeros(m, z.dtype)
    q[ok] = p[np.arange(m)[ok], idx[ok]]
    with np.errstate(divide="ignore"):
        loss = (-np.log(q)).sum() / m
    dz = p.copy()
    rows = np.arange(m)[ok]
    # d/dz of -(1/m) sum_i ln q_i  with q_i = p_{i,c_i}:  p_ij * (#valid rows)/m ... spelled out as the
    # chain rule the autograd graph of losses.py:24-32 applies: dL/dp_{i,c_i} = -1/(m q_i), then the
    # quotient and exp/max nodes.  Closed form for one-hot rows: dz = k p - y/m with k = (#rows with a
    # valid index)/m, because sum_ij (dL/dp_ij) p_ij = -k.
    k = z.dtype.type(ok.sum()) / z.dtype.type(m)
    dz = k * p
    dz[rows, idx[ok]] -= z.dtype.type(1) / z.dtype.type(m)
    return loss, dz


# --------------------------------------------------------------------------------------------
# optimisers (optimizer.py)
# --------------------------------------------------------------------------------------------
class RefOptimizer(object):

    def compute_steps(self, params):
        """optimizer.py:12-35: concatenate every raveled gradient, one _step over the flat
        vector, slice it back into parameter shapes"""
        flat = np.concatenate([np.ravel(p.grad) for p in params])
        flat_step = self._step(flat)
        out, pos = [], 0
        for p in params:
            n = int(np.prod(p.shape))
            out.append(flat_step[pos:pos + n].reshape(p.shape))
            pos += n
        return out


class RefSGD(RefOptimizer):                                          # optimizer.py:41-47
    def __init__(self, lr):
        self.lr = lr

    def _step(self, g):
        return -self.lr * g


class RefAdam(RefOptimizer):                                         # optimizer.py:50-79
    def __init__(self, lr=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, epsilon
        self.t, self.m, self.v = 0, 0, 0

    def _step(self, g):
        self.t += 1
        self.m += (1.0 - self.b1) * (g - self.m)
        self.v += (1.0 - self.b2) * (g ** 2 - self.v)
        m_hat = self.m / (1 - self.b1 ** self.t)
        v_hat = self.v / (1 - self.b2 ** self.t)
        return -self.lr * m_hat / (v_hat ** 0.5 + self.eps)


class RefRMSProp(RefOptimizer):                                      # optimizer.py:82-112
    def __init__(self, lr=0.01, decay=0.99, momentum=0.0, epsilon=1e-8):
        self.lr, self.decay, self.momentum, self.eps = lr, decay, momentum, epsilon
        self.ms, self.mom = 0, 0

    def _step(self, g):
        self.ms += (1 - self.decay) * (g ** 2 - self.ms)
        self.mom = self.momentum * self.mom + self.lr * g / (self.ms + self.eps) ** 0.5
        return -self.mom


class RefMomentum(RefOptimizer):                                     # optimizer.py:115-128
    def __init__(self, lr, momentum=0.9):
        self.lr, self.momentum, self.acc = lr, momentum, 0

    def _step(self, g):
        self.acc = self.momentum * self.acc + g
        return -self.lr * self.acc


class RefAdagrad(RefOptimizer):                                      # optimizer.py:131-146
    def __init__(self, lr, epsilon=1e-8):
        self.lr, self.eps, self.G = lr, epsilon, 0

    def _step(self, g):
        self.G += g ** 2
        return -(self.lr / (self.G + self.eps) ** 0.5) * g


class RefAdadelta(RefOptimizer):                                     # optimizer.py:149-164
    def __init__(self, lr=1.0, decay=0.9, epsilon=1e-8):
        self.lr, self.decay, self.eps = lr, decay, epsilon
        self.Eg, self.delta = 0, 0

    def _step(self, g):
        self.Eg += (1 - self.decay) * (g ** 2 - self.Eg)
        std = (self.delta + self.eps) ** 0.5
        d = g * (std / (self.Eg + self.eps) ** 0.5)
        step = -self.lr * d
        self.delta += (1 - self.decay) * (d ** 2 - self.delta)
        return step


# --------------------------------------------------------------------------------------------
# model harness (nn.py, model.py, examples/mnist/run.py:78-84)
# --------------------------------------------------------------------------------------------
class RefMLP(object):

    def __init__(self, widths, optimizer):
        """widths = output width of each Dense; ReLU between them (run.py:59-69)"""
        self.layers = []
        for i, w in enumerate(widths):
            self.layers.append(RefDense(w))
            if i + 1 < len(widths):
                self.layers.append(RefReLU())
        self.optimizer = optimizer

    def params(self):
        return [p for layer in self.layers for p in layer.params() if p is not None]

    def forward(self, x):
        for layer in self.layers:
            x = layer.forward(x)
        return x

    def zero_grad(self):                                             # model.py:63-68
        for p in self.params():
            p.zero_grad()

    def step(self):                                                  # model.py:45-61
        params = self.params()
        for p, s in zip(params, self.optimizer.compute_steps(params)):
            p.assign(p.values + s)     # float32 + float64 -> float64, as in the reference

    def train_step(self, x, labels):
        """run.py:79-84; returns the loss value"""
        self.zero_grad()
        loss = softmax_cross_entropy(self.forward(lift(x)), labels)
        loss.backward()
        self.step()
        return loss.values


def synthetic_mnist(n, seed=0, d_in=784, n_classes=10):
    """the synthetic MNIST-shaped data BASELINE.md prescribes: U[0,1) float32 pixels, uniform
    integer labels, one-hot float64 labels via np.eye (run.py:27-28)"""
    rng = np.random.RandomState(seed)
    x = rng.rand(n, d_in).astype(np.float32)
    y = rng.randint(0, n_classes, n)
    return x, y, np.eye(n_classes)[y]
