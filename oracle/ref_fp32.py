"""CPU oracle, float32 flavour: the MLP training step of the reference restated in SINGLE precision.

TEST INFRASTRUCTURE ONLY (same rules as oracle/ref_numpy.py: imported by tests/ only).

Why it exists.  The reference is a float64 computation from its second step on (gradients are
always float64 and `param += step` promotes the float32 parameters, /root/reference/core/
tensor.py:16-21, model.py:59-61).  The B200 engine's production path keeps float32 parameters,
gradients and Adam state.  north_star asks for "the loss trajectory over 100 steps within 1e-4" on
that path; this module shows what plain float32 arithmetic can hold (<= 2e-6 against the golden
trajectories recorded from the real reference, tests/test_oracle_golden.py) and serves as the
per-step bisection aid for the engine (parameters / Adam state after step k).

What it restates (citations: /root/reference/...):
  examples/mnist/run.py:78-84      zero_grad -> forward -> loss -> backward -> step
  utils/data_iterator.py:22-34     np.random.shuffle(idx) is drawn BEFORE the lazy Xavier draws of
                                   core/layers.py:45-46,53-56 (the iterator runs first)
  core/layers.py:43-49, 97-98      x @ w + b ; ReLU = clip(x, 0) with ReLU'(0) = 1 (ops.py:336-343)
  core/losses.py:24-32             batch-GLOBAL max and normaliser; closed-form gradient
                                   dL/dz = e/S - (1/m) * y * e / sum_j(y e)   (= p - y/m for one-hot y)
  core/optimizer.py:50-79          Adam; the coefficients 1-b1, 1-b2, 1-b1^t, 1-b2^t are Python
                                   floats (double) in the reference and are formed in double here
                                   before they meet the float32 vectors

Parity status: PINNED to tests/golden/mnist_traj.npz and tests/golden/mnist_learn_traj.npz (both
recorded by oracle/make_golden.py from the real reference).
"""
import numpy as np

F32 = np.float32


def xavier_uniform_f32(num_in, num_out):
    """core/initializer.py:83-86: float64 uniform draw from the global RNG, stored as float32"""
    bound = 1.0 * np.sqrt(6.0 / (num_in + num_out))
    return np.random.uniform(low=-bound, high=bound, size=[num_in, num_out]).astype(F32)


class AdamF32(object):
    """optimizer.py:50-79 on a list of float32 arrays; coefficients formed in double"""

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, epsilon
        self.t, self.m, self.v = 0, None, None

    def step(self, params, grads):
        if self.m is None:
            self.m = [np.zeros_like(p) for p in params]
            self.v = [np.zeros_like(p) for p in params]
        self.t += 1
        c1, c2 = F32(1.0 - self.b1), F32(1.0 - self.b2)
        bc1, bc2 = F32(1.0 - self.b1 ** self.t), F32(1.0 - self.b2 ** self.t)
        lr, eps = F32(self.lr), F32(self.eps)
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m += c1 * (g - m)
            v += c2 * (g * g - v)
            p += -lr * (m / bc1) / (np.sqrt(v / bc2) + eps)


class MLPF32(object):
    """Dense/ReLU stack + global-softmax cross-entropy + Adam, everything float32"""

    def __init__(self, widths, lr=1e-3):
        self.widths = list(widths)
        self.w, self.b = None, None
        self.opt = AdamF32(lr=lr)

    def _init(self, num_in):
        dims = [num_in] + self.widths
        self.w, self.b = [], []
        for i in range(len(self.widths)):          # draw order: layer by layer, w then b
            self.w.append(xavier_uniform_f32(dims[i], dims[i + 1]))
            self.b.append(np.zeros((1, dims[i + 1]), F32))

    def params(self):
        return [a for pair in zip(self.w, self.b) for a in pair]

    def loss_and_grads(self, x, labels):
        x = np.asarray(x, F32)
        y = np.asarray(labels, F32)
        if self.w is None:
            self._init(x.shape[1])
        acts, pre = [x], []
        h = x
        n = len(self.w)
        for i in range(n):
            z = h @ self.w[i] + self.b[i]
            pre.append(z)
            h = np.maximum(z, F32(0)) if i + 1 < n else z
            acts.append(h)
        m = x.shape[0]
        e = np.exp(h - h.max())
        S = e.sum(dtype=F32)
        q = (e / S * y).sum(axis=1, dtype=F32)
        loss = -np.log(q).sum(dtype=F32) / F32(m)
        ye = y * e
        dz = e / S - ye / ye.sum(axis=1, keepdims=True, dtype=F32) / F32(m)
        grads = [None] * (2 * n)
        for i in reversed(range(n)):
            grads[2 * i] = acts[i].T @ dz
            grads[2 * i + 1] = dz.sum(axis=0, keepdims=True, dtype=F32)
            if i > 0:
                dz = (dz @ self.w[i].T) * (pre[i - 1] >= 0)
        return float(loss), grads

    def train_step(self, x, labels):
        loss, grads = self.loss_and_grads(x, labels)
        self.opt.step(self.params(), grads)
        return loss


def mnist_style_trajectory(x, onehot, widths=(200, 100, 70, 30, 10), batch=128, steps=100, lr=1e-3):
    """run.py's loop with BatchIterator(shuffle=True): the caller seeds numpy's global RNG; the
    shuffle permutation is drawn first, the weights at the first forward"""
    idx = np.arange(len(x))
    np.random.shuffle(idx)
    xs, ys = x[idx], onehot[idx]
    mlp = MLPF32(widths, lr=lr)
    losses = []
    for start in range(0, len(xs), batch):
        losses.append(mlp.train_step(xs[start:start + batch], ys[start:start + batch]))
        if len(losses) == steps:
            break
    return np.array(losses), mlp


def learnable_mnist(n, seed=0, d_in=784, n_classes=10):
    """synthetic MNIST-shaped data a network can learn: each class has a fixed sparse template of
    bright pixels; a sample is its class template under multiplicative jitter plus full-range
    background noise, clipped to [0, 1], float32.  The reference's loss falls from 7.17 to 5.10
    over 100 Adam steps (random labels: it stays within 0.05), and -- unlike an easier data set
    that reaches the ln(batch) floor after 30 steps -- the trajectory is well conditioned: starting
    the reference in float64 instead of float32 (a 1e-7 relative perturbation of its first step)
    moves it by 1.3e-7, so a 1e-4 bound measures the arithmetic, not chaos."""
    rng = np.random.RandomState(seed)
    templates = (rng.rand(n_classes, d_in) < 0.2) * rng.uniform(0.2, 1.0, (n_classes, d_in))
    y = rng.randint(0, n_classes, n)
    x = templates[y] * rng.uniform(0.0, 1.0, (n, d_in)) + 1.0 * rng.rand(n, d_in)
    x = np.clip(x, 0.0, 1.0).astype(np.float32)
    return x, y, np.eye(n_classes)[y]
