"""GPU diagnostic: the fused small-MLP step against the eager step, teacher-forced (parameters and
Adam state copied from the eager model before every step) on the learnable MNIST-shaped data, so
the first step where the two differ by more than rounding -- and in which parameter -- shows."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import ref_fp32  # noqa: E402
import core._backend as be  # noqa: E402
from core.layers import Dense, ReLU  # noqa: E402
from core.losses import SoftmaxCrossEntropyLoss  # noqa: E402
from core.model import Model  # noqa: E402
from core.nn import Net  # noqa: E402
from core.optimizer import Adam  # noqa: E402
from core.tensor import Tensor  # noqa: E402


def build():
    np.random.seed(0)
    dims = [784, 200, 100, 70, 30, 10]
    layers = []
    for i in range(5):
        layers.append(Dense(dims[i + 1], num_in=dims[i]))
        if i < 4:
            layers.append(ReLU())
    net = Net(layers)
    return net, Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=Adam(lr=1e-3))


def main():
    x, y, onehot = ref_fp32.learnable_mnist(12800, seed=0)
    np.random.seed(0)
    idx = np.arange(len(x))
    np.random.shuffle(idx)
    xs, ys = x[idx], onehot[idx]
    net_a, a = build()
    net_b, b = build()
    for k in range(40):
        xb, yb = Tensor(xs[k * 128:(k + 1) * 128]), Tensor(ys[k * 128:(k + 1) * 128])
        if a._arena is not None and b._arena is not None:
            be.copy_into(b._arena["p"], a._arena["p"])
            for sb, sa in zip(b.optimizer._state, a.optimizer._state):
                be.copy_into(sb, sa)
            b.optimizer._t = a.optimizer._t
            for p in b._arena["params"]:
                p._touch()
        a.zero_grad()
        la = a.loss.loss(a.forward(xb), yb)
        la.backward()
        ga = [p.grad.copy() for p in a._param_list()]
        a.step()
        lb = b.train_step(xb, yb)
        pa = [p.values for p in a._param_list()]
        pb = [p.values for p in b._param_list()]
        d = ["%.1e" % (np.max(np.abs(u - v)) / np.max(np.abs(u))) for u, v in zip(pa, pb)]
        gmin = ["%.1e" % np.min(np.abs(g[g != 0])) if np.any(g != 0) else "0" for g in ga]
        print(k, "%.7f %.7f" % (float(la.values), float(lb.values)), d, "min|g|", gmin[:4])


if __name__ == "__main__":
    main()
