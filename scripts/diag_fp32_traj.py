"""GPU diagnostic: the float32 engine's 100-step MNIST-MLP trajectory against the reference goldens,
and, step by step, its parameters against oracle/ref_fp32.py (the float32 numpy restatement) so a
drift can be bisected to the step and the parameter where it starts.  Prints JSON lines."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import ref_fp32  # noqa: E402
import ref_numpy as R  # noqa: E402


def run(kind, fused=False):
    from core.layers import Dense, ReLU
    from core.losses import SoftmaxCrossEntropyLoss
    from core.model import Model
    from core.nn import Net
    from core.optimizer import Adam
    from core.tensor import Tensor
    from utils.data_iterator import BatchIterator
    gold = np.load(os.path.join(ROOT, "tests", "golden", kind + ".npz"))["losses"]
    data = R.synthetic_mnist if kind == "mnist_traj" else ref_fp32.learnable_mnist
    x, y, onehot = data(12800, seed=0)
    # float32 numpy restatement, same seed
    np.random.seed(0)
    idx = np.arange(len(x))
    np.random.shuffle(idx)
    xs, ys = x[idx], onehot[idx]
    ref = ref_fp32.MLPF32([200, 100, 70, 30, 10])
    ref._init(784)           # weights drawn right after the shuffle, as in the reference
    # engine
    np.random.seed(0)
    widths = [200, 100, 70, 30, 10]
    layers = []
    for i, w in enumerate(widths):
        layers.append(Dense(w))
        if i + 1 < len(widths):
            layers.append(ReLU())
    net = Net(layers)
    model = Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=Adam(lr=1e-3))
    loss_layer = SoftmaxCrossEntropyLoss()
    rows = []
    for k, batch in enumerate(BatchIterator(batch_size=128)(Tensor(x), Tensor(onehot))):
        if k == 100:
            break
        if fused:
            loss = model.train_step(batch.inputs, batch.targets)
            grads = None
        else:
            model.zero_grad()
            loss = loss_layer.loss(model.forward(batch.inputs), batch.targets)
            loss.backward()
            grads = [p.grad.copy() for layer in net.get_parameters() for p in layer.values()]
            model.step()
        rl, rg = ref.loss_and_grads(xs[k * 128:(k + 1) * 128], ys[k * 128:(k + 1) * 128])
        ref.opt.step(ref.params(), rg)
        params = [p.values for layer in net.get_parameters() for p in layer.values()]
        gerr = -1.0 if grads is None else max(
            float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) for a, b in zip(grads, rg))
        perr = [float(np.max(np.abs(a - b))) for a, b in zip(params, ref.params())]
        rows.append(dict(step=k, loss=float(loss.values), d_gold=abs(float(loss.values) - gold[k]),
                         d_ref32=abs(float(loss.values) - rl), ref32_d_gold=abs(rl - gold[k]),
                         grad_rel=gerr, param_abs=perr))
    worst = max(rows, key=lambda r: r["d_gold"])
    print(json.dumps(dict(kind=kind, fused=fused, max_d_gold=worst["d_gold"], at=worst["step"],
                          max_ref32_d_gold=max(r["ref32_d_gold"] for r in rows))))
    for r in rows[:28] + rows[28::8]:
        print(json.dumps(r))


if __name__ == "__main__":
    for kind in ("mnist_traj", "mnist_learn_traj"):
        run(kind, fused="--fused" in sys.argv)
