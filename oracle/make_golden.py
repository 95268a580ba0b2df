"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Outputs (small, committed):
  tests/golden/ops.npz          every case of tests/op_cases.py: forward value + input gradients
  tests/golden/optimizers.npz   3 steps of each optimiser's _compute_step on seeded flat gradients
  tests/golden/mnist_traj.npz   100-step loss trajectory of the examples/mnist MLP (784-200-100-
                                70-30-10, Adam 1e-3, batch 128, np.random.seed(0), synthetic
                                MNIST-shaped data) + first-step gradients' norms + final params'
                                checksums
  tests/golden/mnist_learn_traj.npz  the same loop on LEARNABLE synthetic data (oracle/ref_fp32.py
                                learnable_mnist: class templates + noise): the loss falls by > 1
                                over the 100 steps, so the 1e-4 trajectory bound discriminates
  tests/golden/ref_checkpoint.pkl    a checkpoint in the reference's own format (pickled Net) +
  tests/golden/ref_checkpoint_io.npz an input batch and the forward values it gives
  tests/golden/mlp_step.npz     3 Adam steps of a small wide-style MLP (4 x Dense(64), fp32
                                one-hot labels): losses, first-step gradients, final parameters
"""
import os
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, REF)                      # `core` / `utils` must resolve to the reference
sys.path.insert(1, os.path.join(ROOT, "tests"))
sys.path.insert(2, HERE)

import numpy as np  # noqa: E402

import core  # noqa: E402

assert os.path.abspath(core.__file__).startswith(REF), "core resolved to %s" % core.__file__

import core.ops as ops  # noqa: E402
from core.layers import Dense, ReLU  # noqa: E402
from core.losses import SoftmaxCrossEntropyLoss  # noqa: E402
from core.model import Model  # noqa: E402
from core.nn import Net  # noqa: E402
import core.optimizer as ref_opt  # noqa: E402
from core.tensor import Tensor  # noqa: E402
from utils.data_iterator import BatchIterator  # noqa: E402

import op_cases  # noqa: E402
import ref_numpy  # noqa: E402
import ref_fp32  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def gen_ops():
    store = {}
    ce = SoftmaxCrossEntropyLoss()
    for case in op_cases.CASES:
        out, grads = op_cases.run_reference_style(case, Tensor, ops, ce.loss)
        store[case["name"] + "/out"] = out
        for i, g in enumerate(grads):
            store[case["name"] + "/g%d" % i] = g
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **store)
    print("ops.npz: %d cases" % len(op_cases.CASES))


OPTIMIZERS = [
    ("sgd", lambda: ref_opt.SGD(lr=0.05)),
    ("adam", lambda: ref_opt.Adam(lr=1e-3)),
    ("rmsprop", lambda: ref_opt.RMSProp(lr=0.01, momentum=0.5)),
    ("momentum", lambda: ref_opt.Momentum(lr=0.02, momentum=0.9)),
    ("adagrad", lambda: ref_opt.Adagrad(lr=0.1)),
    ("adadelta", lambda: ref_opt.Adadelta(lr=1.0)),
]


def gen_optimizers():
    store = {}
    rng = np.random.RandomState(7)
    grads = [rng.standard_normal(1000) for _ in range(3)]
    store["grads"] = np.stack(grads)
    for name, make in OPTIMIZERS:
        opt = make()
        store[name] = np.stack([np.array(opt._compute_step(g.copy())) for g in grads])
    np.savez_compressed(os.path.join(OUT, "optimizers.npz"), **store)
    print("optimizers.npz")


def gen_mnist_traj(learnable=False):
    np.random.seed(0)
    if learnable:
        x, y, onehot = ref_fp32.learnable_mnist(12800, seed=0)
    else:
        x, y, onehot = ref_numpy.synthetic_mnist(12800, seed=0)
    train_x, train_y = Tensor(x), Tensor(onehot)
    net = Net([Dense(200), ReLU(), Dense(100), ReLU(), Dense(70), ReLU(), Dense(30), ReLU(), Dense(10)])
    model = Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=ref_opt.Adam(lr=1e-3))
    loss_layer = SoftmaxCrossEntropyLoss()
    losses = []
    first_grad_norms = None
    for batch in BatchIterator(batch_size=128)(train_x, train_y):
        model.zero_grad()
        pred = model.forward(batch.inputs)
        loss = loss_layer.loss(pred, batch.targets)
        loss.backward()
        if first_grad_norms is None:
            first_grad_norms = np.array([float(np.linalg.norm(p.grad))
                                         for layer in net.get_parameters() for p in layer.values()])
        model.step()
        losses.append(float(loss.values))
        if len(losses) == 100:
            break
    sums = np.array([float(np.sum(p.values)) for layer in net.get_parameters() for p in layer.values()])
    name = "mnist_learn_traj.npz" if learnable else "mnist_traj.npz"
    np.savez_compressed(os.path.join(OUT, name), losses=np.array(losses),
                        first_grad_norms=first_grad_norms, final_param_sums=sums)
    print("%s: loss %.6f -> %.6f" % (name, losses[0], losses[-1]))


def gen_mlp_step():
    np.random.seed(0)
    rng = np.random.RandomState(0)
    B, D = 32, 64
    x = rng.rand(B, D).astype(np.float32)
    labels = np.eye(D, dtype=np.float32)[rng.randint(0, D, B)]
    net = Net([Dense(D), ReLU(), Dense(D), ReLU(), Dense(D), ReLU(), Dense(D)])
    model = Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=ref_opt.Adam(lr=1e-3))
    loss_layer = SoftmaxCrossEntropyLoss()
    store = {}
    losses = []
    for it in range(3):
        model.zero_grad()
        loss = loss_layer.loss(model.forward(Tensor(x)), Tensor(labels))
        loss.backward()
        if it == 0:
            k = 0
            for layer in net.get_parameters():
                for p in layer.values():
                    store["grad%d" % k] = np.array(p.grad)
                    k += 1
        model.step()
        losses.append(float(loss.values))
    k = 0
    for layer in net.get_parameters():
        for p in layer.values():
            store["param%d" % k] = np.array(p.values)
            k += 1
    store["losses"] = np.array(losses)
    np.savez_compressed(os.path.join(OUT, "mlp_step.npz"), **store)
    print("mlp_step.npz: losses", losses)


def gen_ref_checkpoint():
    """a checkpoint written by the reference's own Model.save (model.py:18-21: the pickled Net),
    and the forward values it must reproduce once loaded"""
    np.random.seed(21)
    x = np.random.rand(5, 4).astype(np.float32)
    # the reference can only pickle a Net that has not run yet (after a forward the layers hold
    # tensors whose closures do not pickle), so: eager initialisation, save, then forward
    net = Net([Dense(3, num_in=4), ReLU(), Dense(2, num_in=3)])
    model = Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=ref_opt.Adam(lr=1e-3))
    model.save(os.path.join(OUT, "ref_checkpoint.pkl"))
    y = model.forward(Tensor(x)).values
    np.savez_compressed(os.path.join(OUT, "ref_checkpoint_io.npz"), x=x, y=np.array(y))


if __name__ == "__main__":
    gen_ops()
    gen_optimizers()
    gen_mnist_traj()
    gen_mnist_traj(learnable=True)
    gen_mlp_step()
    gen_ref_checkpoint()
