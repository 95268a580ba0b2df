"""End-to-end parity of the training path (Dense/ReLU layers, fused CE, autograd sweep, flat
arenas, fused Adam) against goldens recorded from the real reference and against the oracle."""
import os

import numpy as np
import pytest

import op_cases
import ref_numpy as R

pytestmark = pytest.mark.gpu


def _build(widths, lr=1e-3):
    from core.layers import Dense, ReLU
    from core.losses import SoftmaxCrossEntropyLoss
    from core.model import Model
    from core.nn import Net
    from core.optimizer import Adam
    layers = []
    for i, w in enumerate(widths):
        layers.append(Dense(w))
        if i + 1 < len(widths):
            layers.append(ReLU())
    net = Net(layers)
    return net, Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=Adam(lr=lr)), SoftmaxCrossEntropyLoss()


def test_mnist_mlp_loss_trajectory(golden_dir):
    """examples/mnist/run.py's loop on synthetic MNIST-shaped data: 100-step loss trajectory within
    1e-4 of the reference's (north_star), same seed, same batches, float32 engine vs float64 ref"""
    from core.tensor import Tensor
    from utils.data_iterator import BatchIterator
    gold = np.load(os.path.join(golden_dir, "mnist_traj.npz"))
    np.random.seed(0)
    x, y, onehot = R.synthetic_mnist(12800, seed=0)
    train_x, train_y = Tensor(x), Tensor(onehot)
    net, model, loss_layer = _build([200, 100, 70, 30, 10])
    losses, first_norms = [], None
    for batch in BatchIterator(batch_size=128)(train_x, train_y):
        model.zero_grad()
        pred = model.forward(batch.inputs)
        loss = loss_layer.loss(pred, batch.targets)
        loss.backward()
        if first_norms is None:
            first_norms = np.array([float(np.linalg.norm(p.grad)) for layer in net.get_parameters()
                                    for p in layer.values()])
        model.step()
        losses.append(float(loss.values))
        if len(losses) == 100:
            break
    assert np.max(np.abs(np.array(losses) - gold["losses"])) <= 1e-4
    assert np.allclose(first_norms, gold["first_grad_norms"], rtol=1e-4, atol=1e-9)
    sums = np.array([float(np.sum(p.values)) for layer in net.get_parameters() for p in layer.values()])
    assert np.allclose(sums, gold["final_param_sums"], rtol=1e-3, atol=1e-3)


def test_wide_style_mlp_steps(golden_dir):
    """3 Adam steps of a 4 x Dense(64) MLP with fp32 one-hot labels: losses, first-step gradients
    and final parameters against the reference golden"""
    from core.tensor import Tensor
    gold = np.load(os.path.join(golden_dir, "mlp_step.npz"))
    np.random.seed(0)
    rng = np.random.RandomState(0)
    B, D = 32, 64
    x = rng.rand(B, D).astype(np.float32)
    labels = np.eye(D, dtype=np.float32)[rng.randint(0, D, B)]
    net, model, loss_layer = _build([D, D, D, D])
    losses = []
    for it in range(3):
        model.zero_grad()
        loss = loss_layer.loss(model.forward(Tensor(x)), Tensor(labels))
        loss.backward()
        if it == 0:
            k = 0
            for layer in net.get_parameters():
                for p in layer.values():
                    assert op_cases.rel_err(p.grad, gold["grad%d" % k]) <= 1e-5, k
                    k += 1
        model.step()
        losses.append(float(loss.values))
    assert np.max(np.abs(np.array(losses) - gold["losses"])) <= 1e-5
    k = 0
    for layer in net.get_parameters():
        for p in layer.values():
            assert op_cases.rel_err(p.values, gold["param%d" % k]) <= 1e-5, k
            k += 1


@pytest.mark.parametrize("cg", [1, 2])
def test_tensor_core_mlp_step_teacher_forced(cg):
    """a 4-layer 512-wide MLP at batch 1024 runs its 11 GEMMs on the tcgen05 path; one step is
    compared with the oracle started from the same parameters (teacher forcing)"""
    import core._backend as be
    from core.tensor import Tensor
    be.set_gemm_cta_group(cg)
    old = be.TC_MIN_MNK
    be.TC_MIN_MNK = 1 << 20
    try:
        np.random.seed(1)
        rng = np.random.RandomState(1)
        B, D = 1024, 512
        x = rng.rand(B, D).astype(np.float32)
        labels = np.eye(D, dtype=np.float32)[rng.randint(0, D, B)]
        net, model, loss_layer = _build([D, D, D, D])
        model.zero_grad()
        loss = loss_layer.loss(model.forward(Tensor(x)), Tensor(labels))
        # oracle with the very same initial parameters
        mlp = R.RefMLP([D, D, D, D], R.RefAdam(lr=1e-3))
        params = [p for layer in net.get_parameters() for p in layer.values()]
        k = 0
        for layer in mlp.layers:
            if isinstance(layer, R.RefDense):
                layer.w = R.RefTensor(params[k].values.copy(), True)
                layer.b = R.RefTensor(params[k + 1].values.copy(), True)
                k += 2
        rloss = R.softmax_cross_entropy(mlp.forward(R.lift(x)), labels)
        assert abs(float(loss.values) - float(rloss.values)) <= 1e-5
        loss.backward()
        rloss.backward()
        for p, rp in zip(params, mlp.params()):
            assert op_cases.rel_err(p.grad, rp.grad) <= 1e-5
        # the fused arena Adam, checked against the reference update applied to the engine's own
        # gradients (Adam's first step is -lr*g/(|g|+eps): for entries with |g| ~ eps = 1e-8 it
        # turns a 1e-8 absolute gradient difference into an O(lr) step difference, so feeding
        # both sides the same gradient is the meaningful comparison)
        p0 = [p.values.astype(np.float64) for p in params]
        g0 = [p.grad.astype(np.float64) for p in params]
        model.step()
        adam = R.RefAdam(lr=1e-3)
        flat_step = adam._step(np.concatenate([g.ravel() for g in g0]))
        pos = 0
        for p, a, g in zip(params, p0, g0):
            expect = a + flat_step[pos:pos + g.size].reshape(g.shape)
            pos += g.size
            assert op_cases.rel_err(p.values, expect) <= 1e-5
    finally:
        be.TC_MIN_MNK = old
        be.set_gemm_cta_group(0)


def test_generic_step_path_equals_fused():
    """Model.step through compute_step()/`param += step` (the reference's three stages) gives the
    same parameters as the fused arena kernel"""
    from core.tensor import Tensor
    rng = np.random.RandomState(2)
    x = rng.rand(16, 20).astype(np.float32)
    labels = np.eye(8, dtype=np.float32)[rng.randint(0, 8, 16)]
    results = []
    for generic in (False, True):
        np.random.seed(3)
        net, model, loss_layer = _build([12, 8])
        for _ in range(3):
            model.zero_grad()
            loss = loss_layer.loss(model.forward(Tensor(x)), Tensor(labels))
            loss.backward()
            if generic:
                model._step_generic()
            else:
                model.step()
        results.append([p.values.copy() for layer in net.get_parameters() for p in layer.values()])
    for a, b in zip(*results):
        assert op_cases.rel_err(a, b) <= 1e-6


def test_save_load_roundtrip(tmp_path):
    from core.tensor import Tensor
    np.random.seed(4)
    net, model, _ = _build([6, 3])
    x = Tensor(np.random.rand(5, 4).astype(np.float32))
    y0 = model.forward(x).values.copy()
    path = str(tmp_path / "m.pkl")
    model.save(path)
    np.random.seed(5)
    net2, model2, _ = _build([6, 3])
    model2.forward(x)
    model2.load(path)
    assert np.array_equal(model2.forward(x).values, y0)
