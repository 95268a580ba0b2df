"""Captured training step (Model.train_step -> CUDA graph replay) against the eager step.

The recorded step must be the same arithmetic as the five lines of the reference's loop
(run.py:78-83) issued kernel by kernel: losses and parameters are compared bit for bit, across a
change of batch shape (the short last batch of an epoch), on the SIMT path (MNIST-sized layers)
and on the tcgen05 path (TMA descriptors, split-K counters and fused epilogues inside a graph)."""
import numpy as np
import pytest

import ref_numpy as R

pytestmark = pytest.mark.gpu


def _model(widths, seed, opt="adam", d_in=None):
    from core.layers import Dense, ReLU
    from core.losses import SoftmaxCrossEntropyLoss
    from core.model import Model
    from core.nn import Net
    import core.optimizer as O
    np.random.seed(seed)
    layers = []
    dims = [d_in] + list(widths)
    for i, w in enumerate(widths):
        # with d_in the weights are drawn here (not lazily at the first forward), so two models
        # built with the same seed are identical
        layers.append(Dense(w, num_in=dims[i] if d_in is not None else None))
        if i + 1 < len(widths):
            layers.append(ReLU())
    net = Net(layers)
    optimizer = {"adam": lambda: O.Adam(lr=1e-3), "sgd": lambda: O.SGD(lr=1e-2),
                 "rmsprop": lambda: O.RMSProp(lr=1e-3), "momentum": lambda: O.Momentum(lr=1e-2)}[opt]()
    return net, Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=optimizer)


def _eager(model, x, y):
    model.zero_grad()
    pred = model.forward(x)
    loss = model.loss.loss(pred, y)
    loss.backward()
    model.step()
    return loss


def _params(net):
    return [p.values.copy() for layer in net.get_parameters() for p in layer.values()]


def _batches(n_rows, d_in, classes, sizes, seed=0):
    from core.tensor import Tensor
    rng = np.random.RandomState(seed)
    out = []
    for b in sizes:
        x = rng.rand(b, d_in).astype(np.float32)
        lab = rng.randint(0, classes, b)
        out.append((Tensor(x), Tensor(np.eye(classes)[lab])))   # float64 one-hot, as run.py:27-28
    return out


@pytest.mark.parametrize("opt", ["adam", "sgd", "rmsprop", "momentum"])
def test_captured_mnist_step_is_bit_identical(opt):
    widths = [200, 100, 70, 30, 10]
    sizes = [128] * 6 + [80, 80, 80] + [128] * 4 + [80]
    batches = _batches(0, 784, 10, sizes)
    net_a, model_a = _model(widths, 3, opt, d_in=784)
    net_b, model_b = _model(widths, 3, opt, d_in=784)
    for x, y in batches:
        la = float(_eager(model_a, x, y).values)
        lb = float(model_b.train_step(x, y).values)
        assert la == lb
    states = [s for s in model_b._captured.values()]
    assert len(states) == 2 and all(hasattr(s, "graph") for s in states)   # both shapes were recorded
    info = states[0].info()
    assert info["kernel_nodes"] >= 10 and info["blocks"] > 0
    for pa, pb in zip(_params(net_a), _params(net_b)):
        assert np.array_equal(pa, pb)


def test_captured_tensor_core_step_is_bit_identical():
    """256-wide layers at batch 1024: every product is above the tensor-core threshold"""
    import core._backend as be
    widths = [256, 256, 256]
    assert be.use_tensor_cores(1024, 256, 256, be.F32)
    batches = _batches(0, 256, 256, [1024] * 6)
    net_a, model_a = _model(widths, 5, d_in=256)
    net_b, model_b = _model(widths, 5, d_in=256)
    for x, y in batches:
        la = float(_eager(model_a, x, y).values)
        lb = float(model_b.train_step(x, y).values)
        assert la == lb
    for pa, pb in zip(_params(net_a), _params(net_b)):
        assert np.array_equal(pa, pb)


def test_replay_does_not_touch_other_tensors():
    """blocks a recorded step uses stay out of the pool: tensors made between replays survive"""
    from core.tensor import Tensor
    widths = [64, 10]
    batches = _batches(0, 32, 10, [16] * 8)
    net, model = _model(widths, 1)
    keep, expect = [], []
    rng = np.random.RandomState(9)
    for i, (x, y) in enumerate(batches):
        model.train_step(x, y)
        for shape in ((16, 64), (16, 10), (32, 64), (1, 64), (16,)):
            a = rng.rand(*shape).astype(np.float32)
            t = Tensor(a) + 0.0      # result block comes from the pool
            keep.append(t)
            expect.append(a)
    for t, a in zip(keep, expect):
        t._host = None
        assert np.array_equal(t.values, a)


def test_captured_step_loss_copies_are_independent():
    batches = _batches(0, 32, 10, [16] * 6)
    net, model = _model([64, 10], 2)
    held = [model.train_step(x, y) for x, y in batches]
    vals = [float(t.values) for t in held]
    assert len(set(vals)) == len(vals)      # not all aliases of the graph's loss buffer


def test_captured_step_tracks_oracle_trajectory():
    """the replayed MNIST-MLP loop against the oracle's 20-step trajectory: same initial weights
    (np.random.seed(0) draw order), float32 engine vs the oracle's float64 path, rel 1e-4"""
    from core.tensor import Tensor
    x, y, onehot = R.synthetic_mnist(2560, seed=0)
    widths = [200, 100, 70, 30, 10]
    np.random.seed(0)
    mlp = R.RefMLP(widths, R.RefAdam(lr=1e-3))
    want = [float(mlp.train_step(x[i * 128:(i + 1) * 128], onehot[i * 128:(i + 1) * 128]))
            for i in range(20)]
    net, model = _model(widths, 0)
    got = []
    for i in range(20):
        xb, yb = Tensor(x[i * 128:(i + 1) * 128]), Tensor(onehot[i * 128:(i + 1) * 128])
        got.append(float(model.train_step(xb, yb).values))
    assert hasattr(list(model._captured.values())[0], "graph")
    assert np.max(np.abs(np.array(got) - np.array(want)) / np.abs(want)) < 1e-4


def test_host_transfer_inside_capture_fails_loudly():
    import core._backend as be
    g = be.StepGraph()
    with pytest.raises(be.BackendError):
        with g.capture():
            be.from_numpy(np.ones(4, dtype=np.float32))
    # the aborted capture left the stream usable
    a = be.from_numpy(np.arange(4, dtype=np.float32))
    assert np.array_equal(be.to_numpy(be.ew(be.ADD, a, a)), 2 * np.arange(4, dtype=np.float32))


def test_graph_destroyed_during_another_capture_is_deferred():
    """a recorded step released while a second capture is running (the garbage collector can do
    that) must not synchronise the stream: its destruction waits for the capture to end"""
    import core._backend as be
    a = be.from_numpy(np.arange(8, dtype=np.float32))
    g1 = be.StepGraph()
    with g1.capture():
        b = be.ew(be.ADD, a, a)
    g1.replay()
    assert np.array_equal(be.to_numpy(b), 2 * np.arange(8, dtype=np.float32))
    g2 = be.StepGraph()
    with g2.capture():
        c = be.ew(be.MUL, a, a)
        g1.destroy()                       # in the middle of g2's capture
        d = be.ew(be.ADD, c, a)
    g2.replay()
    x = np.arange(8, dtype=np.float32)
    assert np.array_equal(be.to_numpy(d), x * x + x)
    g2.destroy()


def test_recorded_step_is_dropped_when_a_parameter_is_rebound():
    """p.values = ... moves a parameter out of the arena a recorded step names: the next
    train_step must notice, fall back to the eager lines and record again -- same numbers as a
    model that never used a graph"""
    batches = _batches(0, 32, 10, [16] * 10)
    net_a, model_a = _model([64, 10], 7, d_in=32)
    net_b, model_b = _model([64, 10], 7, d_in=32)
    for i, (x, y) in enumerate(batches):
        if i == 5:
            for net in (net_a, net_b):
                w = net.layers[0].params["w"]
                w.values = w.values * 0.5
                w.zero_grad()
        la = float(_eager(model_a, x, y).values)
        lb = float(model_b.train_step(x, y).values)
        assert la == lb, i
    for pa, pb in zip(_params(net_a), _params(net_b)):
        assert np.array_equal(pa, pb)
    assert any(hasattr(s, "graph") for s in model_b._captured.values())
