"""Per-op parity of the CUDA engine (through the ctypes C ABI) against the numpy oracle and the
committed reference goldens: forward value and every input gradient, rel error
max|a-b| / max|ref| <= 1e-5 in float32 (north_star tolerance) and <= 1e-12 in float64."""
import os

import numpy as np
import pytest

import op_cases
import ref_numpy as R

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5
F64_TOL = 1e-12


@pytest.fixture(scope="module")
def engine():
    import core.ops as ops
    from core.losses import SoftmaxCrossEntropyLoss
    from core.tensor import Tensor
    return Tensor, ops, SoftmaxCrossEntropyLoss()


@pytest.fixture(scope="module")
def gold_ops(golden_dir):
    return np.load(os.path.join(golden_dir, "ops.npz"))


@pytest.mark.parametrize("case", op_cases.CASES, ids=lambda c: c["name"])
def test_op_parity(case, engine, gold_ops):
    Tensor, ops, ce = engine
    tol = F32_TOL if case["dtype"] == "float32" else F64_TOL
    out, grads = op_cases.run_reference_style(case, Tensor, ops, ce.loss)
    o_out, o_grads = op_cases.run_oracle(case, R)
    assert out.shape == o_out.shape
    assert op_cases.rel_err(out, o_out) <= tol
    assert op_cases.rel_err(out, gold_ops[case["name"] + "/out"]) <= tol
    for i, (g, og) in enumerate(zip(grads, o_grads)):
        assert g.shape == og.shape
        assert op_cases.rel_err(g, og) <= tol, "grad %d vs oracle" % i
        assert op_cases.rel_err(g, gold_ops[case["name"] + "/g%d" % i]) <= tol, "grad %d vs golden" % i


def test_ce_fused_equals_composed(engine):
    """the fused loss node and the same expression built from primitive ops agree"""
    Tensor, ops, ce = engine
    rng = np.random.RandomState(3)
    for dt in (np.float32, np.float64):
        z = rng.standard_normal((64, 20)).astype(dt)
        y = np.eye(20)[rng.randint(0, 20, 64)].astype(dt)
        a, b = Tensor(z, requires_grad=True), Tensor(z, requires_grad=True)
        la = ce.loss(a, Tensor(y))
        lb = ce.loss_composed(b, Tensor(y))
        la.backward()
        lb.backward()
        tol = 2e-5 if dt == np.float32 else 1e-12
        assert op_cases.rel_err(la.values, lb.values) <= tol
        assert op_cases.rel_err(a.grad, b.grad) <= tol


@pytest.mark.parametrize("shape", [(128, 10), (80, 10), (1, 1), (2048, 8), (37, 101)])
@pytest.mark.parametrize("zdt,ydt", [(np.float32, np.float64), (np.float32, np.float32), (np.float64, np.float64)])
def test_ce_single_launch_forward_equals_staged(shape, zdt, ydt):
    """small logits: the one-CTA forward (stats + per-row q + loss in one launch) is bit-identical
    to the staged kernels the large / data-parallel path uses"""
    import core._backend as be
    B, C = shape
    rng = np.random.RandomState(B + C)
    z = be.from_numpy((rng.standard_normal((B, C)) * 3).astype(zdt))
    y = be.from_numpy(np.eye(C)[rng.randint(0, C, B)].astype(ydt))
    assert be.ce_small_ok(B, C)
    stats1, loss1, q1, dz1 = be.ce_fwd_small(z, y, B, want_dz=True)
    stats2 = be.ce_stats(z)
    loss2, q2 = be.ce_loss(z, y, stats2, B)
    assert np.array_equal(stats1.numpy(), stats2.numpy())
    assert np.array_equal(q1.numpy(), q2.numpy())
    assert np.array_equal(loss1.numpy(), loss2.numpy())
    # ... and the dL/dz it leaves for backward()'s default seed is what the backward kernel writes
    dz2 = be.ce_bwd(z, y, stats2, q2, B, be.ones_scalar(z.dtype))
    assert np.array_equal(dz1.numpy(), dz2.numpy())
    assert be.ce_fwd_small(z, y, B)[3] is None


def test_ce_soft_labels(engine):
    """general (non one-hot) labels: dL/dz = p - y p / (m q)"""
    Tensor, ops, ce = engine
    rng = np.random.RandomState(4)
    z = rng.standard_normal((17, 9))
    y = rng.rand(17, 9)
    t = Tensor(z, requires_grad=True)
    loss = ce.loss(t, Tensor(y))
    loss.backward()
    rt = R.RefTensor(z, requires_grad=True)
    rl = R.softmax_cross_entropy(rt, y)
    rl.backward()
    assert op_cases.rel_err(loss.values, rl.values) <= 1e-12
    assert op_cases.rel_err(t.grad, rt.grad) <= 1e-10


@pytest.mark.parametrize("name,make_engine,make_ref", [
    ("sgd", lambda o: o.SGD(lr=0.05), lambda: R.RefSGD(0.05)),
    ("adam", lambda o: o.Adam(lr=1e-3), lambda: R.RefAdam(1e-3)),
    ("rmsprop", lambda o: o.RMSProp(lr=0.01, momentum=0.5), lambda: R.RefRMSProp(0.01, momentum=0.5)),
    ("momentum", lambda o: o.Momentum(lr=0.02, momentum=0.9), lambda: R.RefMomentum(0.02, 0.9)),
    ("adagrad", lambda o: o.Adagrad(lr=0.1), lambda: R.RefAdagrad(0.1)),
    ("adadelta", lambda o: o.Adadelta(lr=1.0), lambda: R.RefAdadelta(1.0)),
])
def test_optimizer_steps(name, make_engine, make_ref, golden_dir):
    import core._backend as be
    import core.optimizer as O
    gold = np.load(os.path.join(golden_dir, "optimizers.npz"))
    for dt, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
        opt, ref = make_engine(O), make_ref()
        for k in range(3):
            g = gold["grads"][k].astype(dt)
            step = opt._compute_step(be.from_numpy(g)).numpy()
            rstep = ref._step(gold["grads"][k].copy())
            assert op_cases.rel_err(step, rstep) <= tol
            assert op_cases.rel_err(step, gold[name][k]) <= tol


def test_compute_step_interface():
    """BaseOptimizer.compute_step(grads, params) keeps the reference's list-of-dicts contract"""
    import core.optimizer as O
    from core.tensor import Tensor
    rng = np.random.RandomState(5)
    params = [{"w": Tensor(rng.rand(3, 4), requires_grad=True), "b": Tensor(rng.rand(1, 4), requires_grad=True)},
              {}, {"w": Tensor(rng.rand(4, 2), requires_grad=True), "b": Tensor(rng.rand(1, 2), requires_grad=True)}]
    grads = [{k: rng.standard_normal(v.shape) for k, v in p.items()} for p in params]
    steps = O.SGD(lr=0.5).compute_step(grads, params)
    assert [sorted(s) for s in steps] == [["b", "w"], [], ["b", "w"]]
    for s, g in zip(steps, grads):
        for k in s:
            assert np.allclose(np.asarray(s[k]), -0.5 * g[k])


@pytest.mark.parametrize("shape", [(300, 1024), (64, 4096), (257, 516)])
def test_ce_wide_rows_vectorised_path(shape):
    """float32 logits and labels with wide rows take the 128-bit row kernel and the unrolled
    sum-exp pass: loss and gradient against the oracle"""
    from core.losses import SoftmaxCrossEntropyLoss
    from core.tensor import Tensor
    B, C = shape
    rng = np.random.RandomState(B + C)
    z = (rng.standard_normal((B, C)) * 2).astype(np.float32)
    y = np.eye(C, dtype=np.float32)[rng.randint(0, C, B)]
    y[0] = rng.rand(C).astype(np.float32)            # one soft-label row
    t = Tensor(z, requires_grad=True)
    loss = SoftmaxCrossEntropyLoss().loss(t, Tensor(y))
    loss.backward()
    rt = R.RefTensor(z.astype(np.float64), requires_grad=True)
    rl = R.softmax_cross_entropy(rt, y.astype(np.float64))
    rl.backward()
    assert op_cases.rel_err(loss.values, rl.values) <= 1e-5
    assert op_cases.rel_err(t.grad, rt.grad) <= 1e-5
