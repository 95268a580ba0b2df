"""CPU-side checks: the C-ABI library loads and exports every symbol include/tnn_b200.h declares,
host-side shape/stride/shard logic, the double-double exp/log (compiled for the host from the
same header the kernels use), and the host utilities that mirror the reference's utils/."""
import ctypes
import os
import re
import subprocess
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
    import core._backend as be
    return be


def test_library_exports_every_declared_symbol(built_lib):
    be = built_lib
    header = open(os.path.join(ROOT, "include", "tnn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(tnn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) > 50
    lib = ctypes.CDLL(be.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes binding types every one of them
    assert declared == set(be.EXPORTED_SYMBOLS), declared ^ set(be.EXPORTED_SYMBOLS)
    be.load_library()


def test_no_cpu_fallback_without_gpu(built_lib):
    """on a box without a GPU the first device operation fails loudly"""
    be = built_lib
    n = ctypes.c_int(0)
    lib = be.load_library()
    has_gpu = lib.tnn_device_count(ctypes.byref(n)) == 0 and n.value > 0
    if has_gpu:
        pytest.skip("GPU present")
    from core.tensor import Tensor
    with pytest.raises(be.BackendError):
        Tensor([1.0, 2.0])


def test_product_path_does_not_import_oracle():
    for sub in ("core", "utils", "tinynn-autograd_b200"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "ref_numpy" not in src and "import oracle" not in src, f


def test_shape_and_stride_helpers(built_lib):
    be = built_lib
    import core.ops as ops
    assert ops._resolve_shape(24, (3, -1)) == (3, 8)
    assert ops._resolve_shape(24, 24) == (24,)
    with pytest.raises(ValueError):
        ops._resolve_shape(24, (5, -1))
    with pytest.raises(ValueError):
        ops._resolve_shape(24, (-1, -1))
    assert be._bstrides((5,), (6, 5)) == [0, 1]
    assert be._bstrides((1, 5), (6, 5)) == [0, 1]
    assert be._bstrides((6, 1), (6, 5)) == [1, 0]
    assert be._bstrides((), (6, 5)) == [0, 0]
    assert be._bstrides((3, 1), (2, 3, 4)) == [0, 1, 0]
    assert be._cstrides((2, 3, 4)) == [12, 4, 1]
    assert be.device_dtype(np.int64) == np.float64 and be.device_dtype(np.float32) == np.float32
    assert be._round4(70) == 72


def test_duplicate_index_scatter_resolution():
    """getitem_ backward is an assignment (ops.py:285-288): numpy lets the last occurrence of a
    repeated index win; the host resolves that before the (parallel) scatter kernel runs"""
    import core.ops as ops
    rng = np.random.RandomState(0)
    for n, m in ((5, 9), (40, 7), (1000, 50)):
        idx = rng.randint(-m, m, n)
        g = rng.standard_normal(n)
        ref = np.zeros(m)
        ref[idx] = g
        res = ops._last_writer(idx, m)
        out = np.zeros(m)
        if res is None:
            out[np.where(idx < 0, idx + m, idx)] = g
        else:
            uniq, last = res
            out[uniq] = g[last]
        assert np.array_equal(out, ref)
    assert ops._last_writer(np.array([3, 1, 2]), 5) is None


def test_shard_bounds_and_stats_merge():
    import core._dist as dist
    for n, w in ((65536, 8), (10, 3), (7, 8)):
        cuts = [dist.shard_bounds(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        sizes = [b - a for a, b in cuts]
        assert max(sizes) - min(sizes) <= 1
    rng = np.random.RandomState(0)
    z = rng.standard_normal((64, 10)) * 5
    parts = np.array_split(z, 4)
    pairs = [(p.max(), np.exp(p - p.max()).sum()) for p in parts]
    m, s = dist.merge_stats_host(pairs)
    assert m == z.max() and abs(s - np.exp(z - z.max()).sum()) <= 1e-12 * s


def test_exp_log_double_double_on_host(tmp_path):
    """math.cuh's exp_cr/log_cr (the float64 device path) compiled with g++: correctly rounded on
    the reference's known-answer inputs and on a random sweep (vs mpmath when available)"""
    src = tmp_path / "m.cpp"
    src.write_text('#include "math.cuh"\n'
                   'extern "C" void exp_arr(const double* x, double* y, int n){for(int i=0;i<n;i++)y[i]=tnn::exp_cr(x[i]);}\n'
                   'extern "C" void log_arr(const double* x, double* y, int n){for(int i=0;i<n;i++)y[i]=tnn::log_cr(x[i]);}\n')
    so = tmp_path / "libm_test.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                           "-I", os.path.join(ROOT, "tinynn-autograd_b200", "csrc"), str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))

    def run(fn, x):
        y = np.empty_like(x)
        fn(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), len(x))
        return y
    k = np.array([1.0, 3.0, 5.0])
    assert [v.hex() for v in run(lib.exp_arr, k)] == [
        "0x1.5bf0a8b145769p+1", "0x1.415e5bf6fb106p+4", "0x1.28d389970338fp+7"]
    assert [v.hex() for v in run(lib.log_arr, k)] == [
        "0x0.0p+0", "0x1.193ea7aad030bp+0", "0x1.9c041f7ed8d33p+0"]
    rng = np.random.RandomState(0)
    x = rng.uniform(-40, 40, 2000)
    e = run(lib.exp_arr, x)
    assert np.max(np.abs(e - np.exp(x)) / np.spacing(np.exp(x))) <= 1.0
    try:
        import mpmath
    except ImportError:
        return
    mpmath.mp.prec = 200
    cr = np.array([float(mpmath.exp(mpmath.mpf(float(v)))) for v in x[:500]])
    assert np.array_equal(e[:500], cr)
    y = np.exp(rng.uniform(-60, 60, 500))
    crl = np.array([float(mpmath.log(mpmath.mpf(float(v)))) for v in y])
    assert np.array_equal(run(lib.log_arr, y), crl)


# ---- host utilities mirrored from the reference's utils/ and core/initializer.py ---------------
def test_initializers_statistics():
    import core.initializer as I
    shape, tol = (100000, 1), 1e-2
    assert I.get_fans((100, 10)) == (100, 10)
    fi, fo = I.get_fans((64, 5, 5, 128))
    assert fi == 5 * 5 * 128 and fo == 64
    v = I.NormalInit(0.0, 1.0).init(shape)
    assert abs(v.mean()) <= tol and abs(v.std() - 1.0) <= tol
    v = I.TruncatedNormalInit(0.0, 1.0).init(shape)
    assert abs(v.mean()) <= tol and v.min() >= -2.0 and v.max() <= 2.0
    v = I.UniformInit(-1.0, 1.0).init(shape)
    assert v.min() >= -1.0 and v.max() <= 1.0
    assert np.all(I.ConstantInit(3.1).init(shape) == 3.1)
    b = np.sqrt(6.0 / np.sum(I.get_fans(shape)))
    v = I.XavierUniformInit().init(shape)
    assert v.min() >= -b and v.max() <= b
    s = np.sqrt(2.0 / np.sum(I.get_fans(shape)))
    assert abs(I.XavierNormalInit().init(shape).std() - s) <= tol
    b = np.sqrt(6.0 / I.get_fans(shape)[0])
    v = I.HeUniformInit().init(shape)
    assert v.min() >= -b and v.max() <= b
    s = np.sqrt(2.0 / I.get_fans(shape)[0])
    assert abs(I.HeNormalInit().init(shape).std() - s) <= tol
    # same RNG stream as the reference: Xavier draws are np.random.uniform(-a, a, shape)
    np.random.seed(0)
    a = I.XavierUniformInit().init([784, 200])
    np.random.seed(0)
    bound = np.sqrt(6.0 / 984)
    assert np.array_equal(a, np.random.uniform(low=-bound, high=bound, size=[784, 200]))


def test_batch_iterator_on_arrays():
    from utils.data_iterator import BatchIterator
    x = np.random.randint(0, 100, size=(100, 10))
    y = np.random.randint(0, 100, size=(100, 5))
    n = 0
    for bx, by in BatchIterator(batch_size=10)(x, y):
        assert bx.shape == (10, 10) and by.shape == (10, 5)
        n += 1
    assert n == 10
    # ragged tail (50000 = 390*128 + 80 in run.py) and the shuffle's RNG call
    sizes = [len(b.inputs) for b in BatchIterator(batch_size=128, shuffle=False)(np.zeros((1000, 2)), np.zeros((1000, 1)))]
    assert sizes == [128] * 7 + [104]
    np.random.seed(3)
    first = next(iter(BatchIterator(batch_size=4)(np.arange(20), np.arange(20)))).inputs
    np.random.seed(3)
    idx = np.arange(20)
    np.random.shuffle(idx)
    assert first.tolist() == idx[:4].tolist()


def test_seeder_and_timer():
    from utils.seeder import random_seed
    from utils.timer import Timer
    with pytest.raises(ValueError):
        random_seed(2 ** 32)
    random_seed(1)
    a = np.random.rand()
    random_seed(1)
    assert a == np.random.rand()
    t = Timer("t")
    t.start()
    time.sleep(0.05)
    t.pause()
    time.sleep(0.02)
    t.start()
    time.sleep(0.05)
    t.stop()
    assert t.count == 2 and 0.1 <= t.duration <= 0.2


def test_coverage_document_cites_existing_tests_and_files():
    """COVERAGE.md maps every SURVEY section-8 row to tests and evidence files: each one it names exists"""
    text = open(os.path.join(ROOT, "COVERAGE.md")).read()
    sources = ""
    for f in os.listdir(os.path.join(ROOT, "tests")):
        if f.endswith(".py"):
            sources += open(os.path.join(ROOT, "tests", f)).read()
    for name in sorted(set(re.findall(r"\b(test_[a-z0-9_]+)\b", text))):
        if os.path.exists(os.path.join(ROOT, "tests", name + ".py")) or name == "test_autograd":
            continue     # a test file of this repository / the reference's test/test_autograd.py
        assert "def %s(" % name in sources, name
    for prof in sorted(set(re.findall(r"`((?:r01[a-z]_|gemm_)[A-Za-z0-9_.]+\.(?:jsonl|json|md|log))`", text))):
        assert os.path.exists(os.path.join(ROOT, "profiles", prof)), prof


def test_bench_arms_print_the_same_config_and_a_fixed_cpu_sample():
    """bench.py: the B200 arm and the reference arm describe the workload with the same `config`
    object, key for key (the driver compares them), and the CPU arm's sample size is a constant of the
    workload -- it does not depend on --steps"""
    sys.path.insert(0, ROOT)
    import bench
    for cfg, world in ((bench.WIDE, 1), (bench.WIDE, 8), (bench.MNIST, 1)):
        a = bench.make_config(dict(cfg), cfg["batch"], world)
        b = bench.make_config(dict(cfg), cfg["batch"], world)
        assert a == b and a["global_batch"] == cfg["batch"] * world
        assert a["cpu_reference_sample"] == "%d rows per step" % bench.ref_sample_batch(cfg)
    assert bench.ref_sample_batch(bench.WIDE) == 1024 and bench.ref_sample_batch(bench.MNIST) == 128
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "pick_reference_sample" not in src           # the steps-dependent sample size of r01 is gone
    assert bench.gemm_flops_per_step(bench.WIDE, 8192) == 11 * 2 * 8192 * 4096 * 4096


def test_reference_arm_runs_on_cpu_and_reports_its_sample():
    """`bench.py --impl reference` on the MNIST workload (fast): one JSON line with impl, the shared
    config, a cpu_baseline describing the run, and the zero-copy e2e object the contract asks for"""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "mnist", "--steps", "3", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["value"] > 0
    assert line["steps"] == 3 and line["config"]["workload"].startswith("mnist_mlp")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "128 of its 128 rows" in line["cpu_baseline"]["sample"]


def test_numa_binding_helpers():
    sys.path.insert(0, ROOT)
    import core._dist as dist
    assert dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert dist._parse_cpulist("") == set()
    # no GPU / no sysfs entry here: nothing is changed and nothing raises
    before = os.sched_getaffinity(0)
    assert dist.bind_to_local_numa_node(0) is None or isinstance(dist.bind_to_local_numa_node(0), str)
    assert os.sched_getaffinity(0) == before


def test_reference_format_checkpoint_unpickles_without_a_gpu():
    """the fixture written by the reference's Model.save: unpickling needs no device, and the arrays
    come out of the reference's own attribute names"""
    import pickle
    sys.path.insert(0, ROOT)
    from core.model import _pickled_values
    with open(os.path.join(ROOT, "tests", "golden", "ref_checkpoint.pkl"), "rb") as f:
        net = pickle.load(f)
    shapes = [tuple(_pickled_values(layer.params[k]).shape) for layer in net.layers
              for k in ("w", "b") if getattr(layer, "params", None)]
    assert shapes == [(4, 3), (1, 3), (3, 2), (1, 2)]


def test_postponed_iteration_state_machine_without_a_gpu():
    """core/_deferred.py on stand-in tensors (no device): the five lines are only noted; any other
    use runs them in order (forward, loss, backward) exactly once; the lazy objects BECOME the
    computed tensors (identity kept, class switched); reading a parameter's gradient, or starting
    another iteration, flushes the pending one; a replayed step hands out loss and prediction"""
    import types
    import core._deferred as D
    import core.tensor as T
    from core.tensor import Tensor

    class FakeArray(object):
        def __init__(self, shape, tag):
            self.shape, self.dtype, self.tag, self.size = shape, np.dtype(np.float32), tag, int(np.prod(shape))

    def fake_tensor(shape, tag, requires_grad=True):
        t = Tensor.__new__(Tensor)
        t.__dict__.update(_data=FakeArray(shape, tag), _host=None, _grad=None, _grad_zero=requires_grad,
                          _grad_host=None, _gslot=None, _relu_pre=None, _fused_bwd=None,
                          requires_grad=requires_grad, dependency=[])
        return t

    calls = []

    class Net(object):
        def forward(self, x):
            calls.append("forward")
            return fake_tensor((4, 3), "pred")

    class Loss(object):
        _weight = None

        def loss(self, pred, y):
            assert type(pred) is Tensor and pred._data.tag == "pred"      # already adopted
            calls.append("loss")
            out = fake_tensor((), "loss")
            out.backward = lambda grad=None: calls.append("backward")
            return out

    model = types.SimpleNamespace(net=Net(), _note_output=lambda x, out: None)
    x, y = fake_tensor((4, 5), "x", False), fake_tensor((4, 3), "y", False)
    T._DEFERRED[0] = None

    # --- the five lines are only noted
    pred = D.begin(model, x, (4, 3), np.dtype(np.float32))
    assert type(pred) is D.LazyTensor and pred.shape == (4, 3) and pred.ndim == 2 and pred.requires_grad
    loss = D.defer_loss(Loss(), pred, y)
    assert type(loss) is D.LazyTensor and loss.shape == ()
    loss.backward()
    chain = T._DEFERRED[0]
    assert chain is not None and chain.stage == "backward" and calls == []

    # --- a parameter's gradient is read: the postponed lines run, in order, once
    param = fake_tensor((5, 3), "w")
    g = param.grad
    assert g.shape == (5, 3) and calls == ["forward", "loss", "backward"] and T._DEFERRED[0] is None
    assert type(pred) is Tensor and pred._data.tag == "pred" and type(loss) is Tensor and loss._data.tag == "loss"
    chain.materialise()
    assert calls == ["forward", "loss", "backward"]                  # idempotent

    # --- touching the prediction first: only the forward runs; the loss is then computed eagerly
    del calls[:]
    pred = D.begin(model, x, (4, 3), np.dtype(np.float32))
    assert pred._data.tag == "pred" and calls == ["forward"] and type(pred) is Tensor
    assert D.defer_loss(Loss(), pred, y) is None                     # not a pending prediction any more

    # --- a second iteration started while one is pending flushes the first
    del calls[:]
    p1 = D.begin(model, x, (4, 3), np.dtype(np.float32))
    l1 = D.defer_loss(Loss(), p1, y)
    p2 = D.begin(model, x, (4, 3), np.dtype(np.float32))
    assert calls == ["forward", "loss"] and type(p1) is Tensor and type(l1) is Tensor
    assert type(p2) is D.LazyTensor and T._DEFERRED[0].pred is p2

    # --- things that are not the pattern are refused (and fall through to the eager code)
    assert D.defer_loss(Loss(), p2, fake_tensor((4, 2), "y2", False)) is None     # other label shape
    weighted = Loss()
    weighted._weight = np.ones(3)
    assert D.defer_loss(weighted, p2, y) is None
    l2 = D.defer_loss(Loss(), p2, y)
    del calls[:]
    l2.backward(2.0)                                                 # a seeded backward is not postponed
    assert calls == ["forward", "loss", "backward"] and T._DEFERRED[0] is None

    # --- a replayed step serves its results
    p3 = D.begin(model, x, (4, 3), np.dtype(np.float32))
    l3 = D.defer_loss(Loss(), p3, y)
    l3.backward()
    chain = T._DEFERRED[0]
    T._DEFERRED[0] = None                                            # Model.step() takes the chain over
    step = types.SimpleNamespace(prediction_source=lambda: (FakeArray((4, 3), "logits"), FakeArray((4, 3), "dz")))
    import core._backend as be
    orig = (be.clone, be.ones_scalar)
    be.clone = lambda a: FakeArray(a.shape, a.tag + "-copy")
    be.ones_scalar = lambda dt: FakeArray((), "one")
    try:
        del calls[:]
        D.serve_after_replay(chain, step, FakeArray((), "loss-value"))
        assert type(l3) is Tensor and l3._data.tag == "loss-value" and l3._grad.tag == "one"
        assert type(p3) is D.LazyTensor and step.live_prediction() is p3
        D.pin_live_prediction(step)                                  # the next replay is about to run
        assert type(p3) is Tensor and p3._data.tag == "logits-copy" and p3._grad.tag == "dz-copy"
        assert p3.dependency == [] and calls == [] and step.live_prediction is None
    finally:
        be.clone, be.ones_scalar = orig
