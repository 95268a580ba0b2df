"""Edge cases of the op library against numpy: empty and 0-d tensors, ragged sizes that defeat the
128-bit fast paths, unaligned views, deep broadcasting, mixed dtypes, every flavour of numpy
indexing, large reductions."""
import numpy as np
import pytest

import op_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import core.ops as ops
    from core.tensor import Tensor
    return Tensor, ops


def close(a, b, tol):
    return op_cases.rel_err(np.asarray(a), np.asarray(b)) <= tol


def test_empty_and_zero_dim(api):
    Tensor, ops = api
    e = Tensor(np.zeros((0, 5), np.float32), requires_grad=True)
    assert (e + e).values.shape == (0, 5)
    assert (e * 2).sum().values == 0
    assert e.sum(0).values.tolist() == [0] * 5
    assert ops.exp(e).values.shape == (0, 5)
    assert e.T.shape == (5, 0)
    assert (Tensor(np.zeros((0, 3))) @ Tensor(np.zeros((3, 4)))).values.shape == (0, 4)
    assert (Tensor(np.ones((2, 0))) @ Tensor(np.ones((0, 3)))).values.tolist() == [[0, 0, 0], [0, 0, 0]]
    s = Tensor(3.0, requires_grad=True)
    y = s * s + 2
    assert y.values == 11 and y.shape == ()
    y.backward()
    assert s.grad == 6
    with pytest.raises(TypeError):
        len(s)
    with pytest.raises(ValueError):
        Tensor(np.zeros((0,))).max()
    assert Tensor(np.zeros((3, 4)))[1:1].shape == (0, 4)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 1023, 4097, (1 << 20) + 3])
def test_ragged_flat_sizes(api, dtype, n):
    """sizes that are not multiples of the vector width, and an unaligned view of them"""
    Tensor, ops = api
    rng = np.random.RandomState(n)
    a, b = rng.standard_normal(n).astype(dtype), rng.standard_normal(n).astype(dtype) + 3
    tol = 1e-6 if dtype == np.float32 else 1e-14
    ta, tb = Tensor(a), Tensor(b)
    assert np.array_equal((ta + tb).values, a + b)
    assert np.array_equal((ta * tb).values, a * b)
    assert np.array_equal((ta / tb).values, a / b)
    assert np.array_equal(ops.clip(ta, 0.0).values, a.clip(0.0))
    assert close(ops.exp(ta).values, np.exp(a), tol * 4)
    assert close(ta.sum().values, a.astype(np.float64).sum(), 1e-5 if dtype == np.float32 else 1e-12)
    assert ta.max().values == a.max() and ta.min().values == a.min()
    if n > 2:  # views starting at element 1: 4- or 8-byte aligned only
        va, vb = ta[1:], tb[1:]
        assert np.array_equal((va - vb).values, a[1:] - b[1:])
        assert np.array_equal(ops.clip(va, -0.5, 0.5).values, a[1:].clip(-0.5, 0.5))
        assert close(va.sum().values, a[1:].astype(np.float64).sum(), 1e-5 if dtype == np.float32 else 1e-12)


def test_deep_broadcasting_and_grads(api):
    Tensor, ops = api
    rng = np.random.RandomState(0)
    shapes = [((2, 1, 3, 1, 5), (4, 1, 2, 1)), ((1,), (3, 1, 1)), ((5, 1, 7), (1, 6, 1)), ((2, 3, 4, 5, 6), (6,)),
              ((1, 1, 1), (2, 2, 2)), ((7, 1), (1, 9)), ((3, 1, 5, 1, 2, 1, 2, 1), (4, 1, 3, 1, 2))]
    for sa, sb in shapes:
        a, b = rng.standard_normal(sa), rng.standard_normal(sb) + 3
        ta, tb = Tensor(a, requires_grad=True), Tensor(b, requires_grad=True)
        out = ta * tb + ta / tb
        ref = a * b + a / b
        assert out.shape == ref.shape and close(out.values, ref, 1e-13)
        g = rng.standard_normal(ref.shape)
        out.backward(g)
        ga = (g * b + g / b)
        gb = (g * a - g * a / b ** 2)
        assert close(ta.grad, _unb(ga, sa), 1e-12)
        assert close(tb.grad, _unb(gb, sb), 1e-12)


def _unb(g, shape):
    while g.ndim > len(shape):
        g = g.sum(axis=0)
    for i, d in enumerate(shape):
        if d == 1:
            g = g.sum(axis=i, keepdims=True)
    return g


def test_mixed_dtypes_and_weak_scalars(api):
    Tensor, ops = api
    f32 = Tensor(np.array([1.5, 2.5], np.float32), requires_grad=True)
    f64 = Tensor(np.array([0.1, 0.2], np.float64), requires_grad=True)
    assert (f32 * 3.14).values.dtype == np.float32          # Python scalars are weakly typed
    assert (2 - f32).values.tolist() == [0.5, -0.5]
    mixed = f32 * f64
    assert mixed.values.dtype == np.float64
    mixed.backward([1.0, 1.0])
    assert f32.grad.dtype == np.float32 and f64.grad.dtype == np.float64   # grad has the tensor's dtype
    assert np.allclose(f32.grad, [0.1, 0.2]) and np.allclose(f64.grad, [1.5, 2.5])
    t = Tensor([1, 2, 3])
    assert t.values.dtype == np.float64                      # integers are stored as float64
    t += np.array([0.5, 0.5, 0.5], np.float32)
    assert t.values.tolist() == [1.5, 2.5, 3.5]
    t **= 2
    assert t.values.tolist() == [2.25, 6.25, 12.25]
    m = Tensor(np.eye(2))
    m @= Tensor([[2.0, 0.0], [0.0, 3.0]])
    assert m.values.tolist() == [[2, 0], [0, 3]]


KEYS = [
    slice(None), slice(2, None), slice(None, -2), slice(1, 9, 3), slice(None, None, -1), 4, -1,
    (slice(1, 5), slice(None)), (slice(None), 2), (3, 1), (slice(None), slice(1, None, 2)),
    np.array([0, 2, 2, 9, -1]), [1, 1, 1], np.array([], dtype=np.int64),
    (np.array([0, 3]), np.array([1, 2])), (Ellipsis, 0), (None, slice(2, 4)),
    np.arange(10) % 3 == 0,
]


@pytest.mark.parametrize("key", KEYS, ids=[str(i) for i in range(len(KEYS))])
def test_getitem_any_numpy_key(api, key):
    Tensor, ops = api
    rng = np.random.RandomState(1)
    x = rng.standard_normal((10, 4))
    t = Tensor(x, requires_grad=True)
    out = t[key]
    ref = x[key]
    assert out.shape == ref.shape and np.array_equal(out.values, ref)
    g = rng.standard_normal(ref.shape)
    out.backward(g)
    expect = np.zeros_like(x)
    expect[key] = g                      # assignment semantics, last write wins (ops.py:285-288)
    assert np.array_equal(t.grad, expect)


def test_transpose_reshape_pad_shapes(api):
    Tensor, ops = api
    rng = np.random.RandomState(2)
    for shape, axes in [((3, 5), None), ((33, 65), None), ((100, 257), (1, 0)), ((2, 3, 4), (1, 2, 0)),
                        ((5, 1, 7, 2), (3, 0, 2, 1)), ((4,), None), ((2, 3, 4, 5, 6), (4, 2, 0, 3, 1))]:
        x = rng.standard_normal(shape).astype(np.float32)
        t = Tensor(x, requires_grad=True)
        o = t.transpose(axes)
        assert np.array_equal(o.values, x.transpose(axes))
        g = rng.standard_normal(o.shape).astype(np.float32)
        o.backward(g)
        inv = np.argsort(axes if axes is not None else list(reversed(range(x.ndim))))
        assert np.array_equal(t.grad, g.transpose(inv))
    x = rng.standard_normal((4, 6))
    t = Tensor(x, requires_grad=True)
    assert np.array_equal(t.reshape((-1, 8)).values, x.reshape(-1, 8))
    assert np.array_equal(ops.flatten(t).values, x.ravel())
    with pytest.raises(ValueError):
        t.reshape((5, 5))
    for pw in ([(0, 0), (0, 0)], [(2, 1), (0, 3)], 1, (1, 2)):
        p = ops.pad(t, pw)
        assert np.array_equal(p.values, np.pad(x, pw))
    x3 = rng.standard_normal((2, 3, 2))
    t3 = Tensor(x3, requires_grad=True)
    p = ops.pad(t3, [(1, 1), (0, 2), (3, 0)])
    assert np.array_equal(p.values, np.pad(x3, [(1, 1), (0, 2), (3, 0)]))
    g = rng.standard_normal(p.shape)
    p.backward(g)
    assert np.array_equal(t3.grad, g[1:-1, 0:3, 3:])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_axis_reductions_ragged(api, dtype):
    Tensor, ops = api
    rng = np.random.RandomState(3)
    tol = 2e-6 if dtype == np.float32 else 1e-13
    for shape in [(7, 100003), (100003, 7), (1, 5000), (5000, 1), (3, 1001, 5), (129, 130), (2, 3, 4, 5)]:
        x = rng.standard_normal(shape).astype(dtype)
        t = Tensor(x)
        for ax in list(range(len(shape))) + [None, -1]:
            assert close(t.sum(ax).values, x.astype(np.float64).sum(axis=ax), tol * 10)
            assert np.array_equal(t.max(ax).values, x.max(axis=ax))
            assert np.array_equal(t.min(ax).values, x.min(axis=ax))


def test_nan_and_inf_propagate_like_numpy(api):
    Tensor, ops = api
    x = np.array([np.nan, -1.0, np.inf, -np.inf, 0.0, -0.0], np.float32)
    t = Tensor(x, requires_grad=True)
    with np.errstate(all="ignore"):
        assert np.array_equal(ops.clip(t, 0.0).values, x.clip(0.0), equal_nan=True)
        assert np.array_equal(ops.maximum(t, Tensor(np.zeros(6, np.float32))).values, np.maximum(x, 0), equal_nan=True)
        assert np.allclose(ops.exp(t).values, np.exp(x), rtol=1e-6, equal_nan=True)
        assert np.isnan(t.max().values) and np.isnan(t.sum().values)
    r = ops.clip(t, 0.0)
    r.backward(np.ones(6, np.float32))
    assert t.grad.tolist()[1:] == [0, 1, 0, 1, 1]    # mask is x >= 0 (NaN >= 0 is False)


def test_second_order_graph_reuse_and_accumulation(api):
    """a tensor used many times, gradient accumulation across backward() calls, zero_grad"""
    Tensor, ops = api
    x = Tensor(np.array([1.0, 2.0, 3.0]), requires_grad=True)
    y = x * x * x + x * 2 - x / 2 + ops.exp(x) * 0 + (x ** 2).sum()
    y.backward(np.ones(3))
    # the scalar (x**2).sum() is broadcast into all three outputs: it contributes 3 * 2x
    expect = 3 * np.array([1.0, 4.0, 9.0]) + 2 - 0.5 + 3 * 2 * np.array([1.0, 2.0, 3.0])
    assert np.allclose(x.grad, expect)
    y.backward(np.ones(3))
    assert np.allclose(x.grad, 2 * expect)
    x.zero_grad()
    assert x.grad.tolist() == [0, 0, 0]
