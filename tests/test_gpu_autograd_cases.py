"""The reference's own autograd suite (test/test_autograd.py), restated against the CUDA engine.
Same tensors, same operators, same exact-equality asserts (the values are the known-answer
vectors listed in SURVEY.md 8c; /root/reference is not readable on the GPU box, and its sources
may not be copied, so the cases are written out again here)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import core.ops as ops
    from core.tensor import Tensor
    return Tensor, ops


def test_add(api):
    Tensor, ops = api
    t1, t2 = Tensor([1, 3, 5], requires_grad=True), Tensor([5, -2, -9], requires_grad=True)
    t3 = t1 + t2
    assert t3.values.tolist() == [6, 1, -4]
    t3.backward([2, 2, 2])
    assert t1.grad.tolist() == [2, 2, 2] and t2.grad.tolist() == [2, 2, 2]
    t1 = Tensor([[1, 3, 5], [2, 3, 0]], requires_grad=True)
    t2 = Tensor([5, -2, -9], requires_grad=True)
    t3 = t1 + t2
    assert t3.values.tolist() == [[6, 1, -4], [7, 1, -9]]
    t3.backward([[1, 1, 1], [2, 2, 2]])
    assert t1.grad.tolist() == [[1, 1, 1], [2, 2, 2]] and t2.grad.tolist() == [3, 3, 3]
    t2 = Tensor([[5, -2, -9]], requires_grad=True)
    t1 = Tensor([[1, 3, 5], [2, 3, 0]], requires_grad=True)
    t3 = t1 + t2
    t3.backward([[1, 1, 1], [2, 2, 2]])
    assert t1.grad.tolist() == [[1, 1, 1], [2, 2, 2]] and t2.grad.tolist() == [[3, 3, 3]]


def test_mul_div_pow(api):
    Tensor, ops = api
    t1, t2 = Tensor([1, 3, 5], requires_grad=True), Tensor([5, -2, -9], requires_grad=True)
    t3 = t1 * t2
    assert t3.values.tolist() == [5, -6, -45]
    t3.backward([2, 2, 2])
    assert t1.grad.tolist() == [10, -4, -18] and t2.grad.tolist() == [2, 6, 10]
    t1, t2 = Tensor([1, 2, 5], requires_grad=True), Tensor([8, -2, -10], requires_grad=True)
    t3 = t1 / t2
    assert t3.values.tolist() == [0.125, -1, -0.5]
    t3.backward([1, 1, 1])
    assert t1.grad.tolist() == [0.125, -0.5, -0.1]
    assert t2.grad.tolist() == [-0.015625, -0.5, -0.05]
    t1 = Tensor([1, -3, 5], requires_grad=True)
    t2 = t1 ** 3
    assert t2.values.tolist() == [1, -27, 125]
    t2.backward([2, 2, 2])
    assert t1.grad.tolist() == [6, 54, 150]


def test_dot_sum_neg(api):
    Tensor, ops = api
    t1 = Tensor([[1, 3, 5], [5, -2, 9]], requires_grad=True)
    t2 = Tensor([[9, 8, 9, 7], [4, 0, 3, 0], [0, 8, 2, 7]], requires_grad=True)
    t3 = t1 @ t2
    assert t3.values.tolist() == [[21, 48, 28, 42], [37, 112, 57, 98]]
    t3.backward([[1, 2, 3, 4], [4, 3, 2, 1]])
    assert t1.grad.tolist() == [[80, 13, 50], [85, 22, 35]]
    assert t2.grad.tolist() == [[21, 17, 13, 9], [-5, 0, 5, 10], [41, 37, 33, 29]]
    t1, t2 = Tensor([1, 3, 5], requires_grad=True), Tensor([5, -2, -9], requires_grad=True)
    t3 = (t1 + t2).sum()
    assert t3.values == 3
    t3.backward(2)
    assert t1.grad.tolist() == [2, 2, 2] and t2.grad.tolist() == [2, 2, 2]
    t1 = Tensor([1, 3, 5], requires_grad=True)
    t2 = -t1
    assert t2.values.tolist() == [-1, -3, -5]
    t2.backward([1, 2, 3])
    assert t1.grad.tolist() == [-1, -2, -3]


def test_exp_log_bit_exact(api):
    """float64 exp/log must equal numpy bit for bit at {1,3,5} (test_autograd.py:90-96, 182-189)"""
    Tensor, ops = api
    t1 = Tensor([1, 3, 5], requires_grad=True)
    t2 = ops.exp(t1)
    assert t2.values.tolist() == np.exp(t1.values).tolist()
    assert [v.hex() for v in t2.values.tolist()] == [
        "0x1.5bf0a8b145769p+1", "0x1.415e5bf6fb106p+4", "0x1.28d389970338fp+7"]
    t2.backward([1, 2, 3])
    assert t1.grad.tolist() == (np.exp(t1.values) * np.array([1, 2, 3])).tolist()
    t1 = Tensor([1, 3, 5], requires_grad=True)
    t2 = ops.log(t1)
    assert t2.values.tolist() == np.log(t1.values).tolist()
    grad = np.array([1, 2, 3])
    t2.backward(grad)
    assert t1.grad.tolist() == (grad / np.array([1, 3, 5])).tolist()


def test_exp_log_correctly_rounded_sweep(api):
    """the double-double exp/log agree with numpy to <= 1 ulp everywhere and exactly almost
    everywhere (numpy itself is within 1 ulp of the correctly rounded value)"""
    Tensor, ops = api
    rng = np.random.RandomState(0)
    x = rng.uniform(-30, 30, 4096)
    e = ops.exp(Tensor(x)).values
    ne = np.exp(x)
    assert np.max(np.abs(e - ne) / np.spacing(ne)) <= 1.0
    y = np.exp(rng.uniform(-50, 50, 4096))
    l = ops.log(Tensor(y)).values
    nl = np.log(y)
    assert np.max(np.abs(l - nl) / np.spacing(np.abs(nl) + 1e-300)) <= 1.0


def test_minimal_nn(api):
    """100 SGD steps on y = 3.14 x + 30, float64, loss strictly decreasing (test_autograd.py:108-126)"""
    Tensor, ops = api
    np.random.seed(0)
    x = Tensor(np.random.normal(0, 1.0, (100, 3)))
    y = x * 3.14 + 30
    w1 = Tensor(np.random.normal(0, 1.0, (3, 3)), requires_grad=True)
    b1 = Tensor(np.random.normal(0, 1.0, 3), requires_grad=True)
    previous_loss = 1e10
    for _ in range(100):
        w1.zero_grad()
        b1.zero_grad()
        predicted = x @ w1 + b1
        err = predicted - y
        loss = (err ** 2).sum()
        loss.backward()
        w1 -= 0.001 * w1.grad
        b1 -= 0.001 * b1.grad
        assert loss.values < previous_loss
        previous_loss = loss.values


def test_maximum_minimum(api):
    Tensor, ops = api
    t1, t2 = Tensor([1, 3, 5], requires_grad=True), Tensor([5, -2, 9], requires_grad=True)
    t3 = ops.maximum_(t1, t2)
    assert t3.values.tolist() == [5, 3, 9]
    t3.backward([1, 2, 1])
    assert t1.grad.tolist() == [0, 2, 0] and t2.grad.tolist() == [1, 0, 1]
    t1, t2 = Tensor([1, 3, 5], requires_grad=True), Tensor([5, -2, 9], requires_grad=True)
    t3 = ops.minimum_(t1, t2)
    assert t3.values.tolist() == [1, -2, 5]
    t3.backward([1, 2, 1])
    assert t1.grad.tolist() == [1, 0, 1] and t2.grad.tolist() == [0, 2, 0]
    # ties go to ts1 (ops.py:170,179)
    t1, t2 = Tensor([2.0, 2.0], requires_grad=True), Tensor([2.0, 3.0], requires_grad=True)
    ops.maximum_(t1, t2).backward([1, 1])
    assert t1.grad.tolist() == [1, 0] and t2.grad.tolist() == [0, 1]


def test_transpose_shapes(api):
    Tensor, ops = api
    shape = [2, 4, 6]
    data = np.random.randn(*shape)
    t1 = Tensor(data, requires_grad=True)
    t2 = t1.T
    assert list(t2.shape) == shape[::-1]
    assert np.array_equal(t2.values, data.T)
    t2.backward(np.ones_like(t2.values))
    assert list(t1.grad.shape) == shape
    t2 = t1.transpose((2, 0, 1))
    assert list(t2.shape) == [6, 2, 4]
    assert np.array_equal(t2.values, data.transpose((2, 0, 1)))
    t2.backward(np.ones_like(t2.values))
    assert list(t1.grad.shape) == shape


def test_max(api):
    Tensor, ops = api
    t1 = Tensor([[1, 3, 5], [3, 7, -2]], requires_grad=True)
    t2 = ops.max(t1, axis=None)
    t3 = ops.max(t1, axis=0)
    assert t2.values == 7
    assert t3.values.tolist() == [3, 7, 5]
    t2.backward()
    assert t1.grad.tolist() == [[0, 0, 0], [0, 1, 0]]
    t1.zero_grad()
    t3.backward([1, 1, 1])
    assert t1.grad.tolist() == [[0, 0, 1], [1, 1, 0]]


def test_reshape_pad_flatten_clip(api):
    Tensor, ops = api
    t1 = Tensor([[1, 2, 3], [4, 5, 6]], requires_grad=True)
    t2 = ops.reshape(t1, (6,))
    assert t2.values.tolist() == [1, 2, 3, 4, 5, 6]
    t2.backward(np.ones(6))
    assert t1.grad.tolist() == [[1, 1, 1], [1, 1, 1]]
    t1 = Tensor([[1, 2, 3], [4, 5, 6]], requires_grad=True)
    t2 = ops.pad(t1, [(1, 0), (1, 0)])
    assert t2.values.tolist() == [[0, 0, 0, 0], [0, 1, 2, 3], [0, 4, 5, 6]]
    t2.backward(np.ones_like(t2.values))
    assert t1.grad.shape == t1.shape and t1.grad.tolist() == [[1, 1, 1], [1, 1, 1]]
    t1 = Tensor([[1, 2, 3], [4, 5, 6]], requires_grad=True)
    t2 = ops.flatten(t1)
    assert t2.values.tolist() == [1, 2, 3, 4, 5, 6]
    t2.backward(np.ones_like(t2.values))
    assert t1.grad.tolist() == [[1, 1, 1], [1, 1, 1]]
    t1 = Tensor([1, -3, 5], requires_grad=True)
    t2 = ops.clip(t1, 0)
    assert t2.values.tolist() == [1, 0, 5]
    t2.backward(np.array([1, 2, 3]))
    assert t1.grad.tolist() == [1, 0, 3]


def test_semantics_kept_from_reference(api):
    """quirks the engine preserves"""
    Tensor, ops = api
    # Q2: ReLU'(0) = 1, mask from the pre-activation
    t = Tensor(np.array([-1.0, 0.0, 2.0], dtype=np.float32), requires_grad=True)
    ops.clip(t, 0.0).backward([1, 1, 1])
    assert t.grad.tolist() == [0, 1, 1]
    # Q5: a second backward accumulates; non-leaf nodes expose .grad too
    a = Tensor([1.0, 2.0], requires_grad=True)
    b = a * 2
    c = b.sum()
    c.backward()
    c.backward()
    assert a.grad.tolist() == [4, 4] and b.grad.tolist() == [2, 2]
    # Q9: getitem backward assigns (duplicates not accumulated)
    x = Tensor([1.0, 2.0, 3.0], requires_grad=True)
    x[np.array([0, 0, 2])].backward([1, 1, 1])
    assert x.grad.tolist() == [1, 0, 1]
    # Q11: the values setter / in-place ops drop the gradient; backward then raises TypeError
    w = Tensor([1.0, 2.0], requires_grad=True)
    w -= 0.5
    assert w.grad is None and w.values.tolist() == [0.5, 1.5]
    with pytest.raises(TypeError):
        (w * 2).sum().backward()
    # diamonds are differentiated correctly by the single sweep (c = b + b; d = c * c)
    bb = Tensor([2.0], requires_grad=True)
    cc = bb + bb
    (cc * cc).backward()
    assert bb.grad.tolist() == [16.0]
    # errors
    with pytest.raises(AssertionError):
        Tensor([1.0]).backward()
    with pytest.raises(ValueError):
        Tensor(np.ones((2, 3))) + Tensor(np.ones((4,)))
    # comparisons give raw bool arrays; numpy consumes a Tensor through __array__
    assert (Tensor([1, 5]) > Tensor([2, 2])).tolist() == [False, True]
    assert np.argmax(Tensor([[1, 9], [7, 3]]), axis=1).tolist() == [1, 0]
    assert len(Tensor(np.zeros((5, 2)))) == 5
