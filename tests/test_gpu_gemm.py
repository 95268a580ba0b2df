"""tcgen05 3xTF32 GEMM and its tf32 split pre-pass against numpy float64 products, through the
C ABI.  Tolerance: rel 1e-5 of max|ref| (north_star per-op tolerance; the 3xTF32 error model is
~2^-21 per product term)."""
import numpy as np
import pytest

import op_cases

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def be():
    import core._backend as be
    be.init()
    return be


def _ref(a, b, ta, tb, bias):
    A = a.astype(np.float64).T if ta else a.astype(np.float64)
    B = b.astype(np.float64).T if tb else b.astype(np.float64)
    out = A @ B
    if bias is not None:
        out = out + bias.astype(np.float64)
    return out


def test_split_planes_reconstruct(be):
    rng = np.random.RandomState(0)
    for R_, C_ in ((64, 64), (130, 70), (257, 33), (5, 1000)):
        x = (rng.standard_normal((R_, C_)) * 10 ** rng.uniform(-3, 3, (R_, C_))).astype(np.float32)
        d = be.from_numpy(x)
        hi, lo, ldp = be.split_planes(d, transposed=False, also_other=True)
        hit, lot, ldt = be.split_planes(d, transposed=True)
        H, L = hi.numpy()[:, :C_], lo.numpy()[:, :C_]
        HT, LT = hit.numpy()[:, :R_], lot.numpy()[:, :R_]
        # hi/lo are tf32 values (13 low mantissa bits clear) and hi + lo ~= x to ~2^-21
        assert np.all(H.view(np.uint32) & 0x1FFF == 0) and np.all(L.view(np.uint32) & 0x1FFF == 0)
        assert np.max(np.abs((H.astype(np.float64) + L) - x) / np.maximum(np.abs(x), 1e-30)) < 2.0 ** -20
        assert np.array_equal(HT, H.T) and np.array_equal(LT, L.T)


@pytest.mark.parametrize("mn_major", [True, False])
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("shape", [
    (128, 256, 32), (128, 256, 64), (256, 256, 128), (256, 512, 96), (384, 768, 200),
    (100, 300, 52), (129, 257, 36), (1000, 520, 260), (2048, 1024, 512),
])
def test_tf32x3_gemm_shapes(be, cg, shape, mn_major):
    """NN / NT / TN products; mn_major=True feeds transposed operands MN-major from the plain tf32
    planes (default), False goes through transposed planes (both operands K-major)"""
    M, N, K = shape
    be.set_gemm_cta_group(cg)
    old_mn = be.TC_MN_MAJOR
    be.TC_MN_MAJOR = mn_major
    try:
        rng = np.random.RandomState(M * 7 + N * 3 + K)
        a = rng.standard_normal((M, K)).astype(np.float32)
        b = rng.standard_normal((K, N)).astype(np.float32)
        bias = rng.standard_normal((1, N)).astype(np.float32)
        da, db, dbias = be.from_numpy(a), be.from_numpy(b), be.from_numpy(bias)
        old = be.TC_MIN_MNK
        be.TC_MIN_MNK = 0
        try:
            out = be.matmul(da, db, bias=dbias).numpy()
            ref = _ref(a, b, False, False, bias)
            assert op_cases.rel_err(out, ref) <= TOL
            # NT: (M,K) @ (N,K)^T
            bt = np.ascontiguousarray(b.T)
            out = be.matmul(da, be.from_numpy(bt), tb=True).numpy()
            assert op_cases.rel_err(out, _ref(a, bt, False, True, None)) <= TOL
            # TN: (K,M)^T @ (K,N), accumulate into an existing buffer
            at = np.ascontiguousarray(a.T)
            c0 = rng.standard_normal((M, N)).astype(np.float32)
            dc = be.from_numpy(c0)
            be.matmul(be.from_numpy(at), db, ta=True, out=dc, accumulate=True)
            assert op_cases.rel_err(dc.numpy(), _ref(at, b, True, False, None) + c0) <= TOL
            # relu epilogue
            out = be.matmul(da, db, bias=dbias, relu=True).numpy()
            assert op_cases.rel_err(out, np.maximum(ref, 0)) <= TOL
        finally:
            be.TC_MIN_MNK = old
    finally:
        be.TC_MN_MAJOR = old_mn
        be.set_gemm_cta_group(0)


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("ks", [2, 4])
def test_tf32x3_ordered_split_k(be, cg, ks):
    """forced split-K: same result as the unsplit kernel to rounding, bit-identical run to run"""
    rng = np.random.RandomState(11 * ks + cg)
    M, N, K = 700, 520, 2304
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal((1, N)).astype(np.float32)
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    da, db, dbias = be.from_numpy(a), be.from_numpy(b), be.from_numpy(bias)
    ref = _ref(a, b, False, False, bias) + c0
    old = be.TC_MIN_MNK
    be.TC_MIN_MNK = 0
    be.set_gemm_cta_group(cg)
    be.set_gemm_ksplit(ks)
    try:
        outs = []
        for _ in range(3):
            dc = be.from_numpy(c0)
            be.matmul(da, db, bias=dbias, out=dc, accumulate=True)
            outs.append(dc.numpy())
        assert op_cases.rel_err(outs[0], ref) <= TOL
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    finally:
        be.set_gemm_ksplit(0)
        be.set_gemm_cta_group(0)
        be.TC_MIN_MNK = old


def test_tf32x3_better_than_plain_tf32(be):
    """the split really buys fp32-level accuracy: error well below single-pass TF32 (~5e-4)"""
    rng = np.random.RandomState(1)
    M, N, K = 512, 512, 4096
    a = rng.rand(M, K).astype(np.float32)
    b = rng.rand(K, N).astype(np.float32)
    old = be.TC_MIN_MNK
    be.TC_MIN_MNK = 0
    try:
        out = be.matmul(be.from_numpy(a), be.from_numpy(b)).numpy()
    finally:
        be.TC_MIN_MNK = old
    ref = a.astype(np.float64) @ b.astype(np.float64)
    assert op_cases.rel_err(out, ref) <= 2e-6
    # and it is at least as good as numpy's own float32 product on the same data
    assert op_cases.rel_err(out, ref) <= 4 * op_cases.rel_err(a @ b, ref) + 1e-7


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 7), (128, 200, 784), (80, 10, 30), (257, 65, 33), (64, 64, 0)])
def test_simt_gemm(be, dtype, shape):
    M, N, K = shape
    rng = np.random.RandomState(M + N + K)
    a = rng.standard_normal((M, K)).astype(dtype)
    b = rng.standard_normal((K, N)).astype(dtype)
    bias = rng.standard_normal((1, N)).astype(dtype)
    tol = 2e-6 if dtype == np.float32 else 1e-13
    old = be.TC_ENABLED
    be.TC_ENABLED = False
    try:
        da, db = be.from_numpy(a), be.from_numpy(b)
        assert op_cases.rel_err(be.matmul(da, db, bias=be.from_numpy(bias)).numpy(), _ref(a, b, False, False, bias)) <= tol
        at, bt = np.ascontiguousarray(a.T), np.ascontiguousarray(b.T)
        assert op_cases.rel_err(be.matmul(be.from_numpy(at), db, ta=True).numpy(), _ref(at, b, True, False, None)) <= tol
        assert op_cases.rel_err(be.matmul(da, be.from_numpy(bt), tb=True).numpy(), _ref(a, bt, False, True, None)) <= tol
    finally:
        be.TC_ENABLED = old
