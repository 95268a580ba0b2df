"""Quick check of the scaled fp16+bf16 split GEMM (gemm_f16.cu) against float64 numpy products:
orientations, ragged shapes, fused relu / mask outputs, on-device fallback.  Prints max errors."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import core._backend as be  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    be.init()
    be.TC_MIN_MNK = 0
    be.TC_SPLIT = "f16"
    worst = 0.0
    for (M, N, K) in [(128, 256, 64), (256, 256, 128), (256, 512, 96), (384, 768, 200), (100, 300, 52),
                      (129, 257, 36), (1000, 520, 260), (2048, 1024, 512), (4096, 4096, 4096)]:
        rng = np.random.RandomState(M + N + K)
        a = (rng.standard_normal((M, K)) * 3.7e-3).astype(np.float32)
        b = (rng.standard_normal((K, N)) * 41.0).astype(np.float32)
        bias = rng.standard_normal((1, N)).astype(np.float32)
        da, db, dbias = be.from_numpy(a), be.from_numpy(b), be.from_numpy(bias)
        A, B = a.astype(np.float64), b.astype(np.float64)
        ref = A @ B + bias
        e1 = rel(be.matmul(da, db, bias=dbias).numpy(), ref)
        bt = np.ascontiguousarray(b.T)
        e2 = rel(be.matmul(da, be.from_numpy(bt), tb=True).numpy(), A @ B)
        at = np.ascontiguousarray(a.T)
        c0 = rng.standard_normal((M, N)).astype(np.float32)
        dc = be.from_numpy(c0)
        be.matmul(be.from_numpy(at), db, ta=True, out=dc, accumulate=True)
        e3 = rel(dc.numpy(), A @ B + c0)
        e4 = rel(be.matmul(da, db, bias=dbias, relu=True).numpy(), np.maximum(ref, 0))
        # fused relu output -> next product consumes the lazy activation via epilogue statistics
        z, act = be.matmul(da, db, bias=dbias, act=True)
        w2 = (rng.standard_normal((N, 128)) * 0.1).astype(np.float32)
        e5 = rel(be.matmul(act, be.from_numpy(w2)).numpy(),
                 np.maximum(z.numpy().astype(np.float64), 0) @ w2.astype(np.float64))
        # mask epilogue
        pre = rng.standard_normal((M, N)).astype(np.float32)
        dx, masked = be.matmul(da, db, act=True, mask_src=be.from_numpy(pre))
        e6 = rel(masked.numpy(), (A @ B) * (pre >= 0))
        e7 = rel(be.matmul(masked, be.from_numpy(w2)).numpy(),
                 masked.numpy().astype(np.float64) @ w2.astype(np.float64))
        print("M%d N%d K%d: NN %.2e NT %.2e TN+acc %.2e relu %.2e act->next %.2e mask %.2e mask->next %.2e"
              % (M, N, K, e1, e2, e3, e4, e5, e6, e7))
        worst = max(worst, e1, e2, e3, e4, e5, e6, e7)
        be.new_split_epoch()
    # fallback: rows scaled by 1e-30, per-row relative error
    rng = np.random.RandomState(1)
    M, N, K = 512, 512, 512
    a = rng.standard_normal((M, K)).astype(np.float32)
    a[::2] *= np.float32(1e-30)
    b = rng.standard_normal((K, N)).astype(np.float32)
    out = be.matmul(be.from_numpy(a), be.from_numpy(b)).numpy().astype(np.float64)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    denom = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    print("fallback rows*1e-30: max err/(|A||B|) = %.2e" % float(np.max(np.abs(out - ref) / denom)))
    # fallback product with a fused relu output feeding the next product (statistics recomputed on demand)
    z, act = be.matmul(be.from_numpy(a), be.from_numpy(b), act=True)
    w2 = rng.standard_normal((N, 256)).astype(np.float32)
    nxt = be.matmul(act, be.from_numpy(w2)).numpy().astype(np.float64)
    ref2 = np.maximum(ref, 0) @ w2.astype(np.float64)
    print("fallback -> relu -> next product: rel err %.2e" % rel(nxt, ref2))
    a[3, 5] = np.inf
    out = be.matmul(be.from_numpy(a), be.from_numpy(b)).numpy()
    print("inf row non-finite:", bool(np.all(~np.isfinite(out[3]))), " other rows finite:",
          bool(np.all(np.isfinite(out[4:]))))
    print("WORST", worst)


if __name__ == "__main__":
    main()
