"""Times the SIMT GEMM on the MNIST-MLP product shapes (back-to-back launches, CUDA events).
    TNN_SIMT_MAX_SPLIT=1|2|4|8 python scripts/simt_bench.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import core._backend as be  # noqa: E402

be.init()
rng = np.random.RandomState(0)
for (M, K, N) in ((128, 784, 200), (128, 200, 100), (128, 100, 70), (128, 70, 30), (128, 30, 10)):
    a = be.from_numpy(rng.rand(M, K).astype(np.float32))
    w = be.from_numpy(rng.rand(K, N).astype(np.float32))
    b = be.from_numpy(rng.rand(1, N).astype(np.float32))
    for _ in range(20):
        be.matmul(a, w, bias=b, act=True)
    e0, e1 = be.Event(), be.Event()
    e0.record()
    for _ in range(500):
        be.matmul(a, w, bias=b, act=True)
    e1.record()
    print(json.dumps(dict(shape=(M, K, N), max_split=os.environ.get("TNN_SIMT_MAX_SPLIT", "8"),
                          us=round(e1.elapsed_ms_since(e0) * 2, 2))))
