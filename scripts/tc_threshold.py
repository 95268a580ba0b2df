"""Where the tensor-core path (statistics + split + tcgen05 product + the two conditional fallback
launches) overtakes the SIMT GEMM for a product whose operands are fresh every call, as in training.
Prints one line per shape: SIMT ms, tensor-core ms (default split), mixed-split ms."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import core._backend as be  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = be.Event(), be.Event()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    return e1.elapsed_ms_since(e0) / reps


def main():
    be.init()
    rng = np.random.RandomState(0)
    for (M, N, K) in [(256, 256, 256), (512, 256, 256), (512, 512, 256), (1024, 256, 256), (512, 512, 512),
                      (1024, 512, 512), (1024, 1024, 512), (2048, 1024, 512), (2048, 1024, 1024),
                      (4096, 1024, 1024), (128, 784, 200), (8192, 512, 512)]:
        a = be.from_numpy(rng.standard_normal((M, K)).astype(np.float32))
        b = be.from_numpy(rng.standard_normal((K, N)).astype(np.float32))
        out = be.empty((M, N), be.F32)

        def run():
            be.new_split_epoch()          # operands are new every step
            be.matmul(a, b, out=out)
        res = {}
        old_min = be.TC_MIN_MNK
        for mode in ("simt", "f16", "mix"):
            be.TC_ENABLED = mode != "simt"
            be.TC_MIN_MNK = 0
            if mode != "simt":
                be.TC_SPLIT = mode
            res[mode] = timed(run)
        be.TC_ENABLED, be.TC_MIN_MNK, be.TC_SPLIT = True, old_min, "f16"
        print("M%5d N%5d K%5d  log2(MNK)=%.1f  simt %.4f ms  f16 %.4f ms  mix %.4f ms" % (
            M, N, K, np.log2(float(M) * N * K), res["simt"], res["f16"], res["mix"]))


if __name__ == "__main__":
    main()
