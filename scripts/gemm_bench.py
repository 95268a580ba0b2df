"""Micro-benchmark of the tcgen05 3xTF32 GEMM on the three product shapes of the wide MLP
(forward X@W, dX = G@W.T, dW = X.T@G), over the kernel's tunables.

    python scripts/gemm_bench.py [--batch 8192] [--width 4096]
Prints one JSON line per (shape, cg, ksplit, group_m): best-of-10 CUDA-event time, TFLOP/s."""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import core._backend as be  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--cg", default="1,2")
    ap.add_argument("--ksplit", default="0,1")
    ap.add_argument("--group-m", default="1,-8,-4,1,-8,-4")
    ap.add_argument("--sustain", type=int, default=0,
                    help="also time this many back-to-back launches (power-capped steady state)")
    args = ap.parse_args()
    be.init()
    B, D = args.batch, args.width
    rng = np.random.RandomState(0)
    x = be.from_numpy(rng.rand(B, D).astype(np.float32))
    w = be.from_numpy(((rng.rand(D, D) - 0.5) * 0.05).astype(np.float32))
    g = be.from_numpy(rng.standard_normal((B, D)).astype(np.float32))
    be.new_split_epoch()
    shapes = {
        "fwd  X@W   (M=%d,N=%d,K=%d)" % (B, D, D): lambda out: be.matmul(x, w, out=out),
        "dX   G@W.T (M=%d,N=%d,K=%d)" % (B, D, D): lambda out: be.matmul(g, w, tb=True, out=out),
        "dW   X.T@G (M=%d,N=%d,K=%d)" % (D, D, B): lambda out: be.matmul(x, g, ta=True, out=out),
    }
    outs = {k: be.empty((B, D) if not k.startswith("dW") else (D, D), be.F32) for k in shapes}
    flops = 2.0 * B * D * D
    for name, fn in shapes.items():
        fn(outs[name])      # builds the tf32 planes once (cached for the epoch)
    be.sync()
    for cg, ks, gm in itertools.product([int(v) for v in args.cg.split(",")],
                                        [int(v) for v in args.ksplit.split(",")],
                                        [int(v) for v in args.group_m.split(",")]):
        be.set_gemm_cta_group(cg)
        be.set_gemm_ksplit(ks)
        be._lib.tnn_set_gemm_group_m(gm)
        for name, fn in shapes.items():
            for _ in range(3):
                fn(outs[name])
            best = 1e30
            e0, e1 = be.Event(), be.Event()
            for _ in range(10):
                e0.record()
                fn(outs[name])
                e1.record()
                best = min(best, e1.elapsed_ms_since(e0))
            rec = dict(shape=name, cg=cg, ksplit=ks, group_m=gm, ms=round(best, 4),
                       tflops=round(flops / best / 1e9, 1))
            if args.sustain:
                e0.record()
                for _ in range(args.sustain):
                    fn(outs[name])
                e1.record()
                sus = e1.elapsed_ms_since(e0) / args.sustain
                rec.update(sustained_ms=round(sus, 4), sustained_tflops=round(flops / sus / 1e9, 1))
            print(json.dumps(rec))


if __name__ == "__main__":
    main()
