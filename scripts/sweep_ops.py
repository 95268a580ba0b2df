"""BASELINE.json config 3: elementwise / broadcast / reduce sweep, achieved HBM GB/s per op and
size against the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs).

    python scripts/sweep_ops.py [--max-log2 30] [--cpu] > gpurun_out/sweep.json

Algorithmic bytes per element (fp32): add/mul 12, bias-add 8, relu fwd 8, relu bwd 12,
unbroadcast (R,1024)->(1,1024) 4, Adam 28.  Each timing is the best of 10 launches with CUDA
events on the compute stream after 3 warm-ups; the L2 is flushed before every timed launch for
sizes whose working set fits in it.  With --cpu the numpy oracle's ops are timed next to it."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import core._backend as be  # noqa: E402


def timed(fn, flush, reps=10, warm=3):
    for _ in range(warm):
        fn()
    best = 1e30
    e0, e1 = be.Event(), be.Event()
    for _ in range(reps):
        if flush:
            be.l2_flush()
        e0.record()
        fn()
        e1.record()
        best = min(best, e1.elapsed_ms_since(e0))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    be.init()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    l2 = be.device_info()["l2_bytes"]
    rows = []
    for lg in range(20, args.max_log2 + 1, 2):
        n = 1 << lg
        R_, C_ = n // 1024, 1024
        a = be.full((R_, C_), 0.5, be.F32)
        b = be.full((R_, C_), 0.25, be.F32)
        out = be.empty((R_, C_), be.F32)
        bias = be.full((1, C_), 0.125, be.F32)
        cs = be.empty((1, C_), be.F32)
        flush = 3 * n * 4 <= 2 * l2
        m = be.zeros((n,), be.F32)
        v = be.zeros((n,), be.F32)
        h = [1e-3, 0.9, 0.999, 1e-8, 0.1, 0.001]
        ops = [
            ("add", 12, lambda: be.ew(be.ADD, a, b, out=out)),
            ("mul", 12, lambda: be.ew(be.MUL, a, b, out=out)),
            ("bias_add", 8, lambda: be.ew(be.ADD, a, bias, out=out)),
            ("relu_fwd", 8, lambda: be._lib.tnn_relu_fwd(0, out.ptr, a.ptr, n)),
            ("relu_bwd", 12, lambda: be._lib.tnn_relu_bwd(0, out.ptr, b.ptr, a.ptr, n)),
            ("unbroadcast_colsum", 4, lambda: be.colsum(a, out=cs)),
            ("adam", 28, lambda: be.opt_step(be.OPT_ADAM, a.view((n,)), None, b.view((n,)), m, v, h)),
        ]
        for name, bpe, fn in ops:
            ms = timed(fn, flush)
            gbs = n * bpe / (ms * 1e-3) / 1e9
            rows.append(dict(op=name, log2_n=lg, bytes_per_elem=bpe, ms=ms, gbs=gbs,
                             frac_of_measured_hbm=gbs / hbm, l2_flushed=flush))
        del a, b, out, m, v
    result = dict(hbm_peak_gbs=hbm, rows=rows)
    if args.cpu:
        import ref_numpy as R
        cpu = []
        for lg in (24, 26):
            n = 1 << lg
            x = np.full((n // 1024, 1024), 0.5, np.float32)
            y = np.full((n // 1024, 1024), 0.25, np.float32)
            bias = np.full((1, 1024), 0.125, np.float32)
            tx, ty, tb = R.RefTensor(x), R.RefTensor(y), R.RefTensor(bias, requires_grad=True)
            cases = [("add", 12, lambda: R.add(tx, ty)), ("mul", 12, lambda: R.mul(tx, ty)),
                     ("bias_add", 8, lambda: R.add(tx, tb)), ("relu_fwd", 8, lambda: R.clip(tx, 0.0)),
                     ("unbroadcast_colsum", 4, lambda: R.unbroadcast(x, (1, 1024)))]
            for name, bpe, fn in cases:
                best = 1e30
                for _ in range(3):
                    t0 = time.perf_counter()
                    fn()
                    best = min(best, time.perf_counter() - t0)
                cpu.append(dict(op=name, log2_n=lg, gbs=n * bpe / best / 1e9))
        result["cpu_numpy_oracle"] = cpu
        result["cpu_cores"] = len(os.sched_getaffinity(0))
    print(json.dumps(result))


if __name__ == "__main__":
    main()
