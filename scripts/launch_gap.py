import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import core._backend as be
be.init()
a = be.zeros((4,), be.F32)
for n in (13, 130):
    g = be.StepGraph()
    with g.capture():
        for _ in range(n):
            be._lib.tnn_fill(0, a.ptr, 1.0, 1)
    for _ in range(20): g.replay()
    e0, e1 = be.Event(), be.Event()
    e0.record()
    for _ in range(200): g.replay()
    e1.record()
    t = e1.elapsed_ms_since(e0) / 200 * 1e3
    print("graph of %d trivial dependent kernels: %.1f us per replay, %.2f us per kernel" % (n, t, t / n))
    g.destroy()
# eager back-to-back trivial launches
e0, e1 = be.Event(), be.Event()
for _ in range(100): be._lib.tnn_fill(0, a.ptr, 1.0, 1)
e0.record()
for _ in range(2000): be._lib.tnn_fill(0, a.ptr, 1.0, 1)
e1.record()
print("eager trivial launches: %.2f us each" % (e1.elapsed_ms_since(e0) / 2000 * 1e3))
