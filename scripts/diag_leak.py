"""GPU diagnostic: does a training step leave device blocks behind that only the cyclic garbage
collector frees?  Prints pool statistics per step and the types found in reference cycles."""
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import core._backend as be  # noqa: E402
from core.tensor import Tensor  # noqa: E402


def main(workload):
    cfg = dict(bench.WIDE if workload == "wide" else bench.MNIST)
    B, C = (2048 if workload == "wide" else 128), cfg["widths"][-1]
    xs, ys = bench.synthetic_shard(cfg, B, 0, copies=2)
    x_dev = [be.from_numpy(x) for x in xs]
    y_dev = [be.from_numpy(bench.one_hot_host(y, C)) for y in ys]
    stepper = bench.Stepper(cfg, use_graph=False)
    gc.collect()
    gc.disable()
    for i in range(12):
        s0 = be.pool_stats()
        stepper(Tensor(x_dev[i % 2]), Tensor(y_dev[i % 2]))
        s1 = be.pool_stats()
        print(workload, "step", i, "mallocs", s1["cuda_mallocs"] - s0["cuda_mallocs"],
              "in_use MB %.1f" % (s1["in_use"] / 1e6), "reserved MB %.1f" % (s1["reserved"] / 1e6))
    gc.set_debug(gc.DEBUG_SAVEALL)
    n = gc.collect()
    s2 = be.pool_stats()
    print(workload, "gc.collect() found", n, "objects; in_use MB after %.1f" % (s2["in_use"] / 1e6))
    kinds = {}
    for o in gc.garbage:
        kinds[type(o).__name__] = kinds.get(type(o).__name__, 0) + 1
    print(sorted(kinds.items(), key=lambda kv: -kv[1])[:12])
    for o in gc.garbage:
        if type(o).__name__ in ("function", "cell") and len(kinds) < 40:
            pass
    fns = [o for o in gc.garbage if type(o).__name__ == "function"]
    print("functions in cycles:", sorted({f.__qualname__ for f in fns})[:20])
    gc.set_debug(0)
    gc.garbage.clear()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "wide")
