import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import core._backend as be, sweep_ops
be.init()
l2 = be.device_info()["l2_bytes"]
for lg in (20, 22, 24, 26):
    n = 1 << lg
    ns = int(min(48, max(2, -(-2 * l2 // (n * 4)))))
    sets = [dict(a=be.full((n // 1024, 1024), 0.5, be.F32), cs=be.empty((1, 1024), be.F32)) for _ in range(ns)]
    ms = sweep_ops.timed_rotating(lambda k: be.colsum(sets[k]["a"], out=sets[k]["cs"]), ns)
    print("2^%d colsum %.1f us %.0f GB/s (%.0f%%)" % (lg, ms * 1e3, n * 4 / ms / 1e6, n * 4 / ms / 1e6 / 65.53))
