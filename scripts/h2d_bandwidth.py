"""H2D bandwidth of this rank's pinned staging path while the other ranks copy too (torchrun):
prints GB/s for cudaHostAlloc'ed and cudaHostRegister'ed 134 MB buffers."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import core._backend as be  # noqa: E402
import core._dist as dist  # noqa: E402

dist.init_process_group()
n = 8192 * 4096
dev = be.empty((n,), be.F32)
pin = be.PinnedArray((n,), np.float32)
pin.array[:] = 1.0
arr = np.ones(n, np.float32)
reg = be.RegisteredHostArray(arr)
for name, src in (("cudaHostAlloc", pin), ("cudaHostRegister", reg)):
    for _ in range(3):
        be.h2d_prefetch(dev, src)
    be.copy_stream_sync()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        be.h2d_prefetch(dev, src)
    be.copy_stream_sync()
    dt = (time.perf_counter() - t0) / 20
    print("rank %d %s: %.2f ms per 134 MB = %.1f GB/s" % (dist.rank(), name, dt * 1e3, n * 4 / dt / 1e9), flush=True)
    dist.barrier()
dist.destroy_process_group()
