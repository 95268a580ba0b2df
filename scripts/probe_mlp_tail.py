import sys, numpy as np
sys.path.insert(0, ".")
import bench, core._backend as be
from core.tensor import Tensor
cfg = dict(bench.MNIST)
xs, ys = bench.synthetic_shard(cfg, 128, 0, copies=2)
x = [be.from_numpy(v) for v in xs]; y = [be.from_numpy(bench.one_hot_host(v, 10)) for v in ys]
st = bench.Stepper(cfg, True)
for i in range(30):
    st(Tensor(x[i % 2]), Tensor(y[i % 2]))
be.sync()
cap = [s for s in st.model._captured.values() if hasattr(s, "graph")][0]
c = cap.keepalive["tail"].counters.numpy().view(np.uint32)
print("probe deltas (clk): staged, fwd, ce, bwd, rendezvous, tiles:", c[5:11].tolist(), "sum", int(c[5:11].sum()))
