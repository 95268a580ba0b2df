"""Host-side profile of the MNIST-shaped step (launch-bound config): cProfile over 300 steps."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import core._backend as be  # noqa: E402
from bench import MNIST, build_model  # noqa: E402
from core.losses import SoftmaxCrossEntropyLoss  # noqa: E402
from core.tensor import Tensor  # noqa: E402

be.init()
rng = np.random.RandomState(0)
x = be.from_numpy(rng.rand(128, 784).astype(np.float32))
y = be.from_numpy(np.eye(10)[rng.randint(0, 10, 128)])
np.random.seed(0)
net, model = build_model(MNIST)
loss_layer = SoftmaxCrossEntropyLoss()


def step():
    model.zero_grad()
    loss = loss_layer.loss(model.forward(Tensor(x)), Tensor(y))
    loss.backward()
    model.step()
    return loss


for _ in range(20):
    step()
be.sync()
t0 = time.perf_counter()
for _ in range(300):
    step()
t1 = time.perf_counter()
be.sync()
t2 = time.perf_counter()
print("host enqueue %.1f us/step, drained after %.1f us/step" % ((t1 - t0) / 300 * 1e6, (t2 - t0) / 300 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
be.sync()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
