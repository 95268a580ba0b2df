"""Numerical soak of the wide MLP (4 x Dense(4096), batch 8192, Adam 1e-3) on a FIXED pair of batches:
the loss every 25 steps for 300 steps.  Run once per operand split (TNN_GEMM_SPLIT=f16|mix|tf32x3) to
compare the trajectories of the splits.  With the default split it also reports in how many steps,
and for how many products, the on-device guard sent a product to the mixed-split fallback (timing of
the conditional launches: a real one takes ~0.9 ms, one that returns at once ~4 us)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import core._backend as be  # noqa: E402
from bench import WIDE, build_model  # noqa: E402
from core.tensor import Tensor  # noqa: E402

be.init()
rng = np.random.RandomState(0)
B, D, C = 8192, 4096, 4096
xs, ys = [], []
for k in range(2):
    xs.append(Tensor(rng.rand(B, D).astype(np.float32)))
    y = np.zeros((B, C), np.float32)
    y[np.arange(B), rng.randint(0, C, B)] = 1.0
    ys.append(Tensor(y))
np.random.seed(0)
net, model = build_model(WIDE)
out = []
fallback_steps, fallback_products = [], 0
count_fallbacks = be.TC_SPLIT == "f16"
for step in range(301):
    if count_fallbacks:
        be.prof_enable(4)
    model.zero_grad()
    loss = model.loss.loss(model.forward(xs[step % 2]), ys[step % 2])
    loss.backward()
    model.step()
    if count_fallbacks:
        ms, n = be.prof_collect()
        be.prof_enable(0)
        k = int(round(max(0.0, ms - 0.006 * n) / 0.9))
        if k:
            fallback_steps.append((step, k))
            fallback_products += k
    if step % 25 == 0:
        out.append((step, float(loss.values)))
psum = [float(np.abs(p.values).mean()) for layer in net.get_parameters() for p in layer.values()]
print(json.dumps({"split": be.TC_SPLIT, "losses": out, "fallback_steps": fallback_steps,
                  "fallback_products": fallback_products, "products": 11 * 301, "mean_abs_param": psum,
                  "finite": bool(np.all(np.isfinite([l for _, l in out])))}))
