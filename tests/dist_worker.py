"""Worker for tests/test_gpu_dist.py (launched by torch.distributed.run, one rank per GPU):
trains a small MLP data-parallel through the public API and checks, on rank 0, that losses and
parameters equal the single-process oracle on the FULL batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import core._backend as be  # noqa: E402
import core._dist as dist  # noqa: E402
import ref_numpy as R  # noqa: E402
from core.layers import Dense, ReLU  # noqa: E402
from core.losses import SoftmaxCrossEntropyLoss  # noqa: E402
from core.model import Model  # noqa: E402
from core.nn import Net  # noqa: E402
from core.optimizer import SGD  # noqa: E402
from core.tensor import Tensor  # noqa: E402


def main():
    dist.init_process_group()
    rank, world = dist.rank(), dist.world_size()
    rng = np.random.RandomState(0)
    B, D, C = 96, 40, 12
    x = rng.rand(B, D).astype(np.float32)
    labels = np.eye(C, dtype=np.float32)[rng.randint(0, C, B)]
    lo, hi = dist.shard_bounds(B, rank, world)

    np.random.seed(0)
    net = Net([Dense(24), ReLU(), Dense(C)])
    model = Model(net=net, loss=SoftmaxCrossEntropyLoss(), optimizer=SGD(lr=0.5))
    loss_layer = SoftmaxCrossEntropyLoss()
    losses = []
    for _ in range(4):
        model.zero_grad()
        loss = loss_layer.loss(model.forward(Tensor(x[lo:hi])), Tensor(labels[lo:hi]))
        loss.backward()
        model.step()
        losses.append(float(loss.values))
    params = [p.values.copy() for layer in net.get_parameters() for p in layer.values()]

    # every rank must hold bit-identical parameters after the all-reduced update
    flat = np.concatenate([p.ravel() for p in params]).astype(np.float32)
    mine = be.from_numpy(flat)
    dist.allreduce_sum(mine)
    assert np.array_equal(mine.numpy(), flat * world), "replicas diverged"

    if rank == 0:
        np.random.seed(0)
        ref = R.RefMLP([24, C], R.RefSGD(lr=0.5))
        rlosses = [float(ref.train_step(x, labels)) for _ in range(4)]
        assert np.max(np.abs(np.array(losses) - np.array(rlosses))) <= 1e-5, (losses, rlosses)
        for p, rp in zip(params, ref.params()):
            err = np.max(np.abs(p - rp.values)) / max(np.max(np.abs(rp.values)), 1e-30)
            assert err <= 1e-5, err
        print("DIST_OK world=%d losses=%s" % (world, losses))
    # the captured step (CUDA graph holding the NCCL all-gather / all-reduces) must reproduce the
    # eager data-parallel loop bit for bit
    def fresh():
        np.random.seed(1)
        n = Net([Dense(24, num_in=D), ReLU(), Dense(C, num_in=24)])
        return n, Model(net=n, loss=SoftmaxCrossEntropyLoss(), optimizer=SGD(lr=0.5))

    xs, ys = Tensor(x[lo:hi]), Tensor(labels[lo:hi])
    net_e, model_e = fresh()
    eager = []
    for _ in range(6):
        model_e.zero_grad()
        loss = model_e.loss.loss(model_e.forward(xs), ys)
        loss.backward()
        model_e.step()
        eager.append(float(loss.values))
    net_g, model_g = fresh()
    replayed = [float(model_g.train_step(xs, ys).values) for _ in range(6)]
    assert any(hasattr(v, "graph") for v in model_g._captured.values()), "step was not captured"
    assert eager == replayed, (eager, replayed)
    for pe, pg in zip(net_e.get_parameters(), net_g.get_parameters()):
        for k in pe:
            assert np.array_equal(pe[k].values, pg[k].values)
    if rank == 0:
        print("GRAPH_DIST_OK world=%d" % world)

    # chunked all-reduce pipelined with the optimiser kernel (what the wide MLP uses): force it on
    # this small model; with two ranks a SUM is order-independent, so it must match bit for bit
    from core.optimizer import Adam
    results = []
    for min_elems, n_chunks in ((1 << 20, 1), (64, 4)):
        dist.MIN_CHUNK_ELEMS, dist.ALLREDUCE_CHUNKS = min_elems, n_chunks
        np.random.seed(2)
        n = Net([Dense(24, num_in=D), ReLU(), Dense(C, num_in=24)])
        m = Model(net=n, loss=SoftmaxCrossEntropyLoss(), optimizer=Adam(lr=1e-2))
        ls = []
        for _ in range(5):
            m.zero_grad()
            loss = m.loss.loss(m.forward(xs), ys)
            loss.backward()
            m.step()
            ls.append(float(loss.values))
        results.append((ls, [p.values.copy() for layer in n.get_parameters() for p in layer.values()]))
    dist.MIN_CHUNK_ELEMS, dist.ALLREDUCE_CHUNKS = 1 << 20, 1
    assert results[0][0] == results[1][0], (results[0][0], results[1][0])
    for a, b in zip(results[0][1], results[1][1]):
        assert np.array_equal(a, b)
    if rank == 0:
        print("CHUNKED_DIST_OK world=%d" % world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
