"""Mini-batch iterators.

BatchIterator keeps the reference's interface (utils/data_iterator.py): called with (inputs,
targets) it yields Batch(inputs, targets) namedtuples, reshuffling once per call with numpy's
global generator.  On device Tensors the epoch's permutation is uploaded once (8 bytes per row) and
every mini-batch is "rows perm[start:end] of the data set": gathered by one kernel per tensor when
the batch is used -- or straight into a recorded step's input buffer (Model.train_step) -- so the
shuffled copy of the whole data set the reference makes (`inputs[idx]`, 157 MB for MNIST) never
exists.  Tensors that require grad take the reference's two differentiable getitem steps instead;
on numpy arrays it is plain indexing.

PrefetchIterator is the host-fed variant: the data set stays in (pinned) host memory and batch
i+1 is copied to the GPU on the copy stream while batch i trains, double-buffered.
"""
from collections import namedtuple

import numpy as np

Batch = namedtuple("Batch", ["inputs", "targets"])


class BaseIterator(object):

    def __call__(self, inputs, targets):
        raise NotImplementedError


def _batch_bounds(n_rows, batch_size):
    """[(start, stop)] covering n_rows; the last batch may be short (50000 = 390*128 + 80)"""
    return [(lo, min(lo + batch_size, n_rows)) for lo in range(0, n_rows, batch_size)]


def _plain_device_tensors(*tensors):
    """device Tensors that take no part in autograd (data, not parameters) and have rows to gather"""
    from core.tensor import Tensor
    return all(isinstance(t, Tensor) and not t.requires_grad and t.ndim >= 1 for t in tensors)


class BatchIterator(BaseIterator):

    def __init__(self, batch_size=32, shuffle=True):
        self.batch_size = batch_size
        self.shuffle = shuffle

    def __call__(self, inputs, targets):
        n_rows = len(inputs)
        if self.shuffle:
            order = np.arange(n_rows)
            np.random.shuffle(order)            # the reference's RNG call (data_iterator.py:25-26)
            if _plain_device_tensors(inputs, targets):
                yield from self._gathered_batches(inputs, targets, order)
                return
            inputs, targets = inputs[order], targets[order]
        for lo, hi in _batch_bounds(n_rows, self.batch_size):
            yield Batch(inputs=inputs[lo:hi], targets=targets[lo:hi])

    def _gathered_batches(self, inputs, targets, order):
        """fused `x[idx][start:end]`: each batch is a LazyRows window of the uploaded permutation"""
        import core._backend as be
        from core.tensor import Tensor
        idx_dev = be.upload_index(order)
        for lo, hi in _batch_bounds(len(order), self.batch_size):
            window = idx_dev.view((hi - lo,), lo)
            yield Batch(inputs=Tensor(be.LazyRows(inputs._data, window)),
                        targets=Tensor(be.LazyRows(targets._data, window)))


class PrefetchIterator(BaseIterator):
    """Yields device-Tensor batches from HOST float32 arrays, with the H2D copy of batch i+1
    overlapping the step on batch i.

        it = PrefetchIterator(batch_size=8192)
        for batch in it(x_host, y_host):      # numpy arrays; rows are sharded by the caller
            ...train on batch.inputs / batch.targets ...

    Without shuffling the data set is pinned in place (cudaHostRegister) and every batch is DMA'd
    straight out of it; with shuffling the rows of a batch are first gathered into a pinned staging
    buffer.  Two device buffers per stream of data: a buffer is rewritten two iterations later,
    after the compute stream has passed the kernels that read it.  loop=True wraps around
    indefinitely (the caller breaks).

    num_classes=C: `targets` is the integer class-label vector (n,) a data set stores; 4 bytes per
    sample cross PCIe and the dense one-hot rows the loss takes (run.py:27-28 get_one_hot) are
    written on the device (tnn_one_hot) -- for the 4096-class wide MLP that halves the H2D bytes."""

    def __init__(self, batch_size=32, shuffle=False, loop=False, num_classes=None):
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.loop = loop
        self.num_classes = num_classes
        self._cache = None

    def _buffers(self, inputs, targets):
        import core._backend as be
        key = (id(inputs), id(targets), self.batch_size, self.shuffle, self.num_classes)
        if self._cache is not None and self._cache["key"] == key:
            return self._cache
        xs = (self.batch_size,) + tuple(inputs.shape[1:])
        ys = (self.batch_size,) + tuple(targets.shape[1:])
        t_dt = np.float32 if self.num_classes is None else np.int32
        cache = {"key": key, "dev": [(be.empty(xs, be.F32), be.empty(ys, be.F32)) for _ in range(2)]}
        if self.num_classes is not None:
            # int32 labels travel in a float32-sized device vector (same 4-byte elements); the
            # one-hot rows are written next to them on the device
            cache["onehot"] = [be.empty((self.batch_size, self.num_classes), be.F32) for _ in range(2)]
        if self.shuffle:
            cache["stage"] = [(be.PinnedArray(xs, np.float32), be.PinnedArray(ys, t_dt))
                              for _ in range(2)]
        else:
            cache["pinned"] = (be.RegisteredHostArray(inputs), be.RegisteredHostArray(targets))
        self._cache = cache
        return cache

    def __call__(self, inputs, targets):
        import core._backend as be
        import core.tensor as T
        from core.tensor import Tensor
        inputs = np.require(inputs, np.float32, "C")
        if self.num_classes is None:
            targets = np.require(targets, np.float32, "C")
        else:
            if np.ndim(targets) != 1:
                raise ValueError("num_classes is set: targets must be the 1-D integer label vector")
            targets = np.require(targets, np.int32, "C")
        n_rows = len(inputs)
        bounds = _batch_bounds(n_rows, self.batch_size)
        if not bounds:
            return
        buf = self._buffers(inputs, targets)
        order = np.arange(n_rows)

        def stage(step):
            lo, hi = bounds[step % len(bounds)]
            n = hi - lo
            x_dev, y_dev = buf["dev"][step % 2]
            xv, yv = x_dev.view((n,) + x_dev.shape[1:]), y_dev.view((n,) + y_dev.shape[1:])
            if self.shuffle:
                if step % len(bounds) == 0:
                    np.random.shuffle(order)
                x_pin, y_pin = buf["stage"][step % 2]
                be.copy_stream_sync()               # the previous copy out of this staging pair is done
                x_pin.array[:n] = inputs[order[lo:hi]]
                y_pin.array[:n] = targets[order[lo:hi]]
                be.h2d_prefetch(xv, x_pin)
                be.h2d_prefetch(yv, y_pin)
            else:
                px, py = buf["pinned"]
                be.h2d_prefetch(xv, px.row_address(lo))
                be.h2d_prefetch(yv, py.row_address(lo))
            return xv, yv

        step = 0
        pending = stage(0)
        while True:
            be.wait_prefetch()                       # batch `step` is on the device
            xv, yv = pending
            last = (not self.loop) and step + 1 >= len(bounds)
            if not last:
                pending = stage(step + 1)            # overlaps with the caller's work on this batch
            if self.num_classes is not None:
                n = xv.shape[0]
                oh = buf["onehot"][step % 2].view((n, self.num_classes))
                # the dense rows are only written if somebody asks for them: the fused cross-entropy
                # takes the class indices (be.LazyOneHot)
                yv = be.LazyOneHot(yv, self.num_classes, oh)
            yield Batch(inputs=Tensor(xv), targets=Tensor(yv))
            if last:
                return
            # a training iteration the engine postponed (core/_deferred.py) and the loop did not
            # finish with step() still names this batch's buffers, which the next stage() reuses
            T._flush_deferred()
            step += 1
