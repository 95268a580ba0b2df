"""Mini-batch iterator (interface of the reference's utils/data_iterator.py).

Works on numpy arrays and on device Tensors alike: the per-epoch shuffle `inputs[idx]` is one
row-gather kernel on a Tensor, the per-step `inputs[start:end]` a zero-copy view.
"""
from collections import namedtuple

import numpy as np

Batch = namedtuple("Batch", ["inputs", "targets"])


class BaseIterator(object):

    def __call__(self, inputs, targets):
        raise NotImplementedError


class BatchIterator(BaseIterator):

    def __init__(self, batch_size=32, shuffle=True):
        self.batch_size = batch_size
        self.shuffle = shuffle

    def __call__(self, inputs, targets):
        n = len(inputs)
        if self.shuffle:
            order = np.arange(n)
            np.random.shuffle(order)  # same RNG call as data_iterator.py:25-26
            inputs, targets = inputs[order], targets[order]
        for start in range(0, n, self.batch_size):
            stop = start + self.batch_size
            yield Batch(inputs=inputs[start:stop], targets=targets[start:stop])
