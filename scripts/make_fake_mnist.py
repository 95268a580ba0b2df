"""Writes a synthetic MNIST-shaped `mnist.pkl.gz` in the format the reference's unmodified
examples/mnist/run.py loads (run.py:41-51): a gzip-pickled 3-tuple (train, valid, test), each
(X float32 [N, 784], y int64 [N]).  With the file present run.py skips its download
(utils/downloader.py:20-21).

    python scripts/make_fake_mnist.py DIR [n_train] [n_eval]"""
import gzip
import os
import pickle
import sys

import numpy as np

out_dir = sys.argv[1]
n_train = int(sys.argv[2]) if len(sys.argv) > 2 else 12800
n_eval = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
rng = np.random.RandomState(0)
# ten noisy class prototypes, so that the example's accuracy visibly rises above chance
protos = rng.rand(10, 784).astype(np.float32)


def split(n):
    y = rng.randint(0, 10, n).astype(np.int64)
    x = (0.6 * protos[y] + 0.4 * rng.rand(n, 784)).astype(np.float32)
    return x, y


os.makedirs(out_dir, exist_ok=True)
with gzip.open(os.path.join(out_dir, "mnist.pkl.gz"), "wb", compresslevel=1) as f:
    pickle.dump((split(n_train), split(n_eval), split(n_eval)), f, protocol=2)
print("wrote", os.path.join(out_dir, "mnist.pkl.gz"))
