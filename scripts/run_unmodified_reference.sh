#!/bin/bash
# Runs the reference's OWN files, unmodified, against this repository's core/ and utils/ on a GPU:
#   REF_DIR/test/*.py              (pytest; test_autograd.py is the one on the hot path)
#   REF_DIR/examples/mnist/run.py  (two epochs on synthetic MNIST-shaped data)
# REF_DIR is a verbatim, untracked copy of the reference's examples/ and test/ directories made
# only for the duration of the call (the reference tree does not exist on the GPU box and its
# sources are never committed here).  They resolve `core` / `utils` from the working directory
# (test/runtime_path.py:19-28), i.e. this repository.
set -u
REF_DIR=${1:-.ref_tmp}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/make_fake_mnist.py /tmp/fakemnist 12800 2000 > gpurun_out/ref_unmodified.log 2>&1
echo "== pytest $REF_DIR/test (unmodified reference tests)" >> gpurun_out/ref_unmodified.log
timeout 600 python -m pytest -p no:cacheprovider "$REF_DIR/test" -q >> gpurun_out/ref_unmodified.log 2>&1
echo "pytest exit $?" >> gpurun_out/ref_unmodified.log
echo "== python $REF_DIR/examples/mnist/run.py --num_ep 2 --seed 0 (unmodified reference example)" >> gpurun_out/ref_unmodified.log
timeout 600 python "$REF_DIR/examples/mnist/run.py" --num_ep 2 --data_dir /tmp/fakemnist --seed 0 >> gpurun_out/ref_unmodified.log 2>&1
echo "run.py exit $?" >> gpurun_out/ref_unmodified.log
cat gpurun_out/ref_unmodified.log
