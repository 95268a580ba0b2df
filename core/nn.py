"""Sequential network container (interface of the reference's core/nn.py: Net(layers) with
forward / get_parameters / set_parameters / get_phase / set_phase).

forward() is where the engine departs from a plain `for layer in layers` loop: a Dense that is
directly followed by a ReLU runs as ONE GEMM launch whose epilogue also produces the ReLU output
and (on the tensor-core path) that output's tf32 planes for the next layer.  The autograd graph,
the values, the gradients and the `inputs` each layer records are the same as for the unfused loop
(tests/test_gpu_train.py::test_dense_relu_fusion_is_transparent).
"""
import os

from core.layers import Dense
from core.layers import ReLU

FUSE_DENSE_RELU = os.environ.get("TNN_FUSE_RELU", "1") != "0"


def _fusable(layer, follower):
    return FUSE_DENSE_RELU and type(layer) is Dense and type(follower) is ReLU


class Net(object):

    def __init__(self, layers):
        self.layers = layers
        self._phase = "TRAIN"

    # -- execution ---------------------------------------------------------------------------
    def forward(self, inputs):
        layers = self.layers
        out, i = inputs, 0
        while i < len(layers):
            follower = layers[i + 1] if i + 1 < len(layers) else None
            if _fusable(layers[i], follower):
                out = layers[i].forward_fused_relu(out, follower)
                i += 2
            else:
                out = layers[i].forward(out)
                i += 1
        return out

    __call__ = forward

    # -- parameters ----------------------------------------------------------------------------
    def get_parameters(self):
        """one dict per layer (empty for activations), the live objects -- model.py:47 mutates them"""
        return [layer.params for layer in self.layers]

    def set_parameters(self, params):
        if len(params) != len(self.layers):
            raise ValueError("expected %d parameter dicts, got %d" % (len(self.layers), len(params)))
        for layer, new in zip(self.layers, params):
            assert layer.params.keys() == new.keys()
            for key, tensor in new.items():
                assert layer.params[key].shape == tensor.shape
                layer.params[key] = tensor

    def num_parameters(self):
        total = 0
        for layer in self.layers:
            for p in layer.params.values():
                if p is not None:
                    total += int(p._data.size)
        return total

    # -- phase -----------------------------------------------------------------------------------
    def get_phase(self):
        return self._phase

    def set_phase(self, phase):
        self._phase = phase
        for layer in self.layers:
            layer.set_phase(phase)
