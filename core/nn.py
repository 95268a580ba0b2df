"""Feed-forward network container (interface of the reference's core/nn.py)."""
import os

from core.layers import Dense
from core.layers import ReLU

FUSE_DENSE_RELU = os.environ.get("TNN_FUSE_RELU", "1") != "0"


class Net(object):

    def __init__(self, layers):
        self.layers = layers
        self._phase = "TRAIN"

    def forward(self, inputs):
        # nn.py:10-13 is a plain loop over layer.forward; a Dense directly followed by a ReLU is
        # executed as one fused launch (same graph, same recorded layer inputs)
        out = inputs
        layers = self.layers
        i, n = 0, len(layers)
        while i < n:
            layer = layers[i]
            if FUSE_DENSE_RELU and type(layer) is Dense and i + 1 < n and type(layers[i + 1]) is ReLU:
                out = layer.forward_fused_relu(out, layers[i + 1])
                i += 2
            else:
                out = layer.forward(out)
                i += 1
        return out

    def get_parameters(self):
        return [layer.params for layer in self.layers]

    def set_parameters(self, params):
        for layer, new in zip(self.layers, params):
            assert layer.params.keys() == new.keys()
            for key in layer.params.keys():
                assert layer.params[key].shape == new[key].shape
                layer.params[key] = new[key]

    def get_phase(self):
        return self._phase

    def set_phase(self, phase):
        for layer in self.layers:
            layer.set_phase(phase)
        self._phase = phase
