"""Layers (interface of the reference's core/layers.py: Layer, Dense, Activation, Sigmoid, Tanh,
ReLU; `params` dicts, `inputs` bookkeeping, lazy Dense initialisation).

Dense is one fused GEMM(+bias) node; together with a following ReLU it is one launch (see
core/nn.py).  Activations are compositions of the op library, which routes `clip(x, 0.0)` to the
fused ReLU kernels.
"""
import core.ops as ops
from core.initializer import XavierUniformInit
from core.initializer import ZerosInit


class Layer(object):
    """Base class: a name, a dict of parameter tensors and the TRAIN/TEST flag."""

    def __init__(self, name):
        self.name = name
        self.params = {}
        self.grads = {}
        self.is_training = True

    def forward(self, inputs):
        raise NotImplementedError

    def __call__(self, inputs):
        return self.forward(inputs)

    def set_phase(self, phase):
        self.is_training = (phase == "TRAIN")


class Dense(Layer):
    """y = x @ w + b with w of shape (num_in, num_out) and b of shape (1, num_out), both float32
    device tensors.  num_in may be left out: it is then taken from the first input (layers.py:45-46),
    which also fixes WHEN the weights are drawn from numpy's global RNG."""

    def __init__(self, num_out, num_in=None, w_init=XavierUniformInit(), b_init=ZerosInit()):
        Layer.__init__(self, "Linear")
        self.initializers = {"w": w_init, "b": b_init}
        self.shapes = {"w": [num_in, num_out], "b": [1, num_out]}
        self.params = {"w": None, "b": None}
        self.inputs = None
        self.is_init = False
        if num_in is not None:
            self._init_parameters(num_in)

    def _init_parameters(self, input_size):
        self.shapes["w"][0] = input_size
        for key in ("w", "b"):          # draw order w, b: layers.py:53-56
            tensor = self.initializers[key](shape=self.shapes[key])
            tensor.zero_grad()
            self.params[key] = tensor
        self.is_init = True

    def _prepare(self, inputs):
        if not self.is_init:
            self._init_parameters(inputs.shape[1])
        self.inputs = inputs
        return self.params["w"], self.params["b"]

    def forward(self, inputs):
        w, b = self._prepare(inputs)
        return ops.dense_(inputs, w, b)

    def forward_fused_relu(self, inputs, relu_layer):
        """this layer followed by `relu_layer` in one GEMM launch; both record their inputs"""
        w, b = self._prepare(inputs)
        pre_activation, activated = ops.dense_relu_(inputs, w, b)
        relu_layer.inputs = pre_activation
        return activated


class Activation(Layer):

    def __init__(self, name):
        Layer.__init__(self, name)
        self.inputs = None

    def forward(self, inputs):
        self.inputs = inputs
        return self.func(inputs)

    def func(self, x):
        raise NotImplementedError


class ReLU(Activation):
    """clip(x, 0.0): the backward mask is x >= 0, so ReLU'(0) = 1 (layers.py:97-98, ops.py:336-343)"""

    def __init__(self):
        Activation.__init__(self, "ReLU")

    def func(self, x):
        return ops.clip(x, 0.0)


class Tanh(Activation):
    """(1 - e^-x) / (1 + e^-x), which is tanh(x / 2) -- the formula at layers.py:88-89"""

    def __init__(self):
        Activation.__init__(self, "Tanh")

    def func(self, x):
        decay = ops.exp(-x)
        return (1.0 - decay) / (1.0 + decay)


class Sigmoid(Activation):
    """1 / (1 + e^-x).  The reference's version calls np.exp on a Tensor and raises
    (layers.py:79-80); no test pins that, so this one works."""

    def __init__(self):
        Activation.__init__(self, "Sigmoid")

    def func(self, x):
        return 1.0 / (1.0 + ops.exp(-x))
