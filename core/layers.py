"""Network layers and activation layers (interface of the reference's core/layers.py).

Dense keeps the reference contract -- lazy initialisation from inputs.shape[1], `w` of shape
(in, out) and `b` of shape (1, out) as float32 tensors, `self.inputs` recorded -- but its forward
is one fused GEMM+bias node (ops.dense_) instead of `inputs @ w + b`; ReLU is ops.clip(x, 0.0)
exactly as in the reference, which the op library routes to the fused ReLU kernels.
"""
import core.ops as ops
from core.initializer import XavierUniformInit
from core.initializer import ZerosInit


class Layer(object):

    def __init__(self, name):
        self.name = name
        self.params, self.grads = {}, {}
        self.is_training = True

    def forward(self, inputs):
        raise NotImplementedError

    def set_phase(self, phase):
        self.is_training = (phase == "TRAIN")


class Dense(Layer):

    def __init__(self, num_out, num_in=None, w_init=XavierUniformInit(), b_init=ZerosInit()):
        super().__init__("Linear")
        self.initializers = {"w": w_init, "b": b_init}
        self.shapes = {"w": [num_in, num_out], "b": [1, num_out]}
        self.params = {"w": None, "b": None}
        self.is_init = False
        if num_in is not None:
            self._init_parameters(num_in)
        self.inputs = None

    def forward(self, inputs):
        if not self.is_init:  # layers.py:45-46
            self._init_parameters(inputs.shape[1])
        self.inputs = inputs
        return ops.dense_(inputs, self.params["w"], self.params["b"])

    def forward_fused_relu(self, inputs, relu_layer):
        """this layer followed by `relu_layer` in one GEMM launch (core.nn.Net calls this when a
        Dense is directly followed by a ReLU); both layers record their inputs as usual"""
        if not self.is_init:
            self._init_parameters(inputs.shape[1])
        self.inputs = inputs
        z, a = ops.dense_relu_(inputs, self.params["w"], self.params["b"])
        relu_layer.inputs = z
        return a

    def _init_parameters(self, input_size):
        # layers.py:51-57; draw order (w then b) fixes the numpy RNG stream
        self.shapes["w"][0] = input_size
        for key in ("w", "b"):
            self.params[key] = self.initializers[key](shape=self.shapes[key])
            self.params[key].zero_grad()
        self.is_init = True


class Activation(Layer):

    def __init__(self, name):
        super().__init__(name)
        self.inputs = None

    def forward(self, inputs):
        self.inputs = inputs
        return self.func(inputs)

    def func(self, x):
        raise NotImplementedError


class Sigmoid(Activation):
    """The reference's Sigmoid calls np.exp on a Tensor and raises (layers.py:79-80);
    nothing pins that, so this one computes 1 / (1 + exp(-x)) with the op library."""

    def __init__(self):
        super().__init__("Sigmoid")

    def func(self, x):
        return 1.0 / (1.0 + ops.exp(-x))


class Tanh(Activation):
    """(1 - e^-x) / (1 + e^-x), i.e. tanh(x / 2), as written at layers.py:88-89"""

    def __init__(self):
        super().__init__("Tanh")

    def func(self, x):
        e = ops.exp(-x)
        return (1.0 - e) / (1.0 + e)


class ReLU(Activation):

    def __init__(self):
        super().__init__("ReLU")

    def func(self, x):
        return ops.clip(x, 0.0)
