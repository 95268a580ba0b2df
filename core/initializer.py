"""Parameter initialisers (interface of the reference's core/initializer.py).

Host-side on purpose: weights are drawn from numpy's global RNG in the same order and with the
same calls as the reference, so a seed reproduces the reference's parameters bit for bit; the
draw is then uploaded once as a float32 device tensor.
"""
import numpy as np
import scipy.stats as stats

from core.tensor import Tensor


def get_fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    return np.prod(shape[1:]), shape[0]


class Initializer(object):

    def __call__(self, shape):
        return Tensor(self.init(shape), requires_grad=True, dtype=np.float32)

    def init(self, shape):
        raise NotImplementedError


class NormalInit(Initializer):

    def __init__(self, mean=0.0, std=1.0):
        self._mean, self._std = mean, std

    def init(self, shape):
        return np.random.normal(loc=self._mean, scale=self._std, size=shape)


class TruncatedNormalInit(Initializer):

    def __init__(self, mean=0.0, std=1.0):
        self._tn = stats.truncnorm(-2 * std, 2 * std, loc=mean, scale=std)

    def init(self, shape):
        return self._tn.rvs(size=shape)


class UniformInit(Initializer):

    def __init__(self, a=0.0, b=1.0):
        self._a, self._b = a, b

    def init(self, shape):
        return np.random.uniform(low=self._a, high=self._b, size=shape)


class ConstantInit(Initializer):

    def __init__(self, val):
        self._val = val

    def init(self, shape):
        return np.full(shape=shape, fill_value=self._val)


class ZerosInit(ConstantInit):

    def __init__(self):
        super(ZerosInit, self).__init__(0.0)


class _Scaled(Initializer):
    """shared machinery of the Glorot / He families: a fan-dependent scale and a distribution"""
    uniform = True

    def __init__(self, gain=1.0):
        self._gain = gain

    def _scale(self, fan_in, fan_out):
        raise NotImplementedError

    def init(self, shape):
        fan_in, fan_out = get_fans(shape)
        s = self._gain * self._scale(fan_in, fan_out)
        if self.uniform:
            return np.random.uniform(low=-s, high=s, size=shape)
        return np.random.normal(loc=0.0, scale=s, size=shape)


class XavierUniformInit(_Scaled):
    """U(-a, a), a = gain * sqrt(6 / (fan_in + fan_out))  (Glorot & Bengio 2010)"""

    def _scale(self, fan_in, fan_out):
        return np.sqrt(6.0 / (fan_in + fan_out))


class XavierNormalInit(_Scaled):
    """N(0, std), std = gain * sqrt(2 / (fan_in + fan_out))"""
    uniform = False

    def _scale(self, fan_in, fan_out):
        return np.sqrt(2.0 / (fan_in + fan_out))


class HeUniformInit(_Scaled):
    """U(-a, a), a = gain * sqrt(6 / fan_in)  (He et al. 2015)"""

    def _scale(self, fan_in, fan_out):
        return np.sqrt(6.0 / fan_in)


class HeNormalInit(_Scaled):
    """N(0, std), std = gain * sqrt(2 / fan_in)"""
    uniform = False

    def _scale(self, fan_in, fan_out):
        return np.sqrt(2.0 / fan_in)
