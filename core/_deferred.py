"""The reference's five-line training iteration, run as one recorded step without changing the loop.

examples/mnist/run.py:78-83 writes every iteration as

    model.zero_grad(); pred = model.forward(x); loss = loss_layer.loss(pred, y)
    loss.backward(); model.step()

On this engine that is a dozen launches of 4-10 us kernels and the loop is bound by the host's cost per
launch.  `Model.train_step` records the same iteration into a CUDA graph, but it is a new entry
point the unmodified example does not call.  This module lets the five lines themselves reach the
recording: inside a training loop `Model.forward` hands back a Tensor whose values have not been
computed yet, `SoftmaxCrossEntropyLoss.loss` and `backward()` on it only note what was asked for,
and `Model.step()` -- once it has seen the same iteration shape run eagerly -- replays the recorded
step (forward, loss, backward and update) in one launch and fills the loss in.

Nothing observable changes.  Whatever touches a deferred tensor in any other way (its values, an
operator, its gradient, a second backward, a different loss) first runs the postponed lines eagerly,
in order, exactly as written, and the tensor becomes an ordinary one.  After a replayed step the
prediction's values and gradient are served from the recording's buffers (copied out when somebody
asks, or just before the next replay overwrites them if the object is still alive); the only
difference to the eager loop is that such a prediction carries no autograd graph any more and the
layers' `inputs` bookkeeping is not refreshed -- the same as documented for `train_step`.

`Model.defer_loop = False` (or TNN_DEFER_LOOP=0) turns it off.
"""
import os
import weakref

import core._backend as be
import core.tensor as T
from core.tensor import Tensor

ENABLED = os.environ.get("TNN_DEFER_LOOP", "1") != "0"


class Chain(object):
    """One postponed iteration: which of its lines have been requested so far."""
    __slots__ = ("model", "x", "pred", "loss_obj", "y", "loss", "stage", "done")

    def __init__(self, model, x, pred):
        self.model, self.x, self.pred = model, x, pred
        self.loss_obj = self.y = self.loss = None
        self.stage = "forward"           # -> "loss" -> "backward"
        self.done = False

    def materialise(self):
        """run the postponed lines eagerly, in order; the lazy tensors become ordinary ones"""
        if self.done:
            return
        self.done = True
        if T._DEFERRED[0] is self:
            T._DEFERRED[0] = None
        real = self.model.net.forward(self.x)
        self.model._note_output(self.x, real)
        self.pred._adopt(real)
        if self.stage != "forward":
            self.loss._adopt(self.loss_obj.loss(self.pred, self.y))
            if self.stage == "backward":
                self.loss.backward()


class LazyTensor(Tensor):
    """A Tensor of known shape whose storage does not exist yet.

    pending: part of the Chain in core.tensor._DEFERRED; any use resolves the chain eagerly.
    served : the recorded step ran; values / gradient are copied out of the recording's buffers on
             first use (`_lz_source()` -> (values DArray, gradient DArray))."""

    def __init__(self, chain, shape, dtype):
        d = self.__dict__
        d["_lz_state"] = "pending"
        d["_lz_chain"] = chain
        d["_lz_shape"] = tuple(shape)
        d["_lz_dtype"] = dtype
        d["_lz_source"] = None
        d["_host"] = None
        d["_grad"] = None
        d["_grad_zero"] = True
        d["_grad_host"] = None
        d["_gslot"] = None
        d["_relu_pre"] = None
        d["_fused_bwd"] = None
        d["requires_grad"] = True
        d["dependency"] = []

    # -- what is known without computing anything
    @property
    def shape(self):
        return self._lz_shape

    @property
    def dtype(self):
        return self._lz_dtype

    @property
    def ndim(self):
        return len(self._lz_shape)

    # -- everything else resolves
    def _resolve(self):
        if self._lz_state == "pending":
            self._lz_chain.materialise()
        else:
            self._pin()
        if type(self) is LazyTensor:     # the postponed lines raised earlier and never produced it
            raise RuntimeError("this tensor belongs to a postponed training iteration whose lines "
                               "raised when they were run; it has no values")

    @property
    def _data(self):
        self._resolve()                  # the class is Tensor afterwards: plain attribute from here on
        return self.__dict__["_data"]

    @_data.setter
    def _data(self, value):              # a setter makes the property win over the instance dict
        self._resolve()
        self.__dict__["_data"] = value

    @property
    def grad(self):
        self._resolve()
        return self.grad

    @grad.setter
    def grad(self, value):
        self._resolve()
        self.grad = value

    def zero_grad(self):
        self._resolve()
        self.zero_grad()

    def backward(self, grad=None):
        ch = self._lz_chain
        if (self._lz_state == "pending" and grad is None and not ch.done and ch.loss is self
                and ch.stage == "loss" and T._DEFERRED[0] is ch):
            ch.stage = "backward"        # noted; Model.step() decides how it runs
            return
        self._resolve()
        self.backward(grad)

    # -- becoming an ordinary Tensor
    def _become(self, attrs):
        self.__class__ = Tensor
        self.__dict__.clear()
        self.__dict__.update(attrs)

    def _adopt(self, real):
        """take over a computed tensor's storage, graph and gradient state"""
        self._become(real.__dict__)

    def _plain(self, data, grad):
        attrs = dict(_data=data, _host=None, _grad=grad, _grad_zero=False, _grad_host=None, _gslot=None,
                     _relu_pre=None, _fused_bwd=None, requires_grad=True, dependency=[])
        self._become(attrs)

    def _serve(self, source):
        self.__dict__["_lz_state"] = "served"
        self.__dict__["_lz_chain"] = None
        self.__dict__["_lz_source"] = source

    def _pin(self):
        """copy values and gradient out of the recording's buffers (they are rewritten by its next
        replay) and become an ordinary tensor without a graph"""
        data, grad = self._lz_source()
        self._plain(be.clone(data), be.clone(grad))


def pending_chain():
    return T._DEFERRED[0]


def begin(model, x, out_shape, out_dtype):
    """Model.forward in a training loop: the prediction as a LazyTensor"""
    T._flush_deferred()
    chain = Chain(model, x, None)
    chain.pred = LazyTensor(chain, out_shape, out_dtype)
    T._DEFERRED[0] = chain
    return chain.pred


def defer_loss(loss_obj, logits, labels):
    """SoftmaxCrossEntropyLoss.loss on a pending prediction: the loss as a LazyTensor, or None when
    this call is not the second line of a postponed iteration"""
    ch = T._DEFERRED[0]
    if (ch is None or ch.done or ch.pred is not logits or ch.stage != "forward"
            or type(logits) is not LazyTensor or logits._lz_state != "pending"):
        return None
    if getattr(loss_obj, "_weight", None) is not None:
        return None
    y = labels
    if type(y) is not Tensor or y.requires_grad or y.dtype not in (be.F32, be.F64):
        return None
    if logits.ndim != 2 or tuple(y.shape) != tuple(logits.shape):
        return None
    ch.loss_obj, ch.y = loss_obj, y
    ch.loss = LazyTensor(ch, (), logits.dtype)
    ch.stage = "loss"
    return ch.loss


def serve_after_replay(chain, step, loss_data):
    """the recorded step ran in place of the chain's lines: hand out its results"""
    chain.done = True
    # the loss: its value (already a private copy) and the seed gradient backward() gives it
    chain.loss._plain(loss_data, be.ones_scalar(loss_data.dtype))
    pred = chain.pred
    pred._serve(step.prediction_source)
    step.live_prediction = weakref.ref(pred)


def pin_live_prediction(step):
    """before a replay overwrites the recording's buffers: a prediction of the previous replay that
    somebody still holds and has not read takes its copy now"""
    ref = getattr(step, "live_prediction", None)
    if ref is None:
        return
    step.live_prediction = None
    t = ref()
    if t is not None and type(t) is LazyTensor and t._lz_state == "served":
        t._pin()
