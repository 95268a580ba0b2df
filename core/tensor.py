"""Device-resident Tensor with reverse-mode autograd.

Same public surface as the reference's core/tensor.py (Tensor(values, requires_grad, dependency,
dtype), .values, .grad, .shape, operator overloads, .backward(), .zero_grad(), as_tensor), but the
storage lives in B200 HBM and every operator is a CUDA kernel reached through core/_backend.py.

Differences from the reference that are deliberate:
  * backward() is one topological sweep instead of the reference's per-path recursion
    (tensor.py:157-168): identical gradients, each grad_fn runs once.
  * a gradient has its tensor's dtype (the reference makes every gradient float64); float32
    parameters stay float32.  Python scalars adopt the dtype of the tensor they meet.
  * integer / bool inputs are stored as float64.
  * .values and .grad return host copies (numpy arrays) marked READ-ONLY: the reference hands out
    its live arrays (tensor.py:20,31-33), so `t.values[i] = v` or `p.grad *= 0.5` edit the tensor
    there; here such an edit could not reach the device, so it raises instead of being lost.
    Writes go through the setters (`t.values = ...`, `p.grad = ...`).
"""
import numpy as np

import core._backend as be
import core.ops as ops

_SCALAR_TYPES = (int, float, bool, np.integer, np.floating, np.bool_)


# the postponed training iteration, if any (core/_deferred.py): anything that reads or changes a
# gradient, or changes a tensor's storage, first runs it
_DEFERRED = [None]


def _flush_deferred():
    chain = _DEFERRED[0]
    if chain is not None:
        _DEFERRED[0] = None
        chain.materialise()


def _readonly(arr):
    arr.setflags(write=False)
    return arr


def as_tensor(obj, like=None):
    """Coerce to Tensor (tensor.py:7-10).  A Python/numpy scalar meeting a tensor takes that
    tensor's dtype (weak scalar typing)."""
    if isinstance(obj, Tensor):
        return obj
    if like is not None and isinstance(obj, _SCALAR_TYPES):
        return Tensor(be.full((), obj, like.dtype))
    return Tensor(obj)


class Tensor(object):

    __array_priority__ = 1000  # ndarray <op> Tensor defers to the reflected overloads below

    def __init__(self, values, requires_grad=False, dependency=None, dtype=None):
        if isinstance(values, be.DArray):
            data = values if dtype is None else be.astype(values, be.device_dtype(dtype))
        elif isinstance(values, Tensor):
            data = values._data if dtype is None else be.astype(values._data, be.device_dtype(dtype))
        else:
            data = be.from_numpy(np.asarray(values, dtype))
        self._data = data
        self._host = None        # cached host copy of the values
        self._grad = None        # device gradient (DArray) once something was accumulated
        self._grad_zero = False  # True = "gradient is all zeros" without having allocated it
        self._grad_host = None
        self._gslot = None       # view into a flat gradient arena (set by core.model.Model)
        self._relu_pre = None    # pre-activation array when this tensor is a fused ReLU output
        self._fused_bwd = None   # optional: all input gradients of this node from one launch (ops._dense_node)
        self.requires_grad = requires_grad
        if self.requires_grad:
            self._grad_zero = True   # zero_grad() of a tensor that has no arena slot yet
        self.dependency = dependency if dependency is not None else []

    # ------------------------------------------------------------------ storage
    @property
    def values(self):
        """host copy of the tensor (numpy array); one D2H transfer, cached until storage changes"""
        if self._host is None:
            self._host = _readonly(be.to_numpy(self._data))
        return self._host

    @values.setter
    def values(self, new_values):
        # tensor.py:35-38: replaces the storage and drops the gradient
        if _DEFERRED[0] is not None:
            _flush_deferred()
        if isinstance(new_values, be.DArray):
            self._data = new_values
        elif isinstance(new_values, Tensor):
            self._data = new_values._data
        else:
            self._data = be.from_numpy(np.asarray(new_values))
        self._host = None
        self._drop_grad()

    def values_async(self):
        """Start copying the values to the host without waiting for device work queued later; the
        returned object's .result() gives the numpy array.  (Not in the reference's interface: there
        `loss.values` is already host memory.  Here it lets a loop log step i's loss while step i+1
        is running instead of draining the GPU once per step.)"""
        return be.AsyncRead(self._data)

    def _drop_grad(self):
        self._grad = None
        self._grad_zero = False
        self._grad_host = None

    def _touch(self):
        """storage was modified in place on the device (fused optimizer step)"""
        self._host = None
        self._data.split = None

    @property
    def grad(self):
        """host copy of the accumulated gradient, or None (tensor.py:22)"""
        if _DEFERRED[0] is not None:
            _flush_deferred()
        if self._grad_zero:      # zeroed and nothing accumulated since (an arena slot is not
            if self._grad_host is None:   # cleared until something is written or the step needs it)
                self._grad_host = _readonly(np.zeros(self.shape, dtype=self._data.dtype))
            return self._grad_host
        if self._grad is None:
            return None
        if self._grad_host is None:
            self._grad_host = _readonly(be.to_numpy(self._grad))
        return self._grad_host

    @grad.setter
    def grad(self, value):
        if _DEFERRED[0] is not None:
            _flush_deferred()
        self._grad_host = None
        if value is None:
            self._grad = None
            self._grad_zero = False
        elif isinstance(value, be.DArray):
            self._grad, self._grad_zero = value, False
        else:
            self._grad = be.from_numpy(np.asarray(value), dtype=self._data.dtype)
            self._grad_zero = False

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def ndim(self):
        return len(self._data.shape)

    def __repr__(self):
        return "Tensor(shape=%s, requires_grad=%s)" % (self.shape, self.requires_grad)

    def __array__(self, dtype=None, copy=None):
        # lets numpy consume a Tensor with one D2H copy (np.argmax(pred, axis=1) in run.py:89)
        v = self.values
        return v if dtype is None else v.astype(dtype)

    def __len__(self):
        if not self._data.shape:
            raise TypeError("len() of unsized object")
        return self._data.shape[0]

    # ------------------------------------------------------------------ comparisons (raw bool arrays)
    def _compare(self, other, op):
        o = as_tensor(other, like=self)
        return be.to_numpy(be.ew(op, self._data, o._data)).astype(bool)

    def __gt__(self, other):
        return self._compare(other, be.GT)

    def __lt__(self, other):
        return self._compare(other, be.LT)

    def __ge__(self, other):
        return self._compare(other, be.GE)

    def __le__(self, other):
        return self._compare(other, be.LE)

    # ------------------------------------------------------------------ differentiable operators
    def __add__(self, other):
        return ops.add_(self, as_tensor(other, like=self))

    def __radd__(self, other):
        return ops.add_(as_tensor(other, like=self), self)

    def __sub__(self, other):
        return ops.sub_(self, as_tensor(other, like=self))

    def __rsub__(self, other):
        return ops.sub_(as_tensor(other, like=self), self)

    def __mul__(self, other):
        return ops.mul_(self, as_tensor(other, like=self))

    def __rmul__(self, other):
        return ops.mul_(as_tensor(other, like=self), self)

    def __truediv__(self, other):
        return ops.div_(self, as_tensor(other, like=self))

    def __rtruediv__(self, other):
        return ops.div_(as_tensor(other, like=self), self)

    def __pow__(self, other):
        return ops.pow_(self, as_tensor(other, like=self))

    def __rpow__(self, other):
        return ops.pow_(as_tensor(other, like=self), self)

    def __matmul__(self, other):
        return ops.dot_(self, as_tensor(other, like=self))

    def __rmatmul__(self, other):
        return ops.dot_(as_tensor(other, like=self), self)

    def __neg__(self):
        return ops.neg_(self)

    def __getitem__(self, key):
        return ops.getitem_(self, key)

    # ------------------------------------------------------------------ in-place: rebind, not recorded
    def _rebind(self, data):
        if _DEFERRED[0] is not None:
            _flush_deferred()
        self._data = data
        self._host = None
        self._drop_grad()
        return self

    def __iadd__(self, other):
        return self._rebind(be.ew(be.ADD, self._data, as_tensor(other, like=self)._data))

    def __isub__(self, other):
        return self._rebind(be.ew(be.SUB, self._data, as_tensor(other, like=self)._data))

    def __imul__(self, other):
        return self._rebind(be.ew(be.MUL, self._data, as_tensor(other, like=self)._data))

    def __itruediv__(self, other):
        return self._rebind(be.ew(be.DIV, self._data, as_tensor(other, like=self)._data))

    def __ipow__(self, other):
        return self._rebind(be.ew(be.POW, self._data, as_tensor(other, like=self)._data))

    def __imatmul__(self, other):
        return self._rebind(be.matmul(self._data, as_tensor(other, like=self)._data))

    # ------------------------------------------------------------------ methods
    def sum(self, axis=None):
        return ops.sum_(self, axis=axis)

    def max(self, axis=None):
        return ops.max_(self, axis=axis)

    def min(self, axis=None):
        return ops.min_(self, axis=axis)

    def transpose(self, axes=None):
        return ops.transpose_(self, axes=axes)

    def log(self):
        return ops.log_(self)

    def reshape(self, newshape):
        return ops.reshape_(self, newshape)

    def flatten(self):
        return ops.flatten_(self)

    def clip(self, min=None, max=None):
        return ops.clip_(self, min, max)

    @property
    def T(self):
        return ops.transpose_(self, axes=None)

    # ------------------------------------------------------------------ autograd
    def zero_grad(self):
        """tensor.py:170-171.  Lazy: nothing is allocated until a gradient arrives."""
        if _DEFERRED[0] is not None:
            _flush_deferred()
        self._grad_host = None
        if self._gslot is not None:
            # the slot is not cleared here: the first gradient overwrites it, and Model.step clears
            # the slots that received none (while _grad_zero is set the slot's content is undefined)
            self._grad = self._gslot
            self._grad_zero = True
        else:
            self._grad = None
            self._grad_zero = True

    def _coerce_grad(self, g):
        """incoming gradient -> DArray of this tensor's dtype and shape"""
        if isinstance(g, Tensor):
            g = g._data
        elif not isinstance(g, be.DArray):
            g = be.from_numpy(np.asarray(g), dtype=self._data.dtype)
        if g.dtype != self._data.dtype:
            g = be.astype(g, self._data.dtype)
        if g.shape != self._data.shape:
            g = be.broadcast_to(g, self._data.shape)  # ValueError if not broadcastable, like +=
        return g

    def _accumulate(self, g):
        """self.grad += g (tensor.py:163)"""
        if self._grad is None and not self._grad_zero:
            # the reference fails the same way: `None += ndarray` (in-place op without zero_grad)
            raise TypeError("unsupported operand type(s) for +=: 'NoneType' and 'ndarray' "
                            "(call zero_grad() after an in-place update)")
        self._grad_host = None
        if self._gslot is not None and self._grad is self._gslot:
            if self._grad_zero:
                be.copy_into(self._gslot, g)
            else:
                be.add_inplace(self._gslot, g)
        elif self._grad_zero or self._grad is None:
            self._grad = g
        else:
            self._grad = be.ew(be.ADD, self._grad, g)
        self._grad_zero = False

    @staticmethod
    def _run_fused_bwd(node, g, pending):
        """let a node produce all its input gradients at once; False = use the per-input functions"""
        direct = {}
        for dep in node.dependency:
            t = dep["tensor"]
            if (t._gslot is not None and t._grad is t._gslot and not t.dependency
                    and id(t) not in pending):
                direct[id(t)] = (t._gslot, not t._grad_zero)
        results = node._fused_bwd(g, direct)
        if results is None:
            return False
        for dep in node.dependency:
            t = dep["tensor"]
            r = results[id(t)]
            if r is ops.WRITTEN_IN_PLACE:
                t._grad_zero = False
                t._grad_host = None
                continue
            gd = t._coerce_grad(r)
            key = id(t)
            if key in pending:
                pending[key] = be.ew(be.ADD, pending[key], gd)
            else:
                pending[key] = gd
        return True

    def backward(self, grad=None):
        assert self.requires_grad, "Call backward() on a non-requires-grad tensor."
        if _DEFERRED[0] is not None:
            _flush_deferred()
        if grad is None:
            if self._data.shape == ():
                seed = be.ones_scalar(self._data.dtype)      # shared constant, no launch
            else:
                seed = be.full(self._data.shape, 1.0, self._data.dtype)
        else:
            seed = self._coerce_grad(grad)

        # reverse topological order of the sub-graph that requires grad
        topo, seen = [], set()
        stack = [(self, False)]
        while stack:
            node, expanded = stack.pop()
            if expanded:
                topo.append(node)
                continue
            if id(node) in seen:
                continue
            seen.add(id(node))
            stack.append((node, True))
            for dep in node.dependency:
                if id(dep["tensor"]) not in seen:
                    stack.append((dep["tensor"], False))

        pending = {id(self): seed}
        for node in reversed(topo):
            g = pending.pop(id(node), None)
            if g is None:
                continue
            node._accumulate(g)
            if node._fused_bwd is not None and self._run_fused_bwd(node, g, pending):
                continue
            for dep in node.dependency:
                t = dep["tensor"]
                fn = dep["grad_fn"]
                # leaf with a slot in the flat gradient arena: let the kernel write/accumulate
                # straight into it (no temporary, no extra pass)
                if (t._gslot is not None and t._grad is t._gslot and not t.dependency
                        and getattr(fn, "supports_out", False) and id(t) not in pending):
                    fn(g, out=t._gslot, accumulate=not t._grad_zero)
                    t._grad_zero = False
                    t._grad_host = None
                    continue
                gd = t._coerce_grad(fn(g))
                key = id(t)
                if key in pending:
                    pending[key] = be.ew(be.ADD, pending[key], gd)
                else:
                    pending[key] = gd
