"""Shared op-parity case table.

One table drives three consumers:
  * oracle/make_golden.py   runs every case through the REAL reference (imported from
                            /root/reference in the build container) and stores the results in
                            tests/golden/ops.npz
  * tests/test_oracle_golden.py  runs them through oracle/ref_numpy.py and compares to the golden
  * tests/test_gpu_ops.py        runs them through the CUDA engine and compares to oracle + golden

A case is (name, op, input shapes, dtype, kwargs).  Inputs are drawn from a RandomState seeded
from the case name, so every consumer sees identical arrays without shipping them.
"""
import zlib

import numpy as np

BCAST_PATTERNS = [
    ("vec", [(7,), (7,)]),
    ("mat", [(6, 5), (6, 5)]),
    ("row_nd", [(6, 5), (5,)]),
    ("row_2d", [(6, 5), (1, 5)]),
    ("scalar", [(6, 5), ()]),
    ("outer", [(6, 1), (1, 5)]),
    ("nd3", [(2, 3, 4), (3, 1)]),
    ("lead", [(3, 4), (2, 3, 4)]),
]

CASES = []


def _add(name, op, shapes, dtype, **kw):
    CASES.append(dict(name=name, op=op, shapes=shapes, dtype=dtype, kw=kw))


for _dt in ("float32", "float64"):
    for _op in ("add", "sub", "mul", "div", "pow", "maximum", "minimum"):
        for _pn, _shapes in BCAST_PATTERNS:
            _add("%s_%s_%s" % (_op, _pn, _dt), _op, _shapes, _dt)
    for _op in ("exp", "log", "neg"):
        _add("%s_%s" % (_op, _dt), _op, [(5, 9)], _dt)
    _add("clip_lo_%s" % _dt, "clip", [(8, 6)], _dt, lo=0.0, hi=None)
    _add("clip_both_%s" % _dt, "clip", [(8, 6)], _dt, lo=-0.5, hi=0.25)
    _add("clip_hi_%s" % _dt, "clip", [(8, 6)], _dt, lo=None, hi=0.1)
    for _ax in (None, 0, 1):
        _add("sum_ax%s_%s" % (_ax, _dt), "sum", [(7, 5)], _dt, axis=_ax)
    _add("sum_ax1_3d_%s" % _dt, "sum", [(3, 4, 5)], _dt, axis=1)
    for _ax in (None, 0):
        _add("max_ax%s_%s" % (_ax, _dt), "max", [(6, 4)], _dt, axis=_ax)
        _add("min_ax%s_%s" % (_ax, _dt), "min", [(6, 4)], _dt, axis=_ax)
    _add("transpose_none_%s" % _dt, "transpose", [(2, 4, 6)], _dt, axes=None)
    _add("transpose_201_%s" % _dt, "transpose", [(2, 4, 6)], _dt, axes=(2, 0, 1))
    _add("transpose_2d_%s" % _dt, "transpose", [(40, 70)], _dt, axes=None)
    _add("reshape_%s" % _dt, "reshape", [(4, 6)], _dt, newshape=(3, -1))
    _add("flatten_%s" % _dt, "flatten", [(4, 6)], _dt)
    _add("pad_%s" % _dt, "pad", [(3, 4)], _dt, pad_width=[(1, 0), (2, 1)])
    _add("getitem_slice_%s" % _dt, "getitem", [(9, 4)], _dt, key=("slice", 2, 7))
    _add("getitem_rows_%s" % _dt, "getitem", [(9, 4)], _dt, key=("rows", [8, 0, 3, 3, 5]))
    _add("getitem_int_%s" % _dt, "getitem", [(9, 4)], _dt, key=("int", 4))
    _add("getitem_2d_%s" % _dt, "getitem", [(9, 4)], _dt, key=("tuple2", 1, 6, 2))
    for _m, _k, _n in ((4, 3, 5), (33, 70, 30), (128, 100, 70), (65, 129, 31)):
        _add("matmul_%dx%dx%d_%s" % (_m, _k, _n, _dt), "matmul", [(_m, _k), (_k, _n)], _dt)
    for _b, _c in ((4, 3), (128, 10), (33, 50)):
        _add("ce_%dx%d_%s" % (_b, _c, _dt), "ce", [(_b, _c)], _dt)
    _add("dense_relu_%s" % _dt, "dense_relu", [(16, 12), (12, 10), (1, 10)], _dt)


def case_by_name(name):
    for c in CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


def make_inputs(case):
    """(inputs, upstream-gradient seed generator)"""
    rng = np.random.RandomState(zlib.crc32(case["name"].encode()) & 0x7FFFFFFF)
    dt = np.dtype(case["dtype"])
    arrs = [rng.standard_normal(s).astype(dt) for s in case["shapes"]]
    op = case["op"]
    if op == "log":
        arrs[0] = (np.abs(arrs[0]) + 0.5).astype(dt)
    elif op == "pow":
        arrs[0] = (np.abs(arrs[0]) + 0.5).astype(dt)     # positive base: ln(a) is finite
    elif op == "div":
        arrs[1] = np.asarray(np.where(arrs[1] >= 0, arrs[1] + 0.5, arrs[1] - 0.5), dtype=dt)
    elif op == "ce":
        labels = rng.randint(0, case["shapes"][0][1], case["shapes"][0][0])
        arrs.append(np.eye(case["shapes"][0][1])[labels].astype(dt))
    return arrs, rng


def key_of(case):
    k = case["kw"]["key"]
    if k[0] == "slice":
        return slice(k[1], k[2])
    if k[0] == "rows":
        return np.array(k[1])
    if k[0] == "int":
        return k[1]
    if k[0] == "tuple2":
        return (slice(k[1], k[2]), k[3])
    raise ValueError(k)


def run_reference_style(case, Tensor, ops, ce_loss):
    """Evaluate a case with an API shaped like the reference's (core.tensor.Tensor + core.ops):
    works for the real reference and for the CUDA engine alike.
    Returns (out_values, [input_grads]) as numpy arrays."""
    arrs, rng = make_inputs(case)
    op, kw = case["op"], case["kw"]
    n_diff = len(case["shapes"])
    ts = [Tensor(a, requires_grad=(i < n_diff)) for i, a in enumerate(arrs)]
    if op == "add": out = ts[0] + ts[1]
    elif op == "sub": out = ts[0] - ts[1]
    elif op == "mul": out = ts[0] * ts[1]
    elif op == "div": out = ts[0] / ts[1]
    elif op == "pow": out = ts[0] ** ts[1]
    elif op == "maximum": out = ops.maximum_(ts[0], ts[1])
    elif op == "minimum": out = ops.minimum_(ts[0], ts[1])
    elif op == "exp": out = ops.exp(ts[0])
    elif op == "log": out = ops.log(ts[0])
    elif op == "neg": out = -ts[0]
    elif op == "clip": out = ops.clip(ts[0], kw["lo"], kw["hi"])
    elif op == "sum": out = ts[0].sum(axis=kw["axis"])
    elif op == "max": out = ts[0].max(axis=kw["axis"])
    elif op == "min": out = ts[0].min(axis=kw["axis"])
    elif op == "transpose": out = ts[0].transpose(kw["axes"])
    elif op == "reshape": out = ops.reshape(ts[0], kw["newshape"])
    elif op == "flatten": out = ops.flatten(ts[0])
    elif op == "pad": out = ops.pad(ts[0], kw["pad_width"])
    elif op == "getitem": out = ts[0][key_of(case)]
    elif op == "matmul": out = ts[0] @ ts[1]
    elif op == "ce": out = ce_loss(ts[0], ts[1])
    elif op == "dense_relu": out = ops.clip(ts[0] @ ts[1] + ts[2], 0.0)
    else: raise ValueError(op)
    out_vals = np.array(out.values)
    g = rng.standard_normal(out_vals.shape).astype(out_vals.dtype)
    out.backward(g)
    return out_vals, [np.array(t.grad) for t in ts[:n_diff]]


def run_oracle(case, R):
    """Evaluate a case with oracle/ref_numpy.py (module passed as R)."""
    arrs, rng = make_inputs(case)
    op, kw = case["op"], case["kw"]
    n_diff = len(case["shapes"])
    ts = [R.RefTensor(a, requires_grad=(i < n_diff)) for i, a in enumerate(arrs)]
    if op == "add": out = ts[0] + ts[1]
    elif op == "sub": out = ts[0] - ts[1]
    elif op == "mul": out = ts[0] * ts[1]
    elif op == "div": out = ts[0] / ts[1]
    elif op == "pow": out = ts[0] ** ts[1]
    elif op == "maximum": out = R.maximum(ts[0], ts[1])
    elif op == "minimum": out = R.minimum(ts[0], ts[1])
    elif op == "exp": out = R.exp(ts[0])
    elif op == "log": out = R.log(ts[0])
    elif op == "neg": out = -ts[0]
    elif op == "clip": out = R.clip(ts[0], kw["lo"], kw["hi"])
    elif op == "sum": out = ts[0].sum(axis=kw["axis"])
    elif op == "max": out = ts[0].max(axis=kw["axis"])
    elif op == "min": out = ts[0].min(axis=kw["axis"])
    elif op == "transpose": out = R.transpose(ts[0], kw["axes"])
    elif op == "reshape": out = R.reshape(ts[0], kw["newshape"])
    elif op == "flatten": out = R.flatten(ts[0])
    elif op == "pad": out = R.pad(ts[0], kw["pad_width"])
    elif op == "getitem": out = ts[0][key_of(case)]
    elif op == "matmul": out = ts[0] @ ts[1]
    elif op == "ce": out = R.softmax_cross_entropy(ts[0], ts[1])
    elif op == "dense_relu": out = R.clip(ts[0] @ ts[1] + ts[2], 0.0)
    else: raise ValueError(op)
    out_vals = np.array(out.values)
    g = rng.standard_normal(out_vals.shape).astype(out_vals.dtype)
    out.backward(g)
    return out_vals, [np.array(t.grad) for t in ts[:n_diff]]


def rel_err(a, b):
    """max|a-b| / max(max|b|, tiny) -- the per-op metric of BASELINE.md config 2"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("inf")
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
