"""NEP-13 / NEP-18 on ``DeviceChunk``: the chunk-type level of the plug-in boundary.

With these protocols the reference's own NumPy-shaped chunk functions -- ``mean_chunk``,
``moment_chunk``, ``arg_chunk``, ``mean_agg`` ... (``reductions/_common.py``), ``np.transpose``
(``manipulation/_transpose.py:17``), ``_concatenate2`` (``_core_utils.py:191``) -- run unmodified on
device blocks: every ufunc call becomes a one-operator fused program (same JIT cache, same
kernels), every reduction a reduction launch.  This is the *correct but unfused* way in; the fast
path is the expression-level executor.

Mixed host/device operands are part of the contract (``numel`` returns NumPy arrays built from
``x.shape``, ``_dispatch.py:209-238``): small host arrays are uploaded on the fly.
"""
from __future__ import annotations

import math
from numbers import Number

import numpy as np

from . import _codegen as cg
from . import _lib
from . import _runtime as rt
from ._device import DeviceChunk, current_device

HANDLED = {}


def implements(*np_funcs):
    def deco(f):
        for nf in np_funcs:
            HANDLED[nf] = f
        return f
    return deco


def _as_chunk(x):
    if isinstance(x, DeviceChunk):
        return x
    return DeviceChunk.from_numpy(np.asarray(x), current_device())


# ----------------------------------------------------------------------------- ufuncs
def array_ufunc(self, ufunc, method, *inputs, **kwargs):
    if method != "__call__":
        return NotImplemented
    out = kwargs.pop("out", None)
    dtype = kwargs.pop("dtype", None)
    kwargs.pop("casting", None)
    if out is not None or kwargs.pop("where", True) is not True or kwargs:
        return NotImplemented
    prog = cg.Program()
    refs, chunks = [], []
    for x in inputs:
        if isinstance(x, (Number, np.generic)) or (isinstance(x, np.ndarray) and x.ndim == 0):
            refs.append(prog.const(x[()] if isinstance(x, np.ndarray) else x))
        else:
            c = _as_chunk(x)
            refs.append(prog.add_input(c.dtype))
            chunks.append(c)
    res = prog.op(ufunc.__name__, *refs)
    if dtype is not None and np.dtype(dtype) != res.dtype:
        res = prog.op("astype", res, dtype=np.dtype(dtype).name)
    prog.set_output(res)
    shape = np.broadcast_shapes(*[c.shape for c in chunks])
    outc = DeviceChunk.empty(shape, prog.out_dtype, chunks[0].device)
    if outc.size:
        ins = [(c.ptr, c.broadcast_to(shape).strides) for c in chunks]
        for L in rt.fused_launches(prog, _lib.RED_NONE, (), [rt.BlockArgs(shape=shape, inputs=ins, out0=outc.ptr)]):
            L.run()
            outc._keep = L
    return outc


def _unary(name):
    return lambda self: array_ufunc(self, getattr(np, name), "__call__", self)


def _binary(name, reverse=False):
    def f(self, other):
        if not isinstance(other, (DeviceChunk, Number, np.generic, np.ndarray)):
            return NotImplemented
        args = (other, self) if reverse else (self, other)
        return array_ufunc(self, getattr(np, name), "__call__", *args)
    return f


# ----------------------------------------------------------------------------- reductions
def _reduce(x, redop, axis, keepdims, out_dtype, acc_dtype=None, want_arg=False):
    x = _as_chunk(x)
    nd = x.ndim
    axes = tuple(range(nd)) if axis is None else tuple(a % nd for a in ((axis,) if np.isscalar(axis) else axis))
    kd_shape = tuple(1 if d in axes else n for d, n in enumerate(x.shape))
    prog = cg.Program()
    prog.set_output(prog.op("positive", prog.add_input(x.dtype)))
    out = DeviceChunk.empty(kd_shape, out_dtype if not want_arg else x.dtype, x.device)
    arg = DeviceChunk.empty(kd_shape, np.int64, x.device) if want_arg else None
    kw = {}
    if want_arg and len(axes) == nd and nd > 1:
        kw["arg_ravel"] = (x.shape, (0,) * nd, x.shape)
    blk = rt.BlockArgs(shape=x.shape, inputs=[(x.ptr, x.strides)], out0=out.ptr, out1=arg.ptr if want_arg else 0, **kw)
    for L in rt.fused_launches(prog, redop, axes, [blk], acc_dtype=acc_dtype or out_dtype):
        L.run()
        out._keep = L
    res = arg if want_arg else out
    if not keepdims:
        res = res.reshape(tuple(n for d, n in enumerate(x.shape) if d not in axes))
    return res


@implements(np.sum)
def _sum(a, axis=None, dtype=None, keepdims=False, **kw):
    dt = np.dtype(dtype) if dtype is not None else np.zeros(1, dtype=a.dtype).sum().dtype
    return _reduce(a, _lib.RED_SUM, axis, keepdims, dt)


@implements(np.prod)
def _prod(a, axis=None, dtype=None, keepdims=False, **kw):
    dt = np.dtype(dtype) if dtype is not None else np.zeros(1, dtype=a.dtype).prod().dtype
    return _reduce(a, _lib.RED_PROD, axis, keepdims, dt)


@implements(np.max, np.amax)
def _max(a, axis=None, keepdims=False, **kw):
    return _reduce(a, _lib.RED_MAX, axis, keepdims, a.dtype)


@implements(np.min, np.amin)
def _min(a, axis=None, keepdims=False, **kw):
    return _reduce(a, _lib.RED_MIN, axis, keepdims, a.dtype)


@implements(np.any)
def _any(a, axis=None, keepdims=False, **kw):
    return _reduce(a, _lib.RED_ANY, axis, keepdims, np.bool_)


@implements(np.all)
def _all(a, axis=None, keepdims=False, **kw):
    return _reduce(a, _lib.RED_ALL, axis, keepdims, np.bool_)


@implements(np.argmax)
def _argmax(a, axis=None, keepdims=False, **kw):
    r = _reduce(a, _lib.RED_ARGMAX, axis, True, np.int64, want_arg=True)
    return r if keepdims else r.reshape(tuple(n for d, n in enumerate(a.shape) if axis is not None and d != axis % a.ndim))


@implements(np.argmin)
def _argmin(a, axis=None, keepdims=False, **kw):
    r = _reduce(a, _lib.RED_ARGMIN, axis, True, np.int64, want_arg=True)
    return r if keepdims else r.reshape(tuple(n for d, n in enumerate(a.shape) if axis is not None and d != axis % a.ndim))


# ----------------------------------------------------------------------------- windows
@implements(np.lib.stride_tricks.sliding_window_view)
def _sliding_window_view(x, window_shape, axis=None, **kw):
    """Zero-copy: the window axes re-use the strides of the axes they slide along (what NumPy's
    ``as_strided`` does on the host); consumers read the overlapping windows through the descriptors."""
    x = _as_chunk(x)
    window_shape = tuple(window_shape) if np.iterable(window_shape) else (window_shape,)
    if axis is None:
        axis = tuple(range(x.ndim))
    axis = tuple(a % x.ndim for a in ((axis,) if np.isscalar(axis) else axis))
    if len(axis) != len(window_shape):
        raise ValueError("Must provide matching length window_shape and axis")
    shape = list(x.shape)
    for ax, w in zip(axis, window_shape):
        if w <= 0 or shape[ax] < w:
            raise ValueError("window shape cannot be larger than input array shape")
        shape[ax] -= w - 1
    return DeviceChunk(x.buf, tuple(shape) + window_shape, x.dtype,
                       strides=tuple(x.strides) + tuple(x.strides[ax] for ax in axis), offset=x.offset)


# ----------------------------------------------------------------------------- cumulative scans
def _cumulative(a, redop, axis, dtype, nan):
    """``np.cumsum`` / ``np.cumprod`` (+ nan variants) of ONE chunk -- what ``CumReduction._layer``
    (``reductions/_cumulative.py:211-222``) calls per block -- as a single scan launch."""
    x = _as_chunk(a)
    if axis is None:
        if x.ndim != 1:
            x = copy(x).reshape((x.size,))
        axis = 0
    fn = np.cumsum if redop == _lib.RED_SUM else np.cumprod
    acc = np.dtype(dtype) if dtype is not None else fn(np.ones((0,), dtype=x.dtype)).dtype
    prog = cg.Program()
    ref = prog.add_input(x.dtype)
    if nan and x.dtype.kind == "f":
        ref = prog.op("where", prog.op("isnan", ref), prog.typed_const(0 if redop == _lib.RED_SUM else 1, x.dtype), ref)
    prog.set_output(prog.op("astype", ref, dtype=acc))
    out = DeviceChunk.empty(x.shape, acc, x.device)
    if x.size:
        blk = rt.BlockArgs(shape=x.shape, inputs=[(x.ptr, x.strides)], out0=out.ptr)
        for L in rt.scan_launches(prog, redop, axis % x.ndim, [blk], acc):
            L.run()
            out._keep = L
    return out


@implements(np.cumsum)
def _cumsum(a, axis=None, dtype=None, **kw):
    return _cumulative(a, _lib.RED_SUM, axis, dtype, False)


@implements(np.cumprod)
def _cumprod(a, axis=None, dtype=None, **kw):
    return _cumulative(a, _lib.RED_PROD, axis, dtype, False)


@implements(np.nancumsum)
def _nancumsum(a, axis=None, dtype=None, **kw):
    return _cumulative(a, _lib.RED_SUM, axis, dtype, True)


@implements(np.nancumprod)
def _nancumprod(a, axis=None, dtype=None, **kw):
    return _cumulative(a, _lib.RED_PROD, axis, dtype, True)


# ----------------------------------------------------------------------------- views / movement
@implements(np.transpose)
def _transpose(a, axes=None):
    return a.transpose(axes) if axes is not None else a.transpose()


@implements(np.broadcast_to)
def _broadcast_to(a, shape, **kw):
    return a.broadcast_to(shape)


@implements(np.expand_dims)
def _expand_dims(a, axis):
    axis = axis if axis >= 0 else axis + a.ndim + 1
    idx = tuple(slice(None) for _ in range(axis)) + (None,)
    return a[idx]


@implements(np.squeeze)
def _squeeze(a, axis=None):
    axes = [d for d, n in enumerate(a.shape) if n == 1] if axis is None else [ax % a.ndim for ax in np.atleast_1d(axis)]
    return a[tuple(0 if d in axes else slice(None) for d in range(a.ndim))]


def copy(a: DeviceChunk) -> DeviceChunk:
    """Contiguous copy (``getitem``'s copy of small selections, ``_chunk.py:285-317``): one launch of
    the identity chain, any strides (negative steps included)."""
    a = _as_chunk(a)
    out = DeviceChunk.empty(a.shape, a.dtype, a.device)
    if a.size:
        prog = cg.Program()
        prog.set_output(prog.op("astype", prog.add_input(a.dtype), dtype=a.dtype))
        blk = rt.BlockArgs(shape=a.shape, inputs=[(a.ptr, a.strides)], out0=out.ptr)
        for L in rt.fused_launches(prog, _lib.RED_NONE, (), [blk]):
            L.run()
            out._keep = L
    return out


@implements(np.concatenate)
def concatenate(arrays, axis=0, **kw):
    """``concatenate_lookup`` implementation (``_core_utils.py:241``): one tiled gather."""
    from ._executor import _copy_descs

    arrays = [_as_chunk(a) for a in arrays]
    nd = arrays[0].ndim
    axis %= nd
    dt = np.result_type(*[a.dtype for a in arrays])
    arrays = [a if a.dtype == dt else a.astype(dt) for a in arrays]
    shape = list(arrays[0].shape)
    shape[axis] = sum(a.shape[axis] for a in arrays)
    out = DeviceChunk.empty(shape, dt, arrays[0].device)
    copies, pos = [], 0
    for a in arrays:
        if not a.size:
            continue
        src = a if a.strides[-1] == 1 or a.shape[-1] == 1 else copy(a)
        dst = out[tuple(slice(pos, pos + a.shape[axis]) if d == axis else slice(None) for d in range(nd))]
        copies.extend(_copy_descs(src, dst, dt.itemsize))
        pos += a.shape[axis]
    g = rt.GatherLaunch(copies)
    g.run()
    out._keep = (g, arrays)
    return out


@implements(np.stack)
def _stack(arrays, axis=0, **kw):
    arrays = [_expand_dims(_as_chunk(a), axis) for a in arrays]
    return concatenate(arrays, axis=axis if axis >= 0 else axis + arrays[0].ndim)


@implements(np.where)
def _where(cond, x, y):
    prog = cg.Program()
    refs, chunks = [], []
    for v in (cond, x, y):
        if isinstance(v, (Number, np.generic)):
            refs.append(prog.const(v))
        else:
            c = _as_chunk(v)
            refs.append(prog.add_input(c.dtype))
            chunks.append(c)
    prog.set_output(prog.op("where", *refs))
    shape = np.broadcast_shapes(*[c.shape for c in chunks])
    out = DeviceChunk.empty(shape, prog.out_dtype, chunks[0].device)
    ins = [(c.ptr, c.broadcast_to(shape).strides) for c in chunks]
    for L in rt.fused_launches(prog, _lib.RED_NONE, (), [rt.BlockArgs(shape=shape, inputs=ins, out0=out.ptr)]):
        L.run()
        out._keep = L
    return out


def _like(a, dtype=None, shape=None, fill=None):
    dt = np.dtype(dtype) if dtype is not None else a.dtype
    if dt.names is not None:
        # selects the dict path of arg_chunk / arg_combine (reductions/_common.py:724-728)
        raise TypeError("DeviceChunk does not support structured dtypes")
    out = DeviceChunk.empty(a.shape if shape is None else ((shape,) if np.isscalar(shape) else tuple(shape)), dt, a.device)
    if fill is not None and out.size:
        rt.fill(out, fill)
    return out


@implements(np.empty_like)
def _empty_like(a, dtype=None, shape=None, **kw):
    return _like(a, dtype, shape)


@implements(np.zeros_like)
def _zeros_like(a, dtype=None, shape=None, **kw):
    return _like(a, dtype, shape, 0)


@implements(np.ones_like)
def _ones_like(a, dtype=None, shape=None, **kw):
    return _like(a, dtype, shape, 1)


@implements(np.full_like)
def _full_like(a, fill_value, dtype=None, shape=None, **kw):
    return _like(a, dtype, shape, fill_value)


def array_function(self, func, types, args, kwargs):
    f = HANDLED.get(func)
    if f is None:
        raise TypeError(f"np.{getattr(func, '__name__', func)} has no B200 implementation for DeviceChunk "
                        "(there is no host fallback)")
    return f(*args, **kwargs)


# dispatch-table implementations
def divide(a, b, dtype=None):
    """``divide_lookup`` (``_dispatch.py:157-162``)."""
    return array_ufunc(_as_chunk(a), np.true_divide, "__call__", a, b, dtype=dtype)


def numel(x, **kwargs):
    """``numel_lookup`` default (``_dispatch.py:209-238``): shape arithmetic only, host result."""
    shape, axis = x.shape, kwargs.get("axis")
    keepdims, dtype = kwargs.get("keepdims", False), kwargs.get("dtype", np.float64)
    if axis is None:
        prod = np.prod(shape, dtype=dtype)
        return np.full((1,) * len(shape), prod, dtype=dtype) if keepdims else prod
    axis = [axis] if not isinstance(axis, (tuple, list)) else axis
    prod = math.prod(shape[d] for d in axis)
    new = tuple(1 if d in axis else n for d, n in enumerate(shape)) if keepdims else \
        tuple(n for d, n in enumerate(shape) if d not in axis)
    return np.broadcast_to(np.array(prod, dtype=dtype), new)


def tensordot(a, b, axes=2):
    raise NotImplementedError("tensordot on DeviceChunk: use the BlockGEMM expression (dask_array_b200._matmul)")


# ----------------------------------------------------------------------------- install on DeviceChunk
def _install():
    D = DeviceChunk
    D.__array_ufunc__ = array_ufunc
    D.__array_function__ = array_function
    for op, name in [("add", "add"), ("sub", "subtract"), ("mul", "multiply"), ("truediv", "true_divide"),
                     ("floordiv", "floor_divide"), ("mod", "remainder"), ("pow", "power"), ("and", "bitwise_and"),
                     ("or", "bitwise_or"), ("xor", "bitwise_xor"), ("lshift", "left_shift"), ("rshift", "right_shift"),
                     ("lt", "less"), ("le", "less_equal"), ("gt", "greater"), ("ge", "greater_equal"),
                     ("eq", "equal"), ("ne", "not_equal")]:
        setattr(D, f"__{op}__", _binary(name))
        if op not in ("lt", "le", "gt", "ge", "eq", "ne"):
            setattr(D, f"__r{op}__", _binary(name, reverse=True))
    D.__neg__, D.__pos__, D.__abs__, D.__invert__ = _unary("negative"), _unary("positive"), _unary("absolute"), _unary("invert")
    D.__hash__ = None
    D.sum = lambda self, axis=None, dtype=None, keepdims=False, **kw: _sum(self, axis, dtype, keepdims)
    D.max = lambda self, axis=None, keepdims=False, **kw: _max(self, axis, keepdims)
    D.min = lambda self, axis=None, keepdims=False, **kw: _min(self, axis, keepdims)
    D.copy = copy
    D.ravel = lambda self: (self if self.is_contiguous else copy(self)).reshape(-1)
    D.squeeze = lambda self, axis=None: _squeeze(self, axis)

    def astype(self, dtype, copy=True, **kw):
        dtype = np.dtype(dtype)
        if dtype == self.dtype and not copy:
            return self
        prog = cg.Program()
        prog.set_output(prog.op("astype", prog.add_input(self.dtype), dtype=dtype.name))
        out = DeviceChunk.empty(self.shape, dtype, self.device)
        if out.size:
            blk = rt.BlockArgs(shape=self.shape, inputs=[(self.ptr, self.strides)], out0=out.ptr)
            for L in rt.fused_launches(prog, _lib.RED_NONE, (), [blk]):
                L.run()
                out._keep = L
        return out

    def setitem(self, index, value):
        """``arg[:] = value`` / ``x[idx] = chunk`` (``arg_chunk`` ``_common.py:713``): a strided copy."""
        from ._executor import _copy_descs

        dst = self[index]
        if isinstance(value, (Number, np.generic)) or (isinstance(value, np.ndarray) and value.size == 1):
            src = _as_chunk(np.full(dst.shape, np.asarray(value).reshape(-1)[0], dtype=self.dtype))
        else:
            src = _as_chunk(value)
            if src.dtype != self.dtype:
                src = src.astype(self.dtype)
            src = src.broadcast_to(dst.shape)
            if not src.is_contiguous:
                src = copy(src)
        g = rt.GatherLaunch(_copy_descs(src, dst, self.itemsize))
        g.run()
        self._keep_set = (g, src)

    D.astype = astype
    D.__setitem__ = setitem


_install()
