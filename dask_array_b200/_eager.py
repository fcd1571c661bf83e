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
    if ufunc is np.matmul:            # a generalised ufunc, not element-wise: the contraction path
        if kwargs.get("out") is not None:
            return NotImplemented
        return _np_matmul(*inputs)
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


# ----------------------------------------------------------------------------- integer-array gathers
def _index_chunk(ix):
    c = _as_chunk(ix)
    if c.dtype.kind == "b":
        raise NotImplementedError("boolean-mask indexing of a DeviceChunk (data-dependent shape; outside the hot path)")
    if c.dtype.kind not in "iu":
        raise IndexError("arrays used as indices must be of integer type")
    if c.dtype != np.int64:
        c = c.astype(np.int64)
    return c if c.is_contiguous else copy(c)


def _take(src: DeviceChunk, idx: DeviceChunk, n: int, inner: int, out_shape) -> DeviceChunk:
    out = DeviceChunk.empty(out_shape, src.dtype, src.device)
    if out.size:
        _lib.check(_lib.lib.b2_take(src.itemsize, src.ptr, idx.ptr, out.ptr, out.size, n, inner, rt.current_stream_ptr()))
        out._keep = (src, idx)
    return out


def fancy_getitem(a: DeviceChunk, index):
    """The two integer-array gathers of ``_arg_combine`` (``reductions/_common.py:687-697``):
    ``vals.ravel()[local_args]`` (one index array on a 1-D chunk) and the take-along-axis spelled with
    ``np.ogrid``: ``vals[ogrid_0, .., local_args, .., ogrid_k]``.  Anything else is refused."""
    if not isinstance(index, tuple):
        index = (index,)
    if len(index) == 1 and a.ndim == 1:
        idx = _index_chunk(index[0])
        src = a if a.is_contiguous else copy(a)
        return _take(src, idx, a.shape[0], 0, idx.shape)
    if len(index) != a.ndim:
        raise NotImplementedError("DeviceChunk integer-array indexing: one index array per axis (np.ogrid form) or a 1-D take")
    dev = [k for k, ix in enumerate(index) if isinstance(ix, DeviceChunk)]
    if len(dev) != 1:
        raise NotImplementedError("DeviceChunk integer-array indexing: exactly one data-dependent index array (the arg positions)")
    axis = dev[0]
    local = index[axis]
    kept = [d for d in range(a.ndim) if d != axis]
    if local.ndim != len(kept) or any(local.shape[j] != a.shape[d] for j, d in enumerate(kept)):
        raise NotImplementedError("DeviceChunk take-along-axis: the index array must span every other axis")
    for j, d in enumerate(kept):
        g = np.asarray(index[d])
        want = np.arange(a.shape[d]).reshape(tuple(a.shape[d] if q == j else 1 for q in range(len(kept))))
        if g.shape != want.shape or not np.array_equal(g, want):
            raise NotImplementedError("DeviceChunk take-along-axis: the other indices must be the np.ogrid open grid")
    src = a if a.is_contiguous else copy(a)
    inner = math.prod(a.shape[axis + 1:])
    return _take(src, _index_chunk(local), a.shape[axis], max(inner, 1), local.shape)


@implements(np.take)
def _np_take(a, indices, axis=None, **kw):
    a = _as_chunk(a)
    if axis is None:
        a = a.ravel()
        axis = 0
    if a.ndim != 1:
        raise NotImplementedError("np.take on DeviceChunk: 1-D (or axis=None) only")
    return fancy_getitem(a, (indices,))


@implements(np.unravel_index)
def _unravel_index(indices, shape, order="C"):
    """``arg_chunk`` ravel path (``_common.py:711``): flat positions -> per-axis coordinates, on the device."""
    if order != "C":
        raise NotImplementedError("unravel_index: C order only")
    idx = _as_chunk(indices)
    out = []
    for n in reversed(tuple(shape)):
        out.append(idx % np.int64(n))
        idx = idx // np.int64(n)
    return tuple(reversed(out))


@implements(np.ravel_multi_index)
def _ravel_multi_index(multi_index, dims, mode="raise", order="C"):
    """``_common.py:713``: coordinates (device chunks and / or host integers) -> flat positions."""
    if order != "C":
        raise NotImplementedError("ravel_multi_index: C order only")
    acc = None
    for i, n in zip(multi_index, dims):
        acc = i if acc is None else acc * np.int64(n) + i
    return acc if isinstance(acc, DeviceChunk) else _as_chunk(np.asarray(acc, dtype=np.int64))


# ----------------------------------------------------------------------------- contractions on chunks
def _matricize(x: DeviceChunk, free, contracted):
    """(prod(free), prod(contracted)) row-major matrix holding ``x`` with the contracted axes last."""
    moved = x.transpose(tuple(free) + tuple(contracted))
    if not moved.is_contiguous:
        moved = copy(moved)
    m = math.prod(x.shape[d] for d in free)
    k = math.prod(x.shape[d] for d in contracted)
    return moved.reshape((m, k)), m, k


def gemm_tn(a2: DeviceChunk, b2: DeviceChunk) -> DeviceChunk:
    """``a2 @ b2.T`` for row-major (M, K) and (N, K) chunks: bf16 / fp32 on the tensor cores
    (``b2_gemm_tn_pairs``; fp32 through the bf16 x 3 split), every other NumPy number type through the
    exact SIMT kernel (``b2_gemm_tn_simt``: fp64 FMA, integer multiply-add in the result type)."""
    import ctypes as C

    (M, K), (N, K2) = a2.shape, b2.shape
    assert K == K2
    dt = np.result_type(a2.dtype, b2.dtype)
    name = dt.name
    if a2.dtype == b2.dtype and name in ("float32", "bfloat16") and K % 8 == 0 and a2.ptr % 16 == 0 and b2.ptr % 16 == 0:
        out = DeviceChunk.empty((M, N), np.float32, a2.device)
        if not out.size:
            return out
        st = rt.current_stream_ptr()
        if name == "bfloat16":
            _lib.check(_lib.lib.b2_gemm_tn(_lib.dtype_code("bfloat16"), a2.ptr, K, b2.ptr, K, out.ptr, N, M, N, K, 0, st))
            out._keep = (a2, b2)
            return out
        u16 = np.dtype("uint16")
        pa = [DeviceChunk.empty(a2.shape, u16, a2.device) for _ in range(3)]
        pb = [DeviceChunk.empty(b2.shape, u16, b2.device) for _ in range(3)]
        _lib.check(_lib.lib.b2_split3_bf16(a2.ptr, pa[0].ptr, pa[1].ptr, pa[2].ptr, a2.size, st))
        _lib.check(_lib.lib.b2_split3_bf16(b2.ptr, pb[0].ptr, pb[1].ptr, pb[2].ptr, b2.size, st))
        combos = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]
        arrA = (C.c_void_p * 6)(*[pa[i].ptr for i, _ in combos])
        arrB = (C.c_void_p * 6)(*[pb[j].ptr for _, j in combos])
        _lib.check(_lib.lib.b2_gemm_tn_pairs(_lib.dtype_code("bfloat16"), arrA, arrB, 6, K, K, out.ptr, N, M, N, K, 0, st))
        out._keep = (pa, pb)
        return out
    if name not in ("float64", "float32", "int32", "int64", "uint32", "uint64"):
        if dt.kind in "iub":
            dt = np.dtype(np.int64) if dt.kind in "ib" else np.dtype(np.uint64)
        elif dt.kind == "f":
            dt = np.dtype(np.float32)
        else:
            raise NotImplementedError(f"matmul / tensordot of dtype {dt} has no B200 kernel")
    want = np.result_type(a2.dtype, b2.dtype)
    a2 = a2 if a2.dtype == dt else a2.astype(dt)
    b2 = b2 if b2.dtype == dt else b2.astype(dt)
    out = DeviceChunk.empty((M, N), dt, a2.device)
    if out.size:
        if K == 0:
            rt.fill(out, 0)
        else:
            _lib.check(_lib.lib.b2_gemm_tn_simt(_lib.dtype_code(dt), a2.ptr, K, b2.ptr, K, out.ptr, N, M, N, K, 0,
                                                rt.current_stream_ptr()))
        out._keep = (a2, b2)
    return out if out.dtype == want else out.astype(want)


def tensordot(a, b, axes=2):
    """``tensordot_lookup`` implementation (``_core_utils.py:1252-1258``; called per block triple by
    ``_tensordot``, ``linalg/_tensordot.py:20-42``): matricise both operands and run one GEMM."""
    a, b = _as_chunk(a), _as_chunk(b)
    if isinstance(axes, (int, np.integer)):
        la, lb = tuple(range(a.ndim - axes, a.ndim)), tuple(range(axes))
    else:
        la, lb = axes
        la = (la,) if isinstance(la, (int, np.integer)) else tuple(la)
        lb = (lb,) if isinstance(lb, (int, np.integer)) else tuple(lb)
    la, lb = tuple(x % a.ndim for x in la), tuple(x % b.ndim for x in lb)
    if len(la) != len(lb) or any(a.shape[i] != b.shape[j] for i, j in zip(la, lb)):
        raise ValueError("shape-mismatch for sum")
    fa = [d for d in range(a.ndim) if d not in la]
    fb = [d for d in range(b.ndim) if d not in lb]
    a2, _, _ = _matricize(a, fa, la)
    b2, _, _ = _matricize(b, fb, lb)
    out = gemm_tn(a2, b2)
    return out.reshape(tuple(a.shape[d] for d in fa) + tuple(b.shape[d] for d in fb))


@implements(np.tensordot)
def _np_tensordot(a, b, axes=2):
    return tensordot(a, b, axes=axes)


@implements(np.matmul, np.dot)
def _np_matmul(a, b, **kw):
    a, b = _as_chunk(a), _as_chunk(b)
    if a.ndim != 2 or b.ndim != 2:
        raise NotImplementedError("np.matmul on DeviceChunk: 2-D operands")
    return tensordot(a, b, axes=((1,), (0,)))


def einsum(subscripts, *operands, dtype=None, **kwargs):
    """``einsum_lookup`` implementation (``_dispatch.py:146``; called per block by ``chunk_einsum``,
    ``_einsum.py:20-34``).  Explicit or implicit subscripts over one or more operands, each label at most
    once per operand; operands are contracted pairwise, left to right: labels that survive in neither the
    output nor a later operand are summed, labels shared by both operands AND still needed afterwards
    (batch labels) are refused -- a per-label GEMM loop is not worth a launch each."""
    if kwargs.get("out") is not None:
        raise NotImplementedError("einsum(out=...) on DeviceChunk")
    subs = subscripts.replace(" ", "")
    if "." in subs:
        raise NotImplementedError("einsum ellipsis on DeviceChunk")
    if "->" in subs:
        ins, out = subs.split("->")
    else:
        ins = subs
        flat = ins.replace(",", "")
        out = "".join(sorted(c for c in set(flat) if flat.count(c) == 1))
    terms = ins.split(",")
    if len(terms) != len(operands):
        raise ValueError("einsum: number of subscripts does not match the operands")
    ops = [_as_chunk(o) for o in operands]
    for t, o in zip(terms, ops):
        if len(set(t)) != len(t) or len(t) != o.ndim:
            raise NotImplementedError("einsum on DeviceChunk: every label at most once per operand")
    cur, lab = ops[0], terms[0]
    for k in range(1, len(ops)):
        nxt, nl = ops[k], terms[k]
        later = set(out).union(*[set(t) for t in terms[k + 1:]]) if k + 1 < len(terms) else set(out)
        # labels private to one side and not needed later: summed first
        for side in (0, 1):
            x, xl, other = (cur, lab, nl) if side == 0 else (nxt, nl, lab)
            drop = [i for i, c in enumerate(xl) if c not in other and c not in later]
            if drop:
                x = _sum(x, axis=tuple(drop))
                xl = "".join(c for i, c in enumerate(xl) if i not in drop)
            if side == 0:
                cur, lab = x, xl
            else:
                nxt, nl = x, xl
        shared = [c for c in lab if c in nl]
        if any(c in later for c in shared):
            raise NotImplementedError("einsum on DeviceChunk: batch labels (shared by two operands and kept) are not supported")
        cur = tensordot(cur, nxt, axes=([lab.index(c) for c in shared], [nl.index(c) for c in shared]))
        lab = "".join(c for c in lab if c not in shared) + "".join(c for c in nl if c not in shared)
    drop = [i for i, c in enumerate(lab) if c not in out]
    if drop:
        cur = _sum(cur, axis=tuple(drop))
        lab = "".join(c for i, c in enumerate(lab) if i not in drop)
    if sorted(lab) != sorted(out):
        raise ValueError(f"einsum: output labels {out!r} are not produced by the operands")
    res = cur.transpose(tuple(lab.index(c) for c in out)) if lab != out else cur
    if dtype is not None and np.dtype(dtype) != res.dtype:
        res = res.astype(np.dtype(dtype))
    return res


@implements(np.einsum)
def _np_einsum(subscripts, *operands, **kw):
    return einsum(subscripts, *operands, **kw)


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
