"""``Array`` -- the user-facing collection, API-compatible with ``dask_array.Array`` for the
hot path (``dask_array/_collection.py:110-1761``): operators build ``Elemwise`` nodes
(:715-877), methods build reductions (:1300-1500), ``.T`` / ``rechunk`` / basic slices build
the data-movement nodes, ``compute()`` optimises and runs on the GPU(s), ``persist()`` keeps
the blocks device-resident under the same expression name (:285-300).
"""
from __future__ import annotations

import math
import operator
from numbers import Integral, Number

import numpy as np

from . import _codegen as cg
from ._blockwise import Elemwise, Transpose
from ._expr import ArrayExpr, BroadcastTrick, FromArray, HostBlocks, Random, Resident, normalize_chunks
from ._reductions import Reduction, validate_axis


def _as_operand(x):
    if isinstance(x, Array):
        return x.expr
    if isinstance(x, (bool, int, float, np.generic)):
        return x
    if isinstance(x, np.ndarray):
        if x.ndim == 0:
            return x[()]
        return FromArray(x, normalize_chunks("auto", x.shape, dtype=x.dtype))
    raise TypeError(f"cannot use {type(x).__name__} as an operand of a dask_array_b200 Array")


def elemwise(op, *args, out=None, where=True, dtype=None, **kwargs):
    """``elemwise()`` (``core/_blockwise_funcs.py:207-300``).  ``where=mask`` keeps ``out``'s values where the mask is
    False (``_elemwise_handle_where``, ``_core_utils.py:1031-1036``: the ufunc writes into a copy of ``out``), which
    is ``where(mask, op(...), out)`` inside the same fused kernel; without ``out`` those positions would be
    uninitialised memory in NumPy -- refused.  ``out=`` (an Array) is overwritten in place: it takes the result's
    expression (``handle_out``), and is returned.  ``dtype=`` casts the result (``_enforce_dtype`` :1078)."""
    name = cg.canonical_name(op)
    ops = tuple(_as_operand(a) for a in args)
    if not any(isinstance(o, ArrayExpr) for o in ops):
        raise TypeError("elemwise needs at least one Array operand")
    if isinstance(out, tuple):
        if len(out) != 1:
            raise NotImplementedError("The out parameter is not fully supported: one output only")
        out = out[0]
    if out is not None and not isinstance(out, Array):
        raise NotImplementedError("The out parameter is not fully supported. Received type "
                                  f"{type(out).__name__}, expected a dask_array_b200 Array")
    res = Array(Elemwise(name, ops, tuple(sorted(kwargs.items()))))
    if dtype is not None and np.dtype(dtype) != res.dtype:
        res = res.astype(dtype)
    if where is not True:
        if where is False or where is None:
            where = np.bool_(False)
        if out is None:
            raise NotImplementedError("elemwise(where=mask) without out=: the unselected positions are uninitialised "
                                      "memory in NumPy and have no defined value to reproduce")
        keep = out if out.dtype == res.dtype else out.astype(res.dtype)
        res = Array(Elemwise("where", (_as_operand(where), res.expr, keep.expr), ()))
    if out is not None:
        if out.shape != res.shape:
            raise ValueError(f"Mismatched shapes between `out` parameter and result: {out.shape} vs {res.shape}")
        if res.dtype != out.dtype:
            res = res.astype(out.dtype)
        out.expr = res.expr
        return out
    return res


_NUMPY_ALIASES = {"amin": "min", "amax": "max", "round_": "round", "concat": "concatenate", "permute_dims": "transpose"}


class Array:
    __array_priority__ = 11

    def __init__(self, expr: ArrayExpr):
        self.expr = expr

    # ---- metadata
    shape = property(lambda self: self.expr.shape)
    dtype = property(lambda self: self.expr.dtype)
    chunks = property(lambda self: self.expr.chunks)
    ndim = property(lambda self: self.expr.ndim)
    numblocks = property(lambda self: self.expr.numblocks)
    size = property(lambda self: self.expr.size)
    nbytes = property(lambda self: self.expr.nbytes)
    name = property(lambda self: self.expr._name)

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __repr__(self):
        return f"dask_array_b200.Array<{self.name}, shape={self.shape}, dtype={self.dtype}, chunks={self.chunks}>"

    # ---- optimisation / execution
    def optimize(self, fuse=True):
        return Array(self.expr.optimize(fuse=fuse))

    def pprint(self):
        self.expr.pprint()

    def _execute(self):
        from ._executor import Executor

        opt = self.expr.optimize()
        ex = Executor()
        return ex, opt, ex.run(opt)

    def compute(self, **kwargs):
        from ._executor import gather_to_host

        ex, opt, store = self._execute()
        return gather_to_host(ex, opt, store)

    def compile(self):
        """Optimise, run once and keep the launch tape: ``Compiled.run()`` replays the same
        kernels on the same buffers (what a steady-state ``compute()`` of a persisted graph does),
        without re-planning."""
        return Compiled(self)

    def persist(self, **kwargs):
        """Blocks stay on the GPU(s); the returned Array reads them in place."""
        ex, opt, store = self._execute()
        if store.kind != "array":
            raise NotImplementedError("persist of a partial reduction state")
        # keep dependencies' buffers alive through the store itself
        store.keepalive.append([s for s in ex.results.values() if s is not store])
        return Array(Resident(store, opt.chunks, opt.dtype, self.expr._name))

    def __array__(self, dtype=None, copy=None):
        out = self.compute()
        return np.asarray(out, dtype=dtype)

    # ---- element-wise operators (``_collection.py:715-877``)
    def _bin(self, op, other, reverse=False):
        if not isinstance(other, (Array, Number, np.generic, np.ndarray, bool)):
            return NotImplemented
        return elemwise(op, other, self) if reverse else elemwise(op, self, other)

    __add__ = lambda s, o: s._bin(operator.add, o)
    __radd__ = lambda s, o: s._bin(operator.add, o, True)
    __sub__ = lambda s, o: s._bin(operator.sub, o)
    __rsub__ = lambda s, o: s._bin(operator.sub, o, True)
    __mul__ = lambda s, o: s._bin(operator.mul, o)
    __rmul__ = lambda s, o: s._bin(operator.mul, o, True)
    __truediv__ = lambda s, o: s._bin(operator.truediv, o)
    __rtruediv__ = lambda s, o: s._bin(operator.truediv, o, True)
    __floordiv__ = lambda s, o: s._bin(operator.floordiv, o)
    __rfloordiv__ = lambda s, o: s._bin(operator.floordiv, o, True)
    __mod__ = lambda s, o: s._bin(operator.mod, o)
    __rmod__ = lambda s, o: s._bin(operator.mod, o, True)
    __pow__ = lambda s, o: s._bin(operator.pow, o)
    __rpow__ = lambda s, o: s._bin(operator.pow, o, True)
    __and__ = lambda s, o: s._bin(operator.and_, o)
    __rand__ = lambda s, o: s._bin(operator.and_, o, True)
    __or__ = lambda s, o: s._bin(operator.or_, o)
    __ror__ = lambda s, o: s._bin(operator.or_, o, True)
    __xor__ = lambda s, o: s._bin(operator.xor, o)
    __rxor__ = lambda s, o: s._bin(operator.xor, o, True)
    __lshift__ = lambda s, o: s._bin(operator.lshift, o)
    __rshift__ = lambda s, o: s._bin(operator.rshift, o)
    __lt__ = lambda s, o: s._bin(operator.lt, o)
    __le__ = lambda s, o: s._bin(operator.le, o)
    __gt__ = lambda s, o: s._bin(operator.gt, o)
    __ge__ = lambda s, o: s._bin(operator.ge, o)
    __eq__ = lambda s, o: s._bin(operator.eq, o)
    __ne__ = lambda s, o: s._bin(operator.ne, o)
    __neg__ = lambda s: elemwise(operator.neg, s)
    __pos__ = lambda s: elemwise(operator.pos, s)
    __abs__ = lambda s: elemwise(operator.abs, s)
    __invert__ = lambda s: elemwise(operator.invert, s)
    __hash__ = None

    def __matmul__(self, other):
        from ._matmul import matmul

        return matmul(self, other)

    def __array_function__(self, func, types, args, kwargs):
        """``Array.__array_function__`` (:866-923): a NumPy function applied to an Array runs the same-named function
        of this package and stays lazy (``np.sum(x, axis=0)``, ``np.clip``, ``np.reshape``, ``np.concatenate`` ...).
        A function the package does not have is handled like the reference does: a ``FutureWarning``, the Arrays among
        the arguments are computed, and NumPy runs on the results."""
        import warnings

        import dask_array_b200 as module

        if not all(issubclass(t, (Array, np.ndarray)) or t.__name__ == "DeviceChunk" for t in types):
            return NotImplemented
        name = _NUMPY_ALIASES.get(func.__name__, func.__name__)
        target = getattr(module, name, None) if getattr(func, "__module__", None) == "numpy" else None
        if target is None or target is func or not callable(target):
            warnings.warn(f"The `{getattr(func, '__module__', 'numpy')}.{func.__name__}` function is not implemented by "
                          "dask_array_b200: the arguments are computed and NumPy runs on the host results.", FutureWarning)

            def host(v):
                if isinstance(v, Array):
                    return v.compute()
                if isinstance(v, (list, tuple)):
                    return type(v)(host(u) for u in v)
                return v
            return func(*host(args), **{k: host(v) for k, v in kwargs.items()})
        return target(*args, **kwargs)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """``Array.__array_ufunc__`` (:1702): NumPy ufuncs on Arrays stay lazy (``out=`` / ``where=`` / ``dtype=``
        included; the two-output ufuncs ``frexp / modf / divmod`` return a pair of Arrays)."""
        if method != "__call__" or set(kwargs) - {"out", "where", "dtype"}:
            return NotImplemented
        if ufunc.__name__ == "matmul":                # a gufunc, not element-wise: the blocked contraction
            if kwargs:
                return NotImplemented
            return matmul(*inputs)
        if ufunc.nout == 2 and ufunc.__name__ in ("frexp", "modf", "divmod") and not kwargs:
            return globals()[ufunc.__name__](*inputs)
        if ufunc.nout != 1:
            return NotImplemented
        return elemwise(ufunc.__name__, *inputs, **kwargs)

    def astype(self, dtype, **kwargs):
        """``Array.astype`` (:1569-1609)."""
        dtype = np.dtype(dtype)
        if dtype == self.dtype:
            return self
        return Array(Elemwise("astype", (self.expr,), (("dtype", dtype.name),)))

    # ---- data movement
    def transpose(self, *axes):
        """``Array.transpose`` (:934-953)."""
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes or axes == (None,):
            axes = tuple(reversed(range(self.ndim)))
        axes = tuple(a % self.ndim for a in axes)
        if sorted(axes) != list(range(self.ndim)):
            raise ValueError("axes don't match array")
        return Array(Transpose(self.expr, axes))

    T = property(lambda self: self.transpose())

    def swapaxes(self, axis1, axis2):
        return swapaxes(self, axis1, axis2)

    def reshape(self, *shape, merge_chunks=True, limit=None):
        """``Array.reshape`` (``manipulation/_reshape.py:460-522``)."""
        from ._reshape import reshape

        if len(shape) == 1 and not isinstance(shape[0], Integral):
            shape = shape[0]
        return reshape(self, shape, merge_chunks=merge_chunks, limit=limit)

    # ---- small conveniences of the reference's collection (``_collection.py:585-700, 1611-1700``)
    itemsize = property(lambda self: self.dtype.itemsize)
    npartitions = property(lambda self: math.prod(self.numblocks))
    chunksize = property(lambda self: tuple(max(c) for c in self.chunks))
    real = property(lambda self: real(self))
    imag = property(lambda self: imag(self))

    def copy(self):
        """Arrays are immutable descriptions: a copy shares the expression (``Array.copy`` :1688)."""
        return Array(self.expr)

    def conj(self):
        return conj(self)

    def clip(self, min=None, max=None):
        return clip(self, min, max)

    def round(self, decimals=0):
        return round(self, decimals)

    def rechunk(self, chunks="auto", threshold=None, block_size_limit=None, balance=False, method=None):
        """``Array.rechunk`` (:1056).  ``threshold`` / ``method`` shape the reference's task graph only: the
        re-blocking here is always one gather pass (``TasksRechunk``)."""
        from ._rechunk import rechunk

        return Array(rechunk(self.expr, chunks, block_size_limit, balance))

    def __getitem__(self, index):
        from ._slicing import SliceSlicesIntegers, normalize_index

        if not isinstance(index, tuple):
            index = (index,)
        if any(i is None for i in index):
            # np.newaxis entries (``slice_with_newaxes``, slicing/_basic.py): slice without them, then
            # insert the unit axes where they land in the result (integers drop their axis)
            from ._views import expand_dims

            if Ellipsis in index:
                k = index.index(Ellipsis)
                n_real = sum(1 for i in index if i is not None and i is not Ellipsis)
                index = index[:k] + (slice(None),) * (self.ndim - n_real) + index[k + 1:]
            out = self[tuple(i for i in index if i is not None)]
            pos = 0
            for i in index:
                if i is None:
                    out = expand_dims(out, pos)
                    pos += 1
                elif not isinstance(i, Integral):
                    pos += 1
            return out
        return Array(SliceSlicesIntegers(self.expr, normalize_index(index, self.shape)))

    # ---- reductions (``_collection.py:1300-1500`` -> ``reductions/_common.py``)
    def _reduce(self, kind, axis=None, keepdims=False, dtype=None, split_every=None, ddof=0):
        if kind in ("argmin", "argmax") and axis is not None and not isinstance(axis, Integral):
            raise TypeError(f"axis must be either `None` or int, got '{axis}'")
        ax = validate_axis(axis, self.ndim)
        if ax == () and (self.ndim > 0 or kind not in ("argmin", "argmax")):
            # axis=(): nothing is reduced (NumPy semantics) -- an element-wise identity / cast
            from ._reductions import result_dtype

            dt = result_dtype(kind, self.dtype, dtype)
            if kind in ("sum", "prod", "mean"):
                return self.astype(dt)
            if kind == "var":
                return (self - self).astype(dt)
            if kind in ("any", "all"):
                return self != 0
            if kind in ("min", "max", "nanmin", "nanmax"):
                return self
            raise TypeError(f"axis=() is not valid for {kind}")
        from ._reductions import normalize_split_every

        # canonical {axis: n} form at construction, so equivalent spellings share a name
        # (reductions/_reduction.py:715-725)
        se = normalize_split_every(split_every, ax)
        return Array(Reduction(self.expr, kind, ax, bool(keepdims), None if dtype is None else np.dtype(dtype).name,
                               se, ddof))

    def sum(self, axis=None, dtype=None, keepdims=False, split_every=None):
        return self._reduce("sum", axis, keepdims, dtype, split_every)

    def prod(self, axis=None, dtype=None, keepdims=False, split_every=None):
        return self._reduce("prod", axis, keepdims, dtype, split_every)

    def mean(self, axis=None, dtype=None, keepdims=False, split_every=None):
        return self._reduce("mean", axis, keepdims, dtype, split_every)

    def var(self, axis=None, dtype=None, keepdims=False, ddof=0, split_every=None):
        return self._reduce("var", axis, keepdims, dtype, split_every, ddof)

    def std(self, axis=None, dtype=None, keepdims=False, ddof=0, split_every=None):
        """``std = sqrt(var)`` as an Elemwise on the aggregate (``_common.py:625-653``)."""
        result = elemwise("sqrt", self.var(axis, dtype, keepdims, ddof, split_every))
        if dtype is not None and np.dtype(dtype) != result.dtype:
            result = result.astype(dtype)          # _common.py:650-652
        return result

    def min(self, axis=None, keepdims=False, split_every=None):
        return self._reduce("min", axis, keepdims, None, split_every)

    def max(self, axis=None, keepdims=False, split_every=None):
        return self._reduce("max", axis, keepdims, None, split_every)

    def dot(self, other):
        return self @ other

    def squeeze(self, axis=None):
        from ._views import squeeze

        return squeeze(self, axis)

    def ravel(self):
        from ._views import ravel

        return ravel(self)

    flatten = ravel

    def any(self, axis=None, keepdims=False, split_every=None):
        return self._reduce("any", axis, keepdims, None, split_every)

    def all(self, axis=None, keepdims=False, split_every=None):
        return self._reduce("all", axis, keepdims, None, split_every)

    def argmin(self, axis=None, keepdims=False, split_every=None):
        return self._reduce("argmin", axis, keepdims, None, split_every)

    def argmax(self, axis=None, keepdims=False, split_every=None):
        return self._reduce("argmax", axis, keepdims, None, split_every)

    def topk(self, k, axis=-1, split_every=None):
        from ._topk import topk

        return topk(self, k, axis=axis, split_every=split_every)

    def argtopk(self, k, axis=-1, split_every=None):
        from ._topk import argtopk

        return argtopk(self, k, axis=axis, split_every=split_every)

    def map_blocks(self, func, *args, **kwargs):
        from ._overlap import map_blocks

        return map_blocks(func, self, *args, **kwargs)

    def map_overlap(self, func, depth=None, boundary=None, trim=True, **kwargs):
        from ._overlap import map_overlap

        return map_overlap(func, self, depth=depth, boundary=boundary, trim=trim, **kwargs)

    def cumsum(self, axis=None, dtype=None, out=None, method="sequential"):
        return cumsum(self, axis=axis, dtype=dtype, out=out, method=method)

    def cumprod(self, axis=None, dtype=None, out=None, method="sequential"):
        return cumprod(self, axis=axis, dtype=dtype, out=out, method=method)


class Compiled:
    """Computed expression(s) plus the replayable launch tape.  Several arrays compiled
    together share one executor, hence common sub-expressions (uploads, rechunks)."""

    def __init__(self, *arrays: "Array"):
        from ._executor import Executor

        self.executor = Executor()
        self.exprs = [a.expr.optimize() for a in arrays]
        self.stores, self.segments = [], []          # segments: tape index range each top-level expression added
        for e in self.exprs:
            lo = len(self.executor.tape)
            self.stores.append(self.executor.run(e))
            self.segments.append((lo, len(self.executor.tape)))
        self.tape = list(self.executor.tape)
        self._graph = None

    def _independent_lanes(self):
        """Tape segments of top-level expressions that may run on concurrent streams inside the captured graph:
        no computed sub-expression in common (leaves that are already resident do not count) and at most one
        segment with cross-rank synchronisation (those keep one global order).  Returns [] when the tape has to
        stay on one stream."""
        import os

        from . import _executor

        if os.environ.get("B2_GRAPH_LANES", "1") != "1" or len(self.exprs) < 2 or _executor._NVTX >= 2:
            return []
        def names(e, acc):
            if e._name in acc:
                return acc
            acc.add(e._name)
            for d in e.dependencies():
                names(d, acc)
            return acc

        lanes = []
        coll = set(self.executor.collectives)
        n_coll = 0
        for e, (lo, hi) in zip(self.exprs, self.segments):
            if hi <= lo:
                continue
            for n in names(e, set()):
                a, b = self.executor.tape_span.get(n, (0, 0))
                if b > a and not (lo <= a and b <= hi):
                    return []            # reads something another segment's launches produce
            if any(lo <= c < hi for c in coll):
                n_coll += 1
            lanes.append((lo, hi))
        covered = sum(hi - lo for lo, hi in lanes)
        if n_coll > 1 or len(lanes) < 2 or covered != len(self.tape):
            return []
        return lanes

    def run(self):
        if self._graph is not None:
            self._graph.replay()
            return
        for fn in self.tape:
            fn()

    def capture(self):
        """Capture the launch tape into ONE CUDA graph: a replay is then a single graph launch instead
        of one driver call per kernel (the README example drops from ~11 us per launch to the graph's
        fixed cost).  Multi-GPU steps are capturable on the peer-memory path: its barriers keep their
        epoch in a device counter (``b2_peer_barrier_dev``), so the captured launches are argument-stable;
        every rank must then replay the same number of times."""
        import os

        import torch

        from . import _peer

        if self.executor.world.size > 1 and not _peer.enabled():
            raise NotImplementedError("CUDA-graph capture of a multi-GPU step needs the peer-memory path (B2_COMM=peer)")
        for k in self.fused_launches():
            if k.profile:
                raise RuntimeError("per-launch event profiling and graph capture are mutually exclusive")
        torch.cuda.synchronize()
        lanes = self._independent_lanes()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if not lanes:
                for fn in self.tape:
                    fn()
            else:
                # independent expressions become parallel branches of the graph: the tail of one expression's big
                # kernel (its last partial wave) overlaps the start of the next one's instead of idling the SMs
                # The lane with the longest chain of launches (the std() tree: chunk kernel -> exchange -> two levels
                # -> sqrt) runs on a HIGH-priority stream: its big kernel is scheduled first, and its serial tail of
                # small launches then overlaps the other lanes' big kernels instead of trailing the whole step.
                main = torch.cuda.current_stream()
                order = sorted(range(len(lanes)), key=lambda i: lanes[i][0] - lanes[i][1])      # longest first
                prio = os.environ.get("B2_LANE_PRIORITY", "1") == "1"
                streams = {}
                for rank_, i in enumerate(order):
                    if rank_ == 0 and prio:
                        streams[i] = torch.cuda.Stream(priority=-1)
                    elif (rank_ == 1 and prio) or (rank_ == 0 and not prio):
                        streams[i] = main
                    else:
                        streams[i] = torch.cuda.Stream()
                for s_ in streams.values():
                    if s_ is not main:
                        s_.wait_stream(main)
                for i, (lo, hi) in enumerate(lanes):
                    with torch.cuda.stream(streams[i]):
                        for fn in self.tape[lo:hi]:
                            fn()
                for s_ in streams.values():
                    if s_ is not main:
                        main.wait_stream(s_)
        self._graph = g
        self.lanes = len(lanes) or 1
        return self

    def results(self):
        from ._executor import gather_to_host

        return [gather_to_host(self.executor, e, s) for e, s in zip(self.exprs, self.stores)]

    def result(self):
        return self.results()[0]

    def fused_launches(self):
        from ._runtime import FusedLaunch

        out = []
        for st in self.executor.results.values():
            out.extend(k for k in st.keepalive if isinstance(k, FusedLaunch))
        return out


# ----------------------------------------------------------------------------- cumulative scans
def _cumulative(kind, x, axis, dtype, out, method, nan=False):
    """``_cumreduction_expr`` (``reductions/_cumulative.py:425-448``).  ``method`` ("sequential" |
    "blelloch") selects between two task-graph shapes in the reference; the B200 path has one implementation
    (reduce -> scan of the block totals -> scan with carry, or the chained single pass) which -- like
    "blelloch" -- combines block totals with each other before applying them: same values up to rounding,
    except that a product whose INTERMEDIATE block-total products overflow can give inf/NaN where a strict
    left-to-right product would not (the reference's own blelloch test expects that warning,
    tests/test_reductions.py:800-803)."""
    from ._reductions import CumReduction

    if out is not None:
        raise NotImplementedError("out= is not supported")
    if method not in ("sequential", "blelloch"):
        raise ValueError("Invalid method for cumulative reduction: choose 'sequential' or 'blelloch'")
    x = asarray(x)
    if axis is None:
        if x.ndim > 1:
            from ._views import ravel

            x = ravel(x)            # ``_prepare_cumulative`` (:77-97): flatten first, then scan the vector
        axis = 0
    axis = validate_axis(axis, x.ndim)[0]
    return Array(CumReduction(x.expr, kind, axis, None if dtype is None else np.dtype(dtype).name, bool(nan)))


def cumsum(x, axis=None, dtype=None, out=None, method="sequential"):
    """``cumsum`` (``reductions/_cumulative.py:451-484``)."""
    return _cumulative("cumsum", x, axis, dtype, out, method)


def cumprod(x, axis=None, dtype=None, out=None, method="sequential"):
    """``cumprod`` (``reductions/_cumulative.py:487-520``)."""
    return _cumulative("cumprod", x, axis, dtype, out, method)


def nancumsum(x, axis, dtype=None, out=None, *, method="sequential"):
    """``nancumsum`` (``reductions/_cumulative.py:523-557``)."""
    return _cumulative("cumsum", x, axis, dtype, out, method, nan=True)


def nancumprod(x, axis, dtype=None, out=None, *, method="sequential"):
    """``nancumprod`` (``reductions/_cumulative.py:560-594``)."""
    return _cumulative("cumprod", x, axis, dtype, out, method, nan=True)


# ----------------------------------------------------------------------------- NaN-aware reducers
def _nan_to(x, value):
    """``where(isnan(x), value, x)`` fused in front of the reduction chunk step (the role of
    ``chunk.nansum`` / ``np.nanmin`` ... in ``reductions/_common.py:170-266``)."""
    x = asarray(x)
    if x.dtype.kind != "f":
        return x
    return elemwise("where", elemwise("isnan", x), value, x)


def nansum(a, axis=None, dtype=None, keepdims=False, split_every=None):
    """``_common.py:170-183``."""
    return _nan_to(a, 0).sum(axis=axis, dtype=dtype, keepdims=keepdims, split_every=split_every)


def nanprod(a, axis=None, dtype=None, keepdims=False, split_every=None):
    return _nan_to(a, 1).prod(axis=axis, dtype=dtype, keepdims=keepdims, split_every=split_every)


def _nancount(a, axis, keepdims, split_every):
    a = asarray(a)
    if a.dtype.kind != "f":
        n = 1
        for ax in validate_axis(axis, a.ndim):
            n *= a.shape[ax]
        return n
    return elemwise("logical_not", elemwise("isnan", a)).sum(axis=axis, keepdims=keepdims, split_every=split_every)


def nanmean(a, axis=None, dtype=None, keepdims=False, split_every=None):
    """``_common.py:346-365``: sum of the non-NaN values over their count (0/0 -> NaN, as NumPy)."""
    a = asarray(a)
    dt = np.dtype(dtype) if dtype is not None else np.mean(np.zeros((1,), dtype=a.dtype)).dtype
    total = nansum(a, axis=axis, dtype=dt, keepdims=keepdims, split_every=split_every)
    return elemwise("true_divide", total, _nancount(a, axis, keepdims, split_every)).astype(dt)


def nanvar(a, axis=None, dtype=None, keepdims=False, ddof=0, split_every=None):
    """``_common.py:596-622`` by its definition: two more passes (mean, then squared deviations of
    the non-NaN values); the fused kernels make each pass one read of ``a``."""
    a = asarray(a)
    dt = np.dtype(dtype) if dtype is not None else np.var(np.ones((1,), dtype=a.dtype)).dtype
    mu = nanmean(a, axis=axis, dtype=dt, keepdims=True, split_every=split_every)
    d = elemwise("subtract", a.astype(dt), mu)
    if a.dtype.kind == "f":          # drop the entries that were NaN in the INPUT (inf - inf stays NaN, as np.nanvar)
        d = elemwise("where", elemwise("isnan", a), 0, d)
    ss = elemwise("multiply", d, d).sum(axis=axis, dtype=dt, keepdims=keepdims, split_every=split_every)
    n = _nancount(a, axis, keepdims, split_every)
    den = elemwise("subtract", n, ddof) if isinstance(n, Array) else n - ddof
    if isinstance(den, Array):
        den = elemwise("where", elemwise("less_equal", den, 0), np.float64(np.nan), den)
    elif den <= 0:
        den = np.nan
    return elemwise("true_divide", ss, den).astype(dt)


def nanstd(a, axis=None, dtype=None, keepdims=False, ddof=0, split_every=None):
    result = elemwise("sqrt", nanvar(a, axis, dtype, keepdims, ddof, split_every))
    if dtype is not None and np.dtype(dtype) != result.dtype:
        result = result.astype(dtype)
    return result


def nanmin(a, axis=None, keepdims=False, split_every=None):
    """``_common.py:196-229``: NaNs skipped by the accumulator itself (``B2AccNanMinMax``)."""
    return asarray(a)._reduce("nanmin", axis, keepdims, None, split_every)


def nanmax(a, axis=None, keepdims=False, split_every=None):
    return asarray(a)._reduce("nanmax", axis, keepdims, None, split_every)


def nanargmin(a, axis=None, keepdims=False, split_every=None):
    """``_common.py:815-827`` (NaN -> +inf; an all-NaN slice is not diagnosed here)."""
    return _nan_to(a, np.inf).argmin(axis=axis, keepdims=keepdims, split_every=split_every)


def nanargmax(a, axis=None, keepdims=False, split_every=None):
    return _nan_to(a, -np.inf).argmax(axis=axis, keepdims=keepdims, split_every=split_every)


def tensordot(a, b, axes=2):
    """``linalg/_tensordot.py:45-136``."""
    from ._matmul import tensordot as _td

    return _td(a, b, axes=axes)


def einsum(*operands, **kwargs):
    """``_einsum.py:181-271``."""
    from ._matmul import einsum as _es

    return _es(*operands, **kwargs)


def dot(a, b):
    return asarray(a) @ asarray(b)


def _freeze(split_every):
    if isinstance(split_every, dict):
        return dict(split_every)
    return split_every


def compile(*arrays):   # noqa: A001  (mirrors dask.compute's variadic form)
    return Compiled(*arrays)


def compute(*arrays):
    """``dask.compute(a, b, ...)``: shared sub-expressions are evaluated once."""
    return tuple(Compiled(*arrays).results())


def from_host_blocks(get_block, shape, chunks, dtype, token=None):
    """Array whose blocks come from ``get_block(block id) -> host ndarray``."""
    chunks = normalize_chunks(chunks, tuple(shape), dtype=np.dtype(dtype))
    token = token if token is not None else f"{id(get_block):x}"
    return Array(HostBlocks(get_block, chunks, np.dtype(dtype).name, token))


# ----------------------------------------------------------------------------- creation
def from_array(x, chunks="auto", **kwargs):
    """``da.from_array`` (``io/_from_array.py``)."""
    from ._device import DeviceChunk

    if isinstance(x, DeviceChunk):
        return _from_device(x, chunks)
    x = np.asarray(x)
    return Array(FromArray(x, normalize_chunks(chunks, x.shape, dtype=x.dtype)))


def _from_device(x, chunks):
    """``from_array`` of an array that already lives on the GPU (``io/_from_array.py:148-152`` keeps such chunks in
    their own type): the blocks this rank owns are cut on the device -- views copied once into their own contiguous
    blocks, no host round trip -- and handed to the graph as resident blocks, like ``persist()`` leaves them."""
    from . import _eager
    from ._exchange import owner_of
    from ._executor import BlockStore, World

    blocks = normalize_chunks(chunks, x.shape, dtype=x.dtype)
    geometry = BroadcastTrick(0, tuple(x.shape), blocks, x.dtype.name)      # only its block grid is used
    store = BlockStore(geometry)
    world = World()
    for bid in geometry.block_ids():
        if owner_of(geometry, bid, world.size) != world.rank:
            continue
        start, shape = geometry.block_start(bid), geometry.block_shape(bid)
        view = x[tuple(slice(s, s + n) for s, n in zip(start, shape))] if x.ndim else x
        store.blocks[bid] = view if view.is_contiguous else _eager.copy(view)
    store.keepalive.append(x)
    token = f"{x.ptr:x}-{id(x):x}-{abs(hash(blocks)):x}"
    return Array(Resident(store, blocks, x.dtype.name, token))


def asarray(x, **kwargs):
    return x if isinstance(x, Array) else from_array(x, **kwargs)


def _creation(value, shape, chunks, dtype):
    shape = (shape,) if isinstance(shape, Integral) else tuple(shape)
    dtype = np.dtype(dtype if dtype is not None else np.float64)
    chunks = "auto" if chunks is None else chunks            # the reference's default (creation/_ones_zeros.py)
    return Array(BroadcastTrick(value, shape, normalize_chunks(chunks, shape, dtype=dtype), dtype.name))


def ones(shape, dtype=None, chunks=None, **kw):
    """``da.ones`` (``creation/_ones_zeros.py:124``)."""
    return _creation(1, shape, chunks, dtype)


def zeros(shape, dtype=None, chunks=None, **kw):
    return _creation(0, shape, chunks, dtype)


def full(shape, fill_value, dtype=None, chunks=None, **kw):
    if dtype is None:
        dtype = np.asarray(fill_value).dtype
    return _creation(fill_value, shape, chunks, dtype)


_NO_START = object()


def _host_sequence(label, params, dtype, chunks, make_block):
    import hashlib

    token = hashlib.sha1(repr((label, params, dtype.name, chunks)).encode()).hexdigest()[:16]
    return Array(HostBlocks(make_block, chunks, dtype.name, f"{label}-{token}"))


def arange(start=_NO_START, stop=None, step=1, *, chunks="auto", like=None, dtype=None):
    """``da.arange`` (``creation/_arange.py:127-185``): block ``k`` is ``np.arange`` over its own
    ``[start + first_k * step, start + (first_k + len_k) * step)`` (``Arange._layer`` :102-123, ``_chunk.arange``
    :320-332), generated on the host and staged once like the random streams."""
    if start is _NO_START:
        if stop is None:
            raise TypeError("arange() requires stop to be specified.")
        start = 0
    elif stop is None:
        start, stop = 0, start
    if start != 0 and not np.isclose(start + step - start, step, atol=0):
        # very large start with a small float step: build from zero and shift (:180-183)
        return arange(0, stop - start, step, chunks=chunks, dtype=dtype) + start
    num = int(max(np.ceil((stop - start) / step), 0))
    dt = np.dtype(dtype) if dtype is not None else np.arange(type(start)(0), type(stop)(0), step).dtype
    blocks = normalize_chunks(chunks, (num,), dtype=dt)
    firsts = np.concatenate([[0], np.cumsum(blocks[0])]).tolist()

    def make_block(bid):
        first, n = firsts[bid[0]], blocks[0][bid[0]]
        res = np.arange(start + first * step, start + (first + n) * step, step, dtype=dt)
        return res[:-1] if len(res) > n else res
    return _host_sequence("arange", (start, stop, step), dt, blocks, make_block)


def linspace(start, stop, num=50, endpoint=True, retstep=False, chunks="auto", dtype=None):
    """``da.linspace`` (``creation/_linspace.py:104-145``): ``step = (stop - start) / (num - 1 | num)``; block ``k``
    is ``np.linspace`` from its running start over its own length (``Linspace._layer`` :84-101)."""
    num = int(num)
    dt = np.dtype(dtype) if dtype is not None else np.linspace(0, 1, 1).dtype
    div = (num - 1) if endpoint else num
    step = float(stop - start) / (div if div else 1)
    blocks = normalize_chunks(chunks, (num,), dtype=dt)
    edges, at = [], start
    for n in blocks[0]:
        edges.append((at, at + ((n - 1) if endpoint else n) * step))
        at = at + step * n

    def make_block(bid):
        lo, hi = edges[bid[0]]
        return np.linspace(lo, hi, blocks[0][bid[0]], endpoint=endpoint, dtype=dt)
    out = _host_sequence("linspace", (start, stop, num, bool(endpoint)), dt, blocks, make_block)
    return (out, step) if retstep else out


class _RandomGenerator:
    """``da.random.default_rng(seed)`` (``random/_generator.py:425-441``): one live ``SeedSequence`` per
    generator (fresh OS entropy when ``seed`` is None); every draw spawns one child per block from it
    (``_spawn_bitgens``, ``random/_expr.py:29-32``), which advances the sequence -- successive draws differ,
    a fixed seed reproduces the same succession, exactly like the reference."""

    def __init__(self, seed=None):
        if isinstance(seed, np.random.SeedSequence):
            self._seed_seq = seed
        elif isinstance(seed, np.random.Generator):
            self._seed_seq = seed.bit_generator.seed_seq
        elif isinstance(seed, np.random.BitGenerator):
            self._seed_seq = seed.seed_seq
        else:
            self._seed_seq = np.random.SeedSequence(seed)

    def _make(self, dist, size, chunks, dtype, args=()):
        shape = () if size is None else (size,) if isinstance(size, Integral) else tuple(size)
        chunks = normalize_chunks("auto" if chunks is None else chunks, shape, dtype=np.dtype(dtype))   # random/_expr.py:86-90
        ss = self._seed_seq
        first = ss.n_children_spawned
        nblocks = 1
        for c in chunks:
            nblocks *= len(c)
        ss.spawn(nblocks)                       # advance, as _spawn_bitgens does
        ent = ss.entropy
        ent = tuple(int(e) for e in ent) if isinstance(ent, (list, tuple, np.ndarray)) else int(ent)
        seed = (ent, tuple(int(k) for k in ss.spawn_key), int(first))
        return Array(Random(seed, dist, shape, chunks, np.dtype(dtype).name, args))

    def random(self, size=None, dtype=np.float64, chunks="auto", **kw):
        return self._make("random", size, chunks, dtype)

    def standard_normal(self, size=None, dtype=np.float64, chunks="auto", **kw):
        return self._make("standard_normal", size, chunks, dtype)

    def integers(self, low, high=None, size=None, dtype=np.int64, chunks="auto", **kw):
        if high is None:
            low, high = 0, low
        return self._make("integers", size, chunks, dtype, (int(low), int(high)))


class _RandomModule:
    default_rng = staticmethod(lambda seed=None: _RandomGenerator(seed))

    @staticmethod
    def random(size=None, chunks="auto", dtype=np.float64, **kw):
        return _RandomGenerator(None).random(size, dtype=dtype, chunks=chunks)


random = _RandomModule()


# ----------------------------------------------------------------------------- free functions
def _ufunc(name):
    def f(*args, **kwargs):
        return elemwise(name, *args, **kwargs)
    f.__name__ = name
    f.__doc__ = f"Element-wise ``np.{name}`` (``_ufunc.py:284-392``)."
    return f


UFUNC_NAMES = [
    "add", "subtract", "multiply", "divide", "true_divide", "floor_divide", "negative", "positive", "power",
    "float_power", "remainder", "mod", "fmod", "exp", "exp2", "log", "log2", "log10", "log1p", "expm1",
    "logaddexp", "sqrt", "square", "cbrt", "reciprocal", "sin", "cos", "tan", "arcsin", "arccos", "arctan",
    "arctan2", "hypot", "sinh", "cosh", "tanh", "arcsinh", "arccosh", "arctanh", "deg2rad", "rad2deg",
    "degrees", "radians", "greater", "greater_equal", "less", "less_equal", "not_equal", "equal",
    "logical_and", "logical_or", "logical_xor", "logical_not", "maximum", "minimum", "fmax", "fmin",
    "bitwise_and", "bitwise_or", "bitwise_xor", "bitwise_not", "invert", "left_shift", "right_shift",
    "isfinite", "isinf", "isnan", "signbit", "copysign", "nextafter", "floor", "ceil", "trunc", "rint",
    "fabs", "sign", "absolute", "abs",
]


def where(cond, x, y):
    return elemwise("where", cond, x, y)


def divmod(x, y):   # noqa: A001
    """``divmod`` (``_ufunc.py:447-451``): ``(x // y, x % y)``."""
    return elemwise("floor_divide", x, y), elemwise("remainder", x, y)


def modf(x):
    """``modf`` (``_ufunc.py:438-444``): fractional and integral parts, both with the sign of ``x`` (``modf(-2.) ==
    (-0., -2.)``, ``modf(inf) == (0., inf)``), each ONE fused kernel."""
    x = asarray(x)
    if x.dtype.kind != "f":
        x = x.astype(np.result_type(x.dtype, np.float64) if x.dtype.itemsize > 2 else np.float32 if x.dtype.itemsize == 2 else np.float16)
    ip = elemwise("trunc", x)
    frac = elemwise("copysign", elemwise("where", elemwise("isinf", x), 0.0, elemwise("subtract", x, ip)), x)
    return frac.astype(x.dtype), ip


def frexp(x):
    """``frexp`` (``_ufunc.py:429-436``): mantissa and int32 exponent."""
    x = asarray(x)
    return elemwise("frexp_mantissa", x), elemwise("frexp_exponent", x)


def _method(name):
    def f(a, *args, **kwargs):
        return getattr(asarray(a), name)(*args, **kwargs)
    f.__name__ = name
    return f


def transpose(a, axes=None):
    return asarray(a).transpose(axes) if axes is not None else asarray(a).transpose()


def swapaxes(a, axis1, axis2):
    """``swapaxes`` (``manipulation/_transpose.py:243-261``)."""
    a = asarray(a)
    if axis1 == axis2:
        return a
    order = list(range(a.ndim))
    i, j = validate_axis(axis1, a.ndim)[0], validate_axis(axis2, a.ndim)[0]
    order[i], order[j] = order[j], order[i]
    return a.transpose(order)


def moveaxis(a, source, destination):
    """``moveaxis`` (``manipulation/_transpose.py:264-285``)."""
    a = asarray(a)
    src = tuple(s % a.ndim for s in ((source,) if isinstance(source, Integral) else source))
    dst = tuple(d % a.ndim for d in ((destination,) if isinstance(destination, Integral) else destination))
    if len(src) != len(dst):
        raise ValueError("`source` and `destination` arguments must have the same number of elements")
    order = [n for n in range(a.ndim) if n not in src]
    for d, s_ in sorted(zip(dst, src)):
        order.insert(d, s_)
    return a.transpose(order)


def rollaxis(a, axis, start=0):
    """``rollaxis`` (``manipulation/_transpose.py:288-313``)."""
    a = asarray(a)
    n = a.ndim
    axis = validate_axis(axis, n)[0]
    if start < 0:
        start += n
    if not 0 <= start < n + 1:
        raise ValueError("'%s' arg requires %d <= %s < %d, but %d was passed in" % ("start", -n, "start", n + 1, start))
    if axis < start:
        start -= 1
    if axis == start:
        return a
    order = list(range(n))
    order.remove(axis)
    order.insert(start, axis)
    return a.transpose(order)


def clip(a, a_min=None, a_max=None, **kwargs):
    """``clip`` (``_ufunc.py``): one-sided bounds are ``maximum`` / ``minimum``."""
    a_min = kwargs.pop("min", a_min)
    a_max = kwargs.pop("max", a_max)
    if kwargs:
        raise TypeError(f"clip() got unexpected arguments {sorted(kwargs)}")
    if a_min is None and a_max is None:
        raise ValueError("One of max or min must be given")
    if a_min is None:
        return elemwise("minimum", a, a_max)
    if a_max is None:
        return elemwise("maximum", a, a_min)
    return elemwise("clip", a, a_min, a_max)


def round(a, decimals=0):   # noqa: A001
    """``round`` / ``around`` (``_ufunc.py:453-462``): NumPy's own scheme -- scale by ``10**decimals``, ``rint``,
    scale back (``decimals < 0``: divide first) -- as one fused kernel; integers with ``decimals >= 0`` are unchanged."""
    a = asarray(a)
    decimals = int(decimals)
    if a.dtype.kind in "biu":
        if decimals >= 0:
            return a
        raise NotImplementedError("round of an integer array to negative decimals")
    if decimals == 0:
        return elemwise("rint", a)
    scale = a.dtype.type(10.0 ** abs(decimals))
    if decimals > 0:
        return elemwise("true_divide", elemwise("rint", elemwise("multiply", a, scale)), scale)
    return elemwise("multiply", elemwise("rint", elemwise("true_divide", a, scale)), scale)


around = round


def _no_complex(a):
    a = asarray(a)
    if a.dtype.kind == "c":
        raise NotImplementedError("complex arrays have no B200 kernels")
    return a


def real(a):
    return _no_complex(a)


def conj(a):
    return _no_complex(a)


conjugate = conj


def imag(a):
    a = _no_complex(a)
    return Array(BroadcastTrick(0, a.shape, a.chunks, a.dtype.name))


def rechunk(a, chunks="auto", threshold=None, block_size_limit=None, balance=False, method=None):
    return asarray(a).rechunk(chunks, threshold, block_size_limit, balance, method)


def matmul(a, b):
    from ._matmul import matmul as _mm

    return _mm(asarray(a), asarray(b))
