"""Activation of the B200 backend behind the unchanged ``import dask_array as da``.

The reference has no FFI: a replacement plugs in through Python duck typing at three levels.  This module
is the reference-facing side of all three; everything below it is ``libb200da.so`` (``include/b200da.h``).

``register()``            B1   ``register_chunk_type(DeviceChunk)``                  ``_chunk_types.py:31-54``
                          B1'  ``concatenate_lookup / tensordot_lookup``             ``_core_utils.py:60-61``
                               ``einsum / divide / numel / nannumel / empty / take`` ``_dispatch.py:145-151,248-256``
                          B1'' creation backend ``"b200"``                           ``_backends_array.py:14-99``
                               collection type for ``DeviceChunk`` metas             ``_backends.py:27-35,95-103``
``lower_reference(expr)`` B2   a reference expression tree (``Elemwise``, ``Transpose``, ``Ones/Zeros/Full``,
                               ``FromArray``, typed ``Reduction``s, ``Blockwise`` chunk steps, ``PartialReduce``,
                               ``ArgChunk``, ``Rechunk``, ``SliceSlicesIntegers``, ``FusedBlockwise``) becomes
                               this package's expression tree: ONE kernel per ``FusedBlockwise`` node
                               (``_blockwise.py:1574-1728``) instead of one Python task per block
``FusedPlan.from_reference`` the kernel program of one reference ``FusedBlockwise`` (its ``.exprs`` walked root
                               first, reading ``.op / .elemwise_args / .user_kwargs / .axes``)
``compute(x)``                 lower + optimise + run + gather: what ``x.compute(scheduler=...)`` cannot express,
                               because a dask scheduler only sees the materialised per-block graph
``get(dsk, keys)``        B3   a dask scheduler (``Array.__dask_scheduler__``, ``_collection.py:111``): runs ANY
                               materialised graph, every chunk function executing on ``DeviceChunk`` through
                               NEP-13 / NEP-18 (correct, one launch per NumPy call; the fused path is ``compute``)

Nothing here imports the reference at module import time: ``register()`` imports ``dask_array`` when called.
The reference's classes are recognised by NAME and read through their public attributes only, so the module
is testable without ``dask`` (``tests/test_plugin.py`` uses stand-in objects and, in the build container, the
reference's own modules through ``tests/golden/_refshim.py``).
"""
from __future__ import annotations

import functools
import importlib
from numbers import Number

import numpy as np

from . import _codegen as cg
from . import _eager
from ._device import DeviceChunk

_REGISTERED = {}


# ----------------------------------------------------------------------------- B1 / B1' / B1''
def _nannumel(x, **kwargs):
    """``nannumel_lookup`` (``_dispatch.py:241-243``): ``np.sum(~np.isnan(x), **kwargs)`` on the device."""
    return np.sum(np.logical_not(np.isnan(x)), **kwargs)


def _empty(shape, dtype=float, order="C", **kw):
    """``empty_lookup`` (``_dispatch.py:250``)."""
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
    return DeviceChunk.empty(shape, np.dtype(dtype))


def _take(a, indices, axis=0, **kw):
    """``take_lookup`` (``_dispatch.py:248``): 1-D integer gathers (the form ``_arg_combine`` needs)."""
    return np.take(a, indices, axis=axis)


def to_device(x):
    """A host block becomes a ``DeviceChunk`` (``from_array`` / ``to_backend`` per-block step)."""
    if isinstance(x, DeviceChunk):
        return x
    return DeviceChunk.from_numpy(np.asarray(x))


def make_backend_entrypoint(base):
    """``ArrayBackendEntrypoint`` subclass (``_backends_array.py:14-82``) creating ``DeviceChunk`` blocks."""

    class B200BackendEntrypoint(base):
        @property
        def RandomState(self):
            # host RNG (same bit streams as the reference's numpy backend, random/_expr.py:29-62); blocks are
            # uploaded by the leaf task.  A device RNG would not be bit-compatible with the reference.
            return np.random.RandomState

        @property
        def default_bit_generator(self):
            return np.random.PCG64

        @staticmethod
        def _filled(shape, value, dtype):
            shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
            dt = np.dtype(dtype if dtype is not None else np.asarray(value).dtype if value is not None else float)
            out = DeviceChunk.empty(shape, dt)
            if value is not None and out.size:
                from . import _runtime as rt

                rt.fill(out, value)
            return out

        @classmethod
        def ones(cls, shape, *, dtype=None, meta=None, **kwargs):
            return cls._filled(shape, 1, dtype or float)

        @classmethod
        def zeros(cls, shape, *, dtype=None, meta=None, **kwargs):
            return cls._filled(shape, 0, dtype or float)

        @classmethod
        def empty(cls, shape, *, dtype=None, meta=None, **kwargs):
            return cls._filled(shape, None, dtype or float)

        @classmethod
        def full(cls, shape, fill_value, *, dtype=None, meta=None, **kwargs):
            return cls._filled(shape, fill_value, dtype)

        @staticmethod
        def arange(start, /, stop=None, step=1, *, dtype=None, meta=None, **kwargs):
            host = np.arange(start, stop, step, dtype=dtype) if stop is not None else np.arange(start, dtype=dtype)
            return to_device(host)

        @classmethod
        def to_backend_dispatch(cls):
            return to_device

        @classmethod
        def to_backend(cls, data, **kwargs):
            """``to_backend`` (``creation/_utils.py:277-301``): every block uploaded by ``map_blocks``."""
            meta = DeviceChunk.empty((0,) * data.ndim, data.dtype) if _gpu() else None
            return data.map_blocks(to_device, meta=meta) if meta is not None else data.map_blocks(to_device)

    return B200BackendEntrypoint


def _gpu() -> bool:
    import torch

    return torch.cuda.is_available()


def register(dask_array=None) -> dict:
    """Perform the B1 / B1' / B1'' registrations against the reference package (``dask_array``, imported
    here unless passed in).  Idempotent.  Returns what was registered, by name."""
    mod = lambda name: importlib.import_module(name)      # noqa: E731
    done = {}
    ct = mod("dask_array._chunk_types")
    if DeviceChunk not in ct._HANDLED_CHUNK_TYPES:
        ct.register_chunk_type(DeviceChunk)                                      # _chunk_types.py:31
    done["chunk_type"] = DeviceChunk
    cu = mod("dask_array._core_utils")
    cu.concatenate_lookup.register(DeviceChunk, _eager.concatenate)              # _core_utils.py:60,1252
    cu.tensordot_lookup.register(DeviceChunk, _eager.tensordot)                  # _core_utils.py:61,1256
    dp = mod("dask_array._dispatch")
    for name, impl in (("einsum_lookup", _eager.einsum), ("divide_lookup", _eager.divide), ("numel_lookup", _eager.numel),
                       ("nannumel_lookup", _nannumel), ("empty_lookup", _empty), ("take_lookup", _take)):
        getattr(dp, name).register(DeviceChunk, impl)                            # _dispatch.py:145-151
        done[name] = impl
    done["concatenate_lookup"], done["tensordot_lookup"] = _eager.concatenate, _eager.tensordot
    ba = mod("dask_array._backends_array")
    entry = make_backend_entrypoint(ba.ArrayBackendEntrypoint)()
    ba.array_creation_dispatch.register_backend("b200", entry)                   # _backends_array.py:91-99
    done["backend"] = entry
    try:
        bk = mod("dask_array._backends")
        bk._register_collection_type(DeviceChunk, bk.get_collection_type_array)  # _backends.py:27-35
        done["collection_type"] = True
    except Exception as e:          # noqa: BLE001  (optional: the object fallback already maps to Array)
        done["collection_type"] = repr(e)
    _REGISTERED.update(done)
    return done


# ----------------------------------------------------------------------------- B2: expression adapter
def _cls(e) -> str:
    return type(e).__name__


def _funcnames(f) -> list:
    """Names of the plain functions inside ``partial`` / toolz ``compose`` wrappers, outermost first."""
    out, stack = [], [f]
    while stack:
        g = stack.pop(0)
        if isinstance(g, functools.partial):
            stack.insert(0, g.func)
            continue
        inner = []
        if hasattr(g, "first") and hasattr(g, "funcs"):            # toolz.functoolz.Compose: first, then funcs
            inner = list(reversed(tuple(g.funcs))) + [g.first]
        elif hasattr(g, "funcs"):
            inner = list(g.funcs)
        if inner:
            stack = inner + stack
            continue
        out.append(getattr(g, "__name__", type(g).__name__))
    return out


def _partial_kwargs(f) -> dict:
    kw, stack = {}, [f]
    while stack:
        g = stack.pop()
        if isinstance(g, functools.partial):
            for k, v in (g.keywords or {}).items():
                kw.setdefault(k, v)
            stack.append(g.func)
        elif hasattr(g, "funcs"):
            stack.extend(list(g.funcs) + ([g.first] if hasattr(g, "first") else []))
    return kw


_CHUNK_KINDS = {
    "sum": "sum", "prod": "prod", "chunk_min": "min", "chunk_max": "max", "min": "min", "max": "max", "amin": "min",
    "amax": "max", "any": "any", "all": "all", "mean_chunk": "mean", "moment_chunk": "var",
}
_AGG_KINDS = {
    "sum": "sum", "prod": "prod", "min": "min", "max": "max", "amin": "min", "amax": "max", "chunk_min": "min",
    "chunk_max": "max", "any": "any", "all": "all", "mean_agg": "mean", "mean_combine": "mean", "moment_agg": "var",
    "moment_combine": "var", "arg_agg": "arg", "arg_combine": "arg",
}
_TYPED = {"Sum": "sum", "Prod": "prod", "Min": "min", "Max": "max", "Any": "any", "All": "all", "Mean": "mean", "Var": "var"}
_TYPED_NAN = {"NanSum": "nansum", "NanProd": "nanprod", "NanMin": "nanmin", "NanMax": "nanmax", "NanMean": "nanmean",
              "NanVar": "nanvar"}


def _attr(e, name, default=None):
    """Reference operands are reachable as attributes (``Expr.__getattr__`` over ``_parameters``) and through
    ``operand(name)``; stand-in objects in tests only have attributes."""
    if hasattr(e, name):
        return getattr(e, name)
    op = getattr(e, "operand", None)
    if op is not None:
        try:
            return op(name)
        except Exception:       # noqa: BLE001
            return default
    return default


def _axes(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    if isinstance(axis, (int, np.integer)):
        axis = (axis,)
    return tuple(sorted(a % ndim for a in axis))


def _contraction_kind(e):
    """``"_tensordot"`` / ``"_matmul"`` when ``e`` is the reference's contraction Blockwise, else None."""
    if _cls(e) != "Blockwise":
        return None
    names = _funcnames(e.func)
    return next((n for n in ("_tensordot", "_matmul") if n in names), None)


def _cast(e, dtype):
    from ._blockwise import Elemwise

    want = np.dtype(dtype)
    return e if e.dtype == want else Elemwise("astype", (e,), (("dtype", want.name),))


def _matmul_of(inner, rec):
    from . import _matmul as mm
    from ._collection import Array

    lhs, _, rhs, _ = list(inner.args)[:4]
    if lhs.ndim != 2 or rhs.ndim != 2:
        raise NotImplementedError("stacked / broadcast batch matmul has no B200 kernel")
    return mm.matmul(Array(rec(lhs)), Array(rec(rhs))).expr


def lower_reference(expr, _memo=None):
    """Convert a reference expression (``dask_array._expr.ArrayExpr`` tree, lowered or not, fused or not) into
    this package's expression tree.  Unknown node types raise ``NotImplementedError`` naming the class --
    never a silent host fallback."""
    from . import _blockwise as bw
    from . import _expr as ex
    from . import _reductions as red
    from ._collection import Array
    from ._rechunk import rechunk as _rechunk
    from ._slicing import SliceSlicesIntegers, normalize_index
    from . import _matmul as mm
    from . import _window as win
    from . import _views as vw

    memo = {} if _memo is None else _memo
    key = getattr(expr, "_name", None) or id(expr)
    if key in memo:
        return memo[key]
    rec = lambda e: lower_reference(e, memo)               # noqa: E731
    name = _cls(expr)
    if name == "FusedBlockwise":
        out = rec(expr.exprs[0])                           # members are reachable from the root; we re-fuse
    elif name == "Elemwise":
        if _attr(expr, "where", True) is not True or _attr(expr, "out") is not None:
            raise NotImplementedError("Elemwise with where= / out= has no B200 kernel")
        ops = []
        for a in expr.elemwise_args:
            if hasattr(a, "chunks") and hasattr(a, "dtype"):
                ops.append(rec(a))
            elif isinstance(a, np.ndarray) and a.ndim == 0:
                ops.append(a[()])
            elif isinstance(a, (Number, np.generic, bool)):
                ops.append(a)
            else:
                raise NotImplementedError(f"Elemwise operand of type {type(a).__name__}")
        kwargs = dict(_attr(expr, "user_kwargs", None) or {})
        out = bw.Elemwise(cg.canonical_name(expr.op), tuple(ops), tuple(sorted(kwargs.items())))
        want = _attr(expr, "dtype")
        if want is not None and np.dtype(want) != out.dtype:
            out = bw.Elemwise("astype", (out,), (("dtype", np.dtype(want).name),))
    elif name == "Transpose":
        out = bw.Transpose(rec(expr.array), tuple(expr.axes))
    elif name in ("Ones", "Zeros", "Full", "Empty", "BroadcastTrick"):
        kw = dict(_attr(expr, "kwargs", None) or {})
        value = {"Ones": 1, "Zeros": 0, "Empty": 0}.get(name, kw.get("fill_value", 0))
        out = ex.BroadcastTrick(value, tuple(expr.shape), ex.normalize_chunks(expr.chunks, tuple(expr.shape)),
                                np.dtype(expr.dtype).name)
    elif name == "Arange":
        from . import _collection as col

        out = col.arange(expr.start, expr.stop, expr.step, chunks=tuple(expr.chunks), dtype=np.dtype(expr.dtype)).expr
    elif name == "Linspace":
        from . import _collection as col

        out = col.linspace(expr.start, expr.stop, int(_attr(expr, "num", 50)), endpoint=bool(_attr(expr, "endpoint", True)),
                           chunks=tuple(expr.chunks), dtype=np.dtype(expr.dtype)).expr
    elif name == "FromArray":
        arr = expr.array
        if isinstance(arr, DeviceChunk):
            arr = arr.to_numpy()                          # re-blocked below; device leaves go through from_array
        out = ex.FromArray(np.asarray(arr), ex.normalize_chunks(expr.chunks, np.asarray(arr).shape))
    elif name == "Sum" and _contraction_kind(expr.array) == "_tensordot":
        # tensordot (linalg/_tensordot.py:100-136): Blockwise(_tensordot, concatenate=False) keeps the contracted
        # axes as size-1 chunks and ``.sum(axis=left_axes)`` folds them -- ONE accumulate-over-k node here
        inner = expr.array
        lhs, _, rhs, _ = list(inner.args)[:4]
        la, lb = ({**(_attr(inner, "kwargs", None) or {}), **_partial_kwargs(inner.func)})["axes"]
        la, lb = tuple(int(a) % lhs.ndim for a in la), tuple(int(b) % rhs.ndim for b in lb)
        if _axes(_attr(expr, "axis"), inner.ndim) != tuple(sorted(la)) or bool(_attr(expr, "keepdims", False)):
            raise NotImplementedError("a Sum over a tensordot partial that is not its own contraction fold")
        out = _cast(mm.tensordot(Array(rec(lhs)), Array(rec(rhs)), axes=(la, lb)).expr, expr.dtype)
    elif name == "Reduction" and _contraction_kind(expr.array) == "_matmul" \
            and "_chunk_sum" in _funcnames(_attr(expr, "chunk")):
        # matmul (linalg/_tensordot.py:253-334): Blockwise(_matmul) partials (M, 1, N) + ``_sum_wo_cat`` over k
        out = _cast(_matmul_of(expr.array, rec), expr.dtype)
    elif name == "Squeeze" and _contraction_kind(expr.array) == "_matmul" \
            and _axes(_attr(expr, "axis"), expr.array.ndim) == (expr.array.ndim - 2,):
        out = _cast(_matmul_of(expr.array, rec), expr.dtype)     # one k block: ``_sum_wo_cat`` is a squeeze (:243-246)
    elif name in ("Concatenate", "Stack"):
        arrs = [Array(rec(a)) for a in expr.args]          # ``args`` = array + the trailing operands (:24-25)
        out = (vw.concatenate if name == "Concatenate" else vw.stack)(arrs, axis=int(expr.axis)).expr
    elif name == "ExpandDims":
        x = Array(rec(expr.array))
        for ax in sorted(int(a) for a in expr.axes):       # positions in OUTPUT coordinates (manipulation/_expand.py:26)
            x = vw.expand_dims(x, ax)
        out = x.expr
    elif name == "Squeeze":
        out = vw.squeeze(Array(rec(expr.array)), axis=_attr(expr, "axis")).expr
    elif name == "BroadcastTo":
        out = vw.broadcast_to(Array(rec(expr.array)), tuple(expr._shape), chunks=_attr(expr, "_chunks")).expr
    elif name in ("CumReduction", "CumReductionBlelloch"):
        names = _funcnames(expr.func)
        kind = next((n for n in names if n in ("cumsum", "cumprod", "nancumsum", "nancumprod")), None)
        if kind is None:
            raise NotImplementedError(f"cumulative reduction {names} has no B200 kernel")
        x = rec(expr.array)
        dt = _attr(expr, "_dtype")
        out = red.CumReduction(x, kind[3:] if kind.startswith("nan") else kind, int(expr.axis) % x.ndim,
                               None if dt is None else np.dtype(dt).name, kind.startswith("nan"))
    elif name in ("Reshape", "ReshapeLowered"):
        from ._reshape import reshape as _reshape

        out = _reshape(Array(rec(expr.array)), tuple(int(n) for n in expr._shape)).expr
        if tuple(out.chunks) != tuple(tuple(c) for c in expr.chunks):
            raise NotImplementedError("a ReshapeLowered whose input was not blocked by reshape_rechunk")
    elif name == "Slice":
        out = Array(rec(expr.array))[tuple(expr.index)].expr
    elif name == "SlidingWindowReduction":
        x = rec(expr.array)
        out = win.SlidingWindowReduction(x, int(expr.window), int(expr.sliding_axis) % x.ndim, int(expr.window_axis),
                                         bool(expr.keepdims), str(expr.reducer), np.dtype(expr.dtype).name)
        if win.native_window_reduction(win.SlidingWindowView(x, (int(expr.window),), (int(expr.sliding_axis) % x.ndim,)),
                                       str(expr.reducer), (x.ndim,), bool(expr.keepdims), np.dtype(expr.dtype)) is None:
            raise NotImplementedError(f"SlidingWindowReduction({expr.reducer}, {np.dtype(expr.dtype)}) has no B200 kernel")
    elif name == "MovingWindowReduction":
        x = Array(rec(expr.array))
        out = _cast(win.moving_window(x, int(expr.window), str(expr.reducer), int(expr.min_count),
                                      int(expr.sliding_axis)).expr, expr.dtype)
    elif name in _TYPED or name in _TYPED_NAN:
        x = Array(rec(expr.array))
        axis = _axes(_attr(expr, "axis"), x.ndim)
        dt = np.dtype(expr.dtype)               # the node's result dtype (== the explicit dtype= when one was given)
        ddof = _partial_kwargs(_attr(expr, "aggregate")).get("ddof", 0) if name in ("Var", "NanVar") else 0
        kw = dict(axis=axis, keepdims=bool(_attr(expr, "keepdims", False)), split_every=_attr(expr, "split_every"))
        if name in _TYPED:
            kind = _TYPED[name]
            if kind in ("sum", "prod", "mean", "var"):
                kw["dtype"] = dt
            if kind == "var":
                kw["ddof"] = ddof
            out = getattr(x, kind)(**kw).expr
        else:
            from . import _collection as col

            fn = getattr(col, _TYPED_NAN[name])
            if name in ("NanSum", "NanProd", "NanMean", "NanVar"):
                kw["dtype"] = dt
            if name == "NanVar":
                kw["ddof"] = ddof
            out = fn(x, **kw).expr
    elif name == "Blockwise":
        names = _funcnames(expr.func)
        kind = next((_CHUNK_KINDS[n] for n in names if n in _CHUNK_KINDS), None)
        kw = {**(_attr(expr, "kwargs", None) or {}), **_partial_kwargs(expr.func)}
        if kind is None or "axis" not in kw:
            raise NotImplementedError(f"generic Blockwise({names}) has no B200 kernel (only reduction chunk steps)")
        arrays = [a for a in expr.args if hasattr(a, "chunks")]
        if len(arrays) != 1:
            raise NotImplementedError("weighted reduction chunk steps")
        x = rec(arrays[0])
        axis = _axes(kw["axis"], x.ndim)
        dt = red.result_dtype(kind, x.dtype, kw.get("dtype") if kind in ("sum", "prod", "mean", "var") else None)
        out = red.ChunkReduce(x, kind, axis, dt)
    elif name == "ArgChunk":
        x = rec(expr.array)
        names = _funcnames(_attr(expr, "chunk_func", _attr(expr, "chunk")))
        kind = "argmin" if any("min" in n for n in names) else "argmax"
        axis = _axes(_attr(expr, "axis"), x.ndim)
        ravel = len(axis) == x.ndim
        out = red.ArgChunk(x, kind, axis, ravel or x.ndim == 1)
    elif name == "PartialReduce":
        x = rec(expr.array)
        names = _funcnames(expr.func)
        kind = next((_AGG_KINDS[n] for n in names if n in _AGG_KINDS), None)
        if kind is None:
            raise NotImplementedError(f"PartialReduce({names}) has no B200 kernel")
        kw = _partial_kwargs(expr.func)
        if kind == "arg":
            leaf = x
            while isinstance(leaf, red.PartialReduce):
                leaf = leaf.operand("array")
            kind = leaf.operand("kind")
        se = {int(k): int(v) for k, v in dict(expr.split_every).items()}
        final = any(n.endswith("_agg") for n in names) or str(_attr(expr, "name", "") or "").endswith("-aggregate") \
            or not bool(_attr(expr, "keepdims", False))
        dt = red.result_dtype(kind, _leaf_dtype(x), _attr(expr, "dtype") if kind in ("sum", "prod", "mean", "var") else None)
        out = red.PartialReduce(x, kind, tuple(sorted(se)), se, bool(_attr(expr, "keepdims", False)), dt, bool(final),
                                kw.get("ddof", 0))
    elif name in ("Rechunk", "TasksRechunk"):
        out = _rechunk(rec(expr.array), _attr(expr, "_chunks", None) or expr.chunks)
    elif name == "SliceSlicesIntegers":
        x = rec(expr.array)
        out = SliceSlicesIntegers(x, normalize_index(tuple(expr.index), x.shape))
    else:
        raise NotImplementedError(f"reference expression {name} has no B200 lowering (hot-path subset: SURVEY.md 8a)")
    memo[key] = out
    return out


def _leaf_dtype(x):
    """dtype of the array a reduction tree reduces (below the chunk step)."""
    from . import _reductions as red

    while isinstance(x, red.PartialReduce):
        x = x.operand("array")
    if isinstance(x, (red.ChunkReduce, red.ArgChunk)):
        return x.operand("array").dtype
    return x.dtype


def fused_plan_from_reference(fused):
    """``FusedPlan.from_reference``: the kernel program of ONE reference ``FusedBlockwise`` node -- its
    ``exprs`` (root first, ``_blockwise.py:1591``) lowered member by member, external dependencies becoming the
    kernel's inputs in first-use order."""
    from . import _blockwise as bw

    if _cls(fused) != "FusedBlockwise":
        raise TypeError(f"expected a FusedBlockwise, got {_cls(fused)}")
    memo = {}
    members = tuple(lower_reference(e, memo) for e in fused.exprs)
    return bw.FusedPlan(bw.FusedBlockwise(members))


def compute(x, optimize: bool = True):
    """Run a reference collection / expression on the B200 backend: lower -> optimise (fusion) -> one launch per
    fused expression and device -> gather.  Returns a NumPy array like ``Array.compute()``."""
    from ._collection import Array

    expr = getattr(x, "expr", x)
    if optimize and hasattr(expr, "optimize") and _cls(expr) != "FusedBlockwise":
        try:
            expr = expr.optimize()
        except Exception:       # noqa: BLE001  (the reference's optimiser needs dask; ours runs below anyway)
            pass
    return Array(lower_reference(expr)).compute()


# ----------------------------------------------------------------------------- B3: scheduler
def _is_legacy_task(v) -> bool:
    return isinstance(v, tuple) and len(v) > 0 and callable(v[0])


def get(dsk, keys, to_host: bool = True, **kwargs):
    """A dask scheduler: ``x.compute(scheduler=dask_array_b200.get)`` / ``dask.config.set(scheduler=get)``
    (``_collection.py:111,158-191``).  Executes the materialised graph depth-first on the calling thread (GPU
    work is stream-ordered, the host never waits between tasks); every task is the reference's own chunk
    function, running on ``DeviceChunk`` blocks through NEP-13 / NEP-18.  Leaf blocks that arrive as host
    ndarrays are uploaded once -- nothing is computed on the CPU.  Accepts new-style ``GraphNode`` values
    (``.dependencies`` + ``__call__(values)``), legacy ``(func, *args)`` tuples, aliases and literal data.
    ``keys`` may be nested lists; the result has the same nesting.  ``to_host``: final results as NumPy."""
    if not _gpu():
        raise RuntimeError("dask_array_b200.get needs a CUDA device (B200); there is no CPU fallback")
    graph = dict(dsk)
    cache = {}

    def hashable(k):
        try:
            hash(k)
            return True
        except TypeError:
            return False

    def leafify(v):
        if isinstance(v, np.ndarray) and v.ndim >= 1 and v.dtype.kind in "biuf" and v.size:
            return to_device(v)
        return v

    def resolve(arg):
        """Legacy argument: a key, a nested list of keys, a nested task, or a literal."""
        if isinstance(arg, list):
            return [resolve(a) for a in arg]
        if _is_legacy_task(arg):
            return arg[0](*[resolve(a) for a in arg[1:]])
        if hashable(arg) and arg in graph:
            return value(arg)
        return arg

    def value(k):
        if k in cache:
            return cache[k]
        stack = [k]
        while stack:                                     # iterative DFS: deep graphs must not hit the recursion limit
            cur = stack[-1]
            if cur in cache:
                stack.pop()
                continue
            node = graph[cur]
            deps = getattr(node, "dependencies", None)
            if deps is not None and callable(node):     # dask._task_spec.GraphNode
                missing = [d for d in deps if d not in cache]
                if missing:
                    stack.extend(missing)
                    continue
                res = node({d: cache[d] for d in deps})
                leaf = not deps
            elif _is_legacy_task(node):
                missing = [d for d in _legacy_deps(node, graph) if d not in cache]
                if missing:
                    stack.extend(missing)
                    continue
                res = node[0](*[resolve(a) for a in node[1:]])
                leaf = not _legacy_deps(node, graph)
            elif hashable(node) and node in graph and node != cur:      # alias
                if node not in cache:
                    stack.append(node)
                    continue
                res, leaf = cache[node], False
            else:
                res, leaf = node, True
            cache[cur] = leafify(res) if leaf else res
            stack.pop()
        return cache[k]

    def fetch(ks):
        if isinstance(ks, list):
            return [fetch(k) for k in ks]
        out = value(ks)
        if to_host and isinstance(out, DeviceChunk):
            return out.to_numpy()
        return out

    return fetch(keys)


def _legacy_deps(task, graph) -> list:
    out = []

    def walk(a):
        if isinstance(a, list):
            for b in a:
                walk(b)
        elif _is_legacy_task(a):
            for b in a[1:]:
                walk(b)
        else:
            try:
                if a in graph:
                    out.append(a)
            except TypeError:
                pass
    for a in task[1:]:
        walk(a)
    return out
