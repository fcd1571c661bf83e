"""Tree reductions: ``chunk -> combine -> aggregate``.

Mirrors ``dask_array/reductions/_reduction.py`` (``Reduction`` :25, ``_lower`` :154-226,
``_normalize_split_every`` :715-725, ``_build_tree_reduce_expr`` :751-806, ``PartialReduce``
:900-983), ``reductions/_common.py`` (which reducers exist and their dtypes) and
``reductions/_arg_reduction.py`` (``ArgChunk`` :16-86).

Device representation of the per-block partials (the reference's dicts / structured arrays):
  sum/prod/min/max/any/all   one DeviceChunk with the kept-dims shape
  mean                       "total" DeviceChunk + the element count (uniform per block,
                             ``mean_chunk`` :270-281 collapses it the same way)
  var/std (moment)           one fp64 DeviceChunk (..., 3) = (n, mean, M2)
  argmin/argmax              "vals" + int64 "arg" DeviceChunks (the dict escape hatch of
                             ``arg_chunk`` :724-728 for chunk types without structured dtypes)
"""
from __future__ import annotations

import math
from numbers import Integral

import numpy as np

from . import _lib
from ._expr import ArrayExpr

REDOPS = {
    "sum": _lib.RED_SUM, "prod": _lib.RED_PROD, "min": _lib.RED_MIN, "max": _lib.RED_MAX,
    "any": _lib.RED_ANY, "all": _lib.RED_ALL, "mean": _lib.RED_SUM, "var": _lib.RED_MOMENT,
    "argmin": _lib.RED_ARGMIN, "argmax": _lib.RED_ARGMAX, "nanmin": _lib.RED_NANMIN, "nanmax": _lib.RED_NANMAX,
}


def normalize_split_every(split_every, axis):
    """``_normalize_split_every`` (``_reduction.py:715-725``); config default 16."""
    split_every = split_every or 16
    if isinstance(split_every, dict):
        return {k: split_every.get(k, 2) for k in axis}
    if isinstance(split_every, Integral):
        n = max(int(split_every ** (1 / (len(axis) or 1))), 2)
        return dict.fromkeys(axis, n)
    raise ValueError("split_every must be a int or a dict")


def validate_axis(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    if isinstance(axis, Integral):
        axis = (axis,)
    out = []
    for a in axis:
        if not -ndim <= a < ndim:
            raise np.exceptions.AxisError(a, ndim)
        out.append(a % ndim)
    if len(set(out)) != len(out):
        raise ValueError("duplicate value in 'axis'")
    return tuple(sorted(out))


def result_dtype(kind, in_dtype, dtype=None) -> np.dtype:
    """Output dtypes of ``reductions/_common.py``: sum/prod follow ``np.sum`` (:59-60), mean
    ``np.mean`` (:325-330), var ``np.var`` (:573-576), min/max the input, any/all bool,
    arg reductions intp."""
    in_dtype = np.dtype(in_dtype)
    if kind in ("sum", "prod"):
        return np.dtype(dtype) if dtype is not None else getattr(np, kind)(np.zeros(1, dtype=in_dtype)).dtype
    if kind == "mean":
        return np.dtype(dtype) if dtype is not None else np.mean(np.zeros((1,), dtype=in_dtype)).dtype
    if kind == "var":
        return np.dtype(dtype) if dtype is not None else np.var(np.ones((1,), dtype=in_dtype)).dtype
    if kind in ("min", "max", "nanmin", "nanmax"):
        return in_dtype
    if kind in ("any", "all"):
        return np.dtype(bool)
    if kind in ("argmin", "argmax"):
        return np.dtype(np.intp)
    raise NotImplementedError(kind)


class Reduction(ArrayExpr):
    """User-level reduction node (``Reduction`` :25; typed subclasses :461-712)."""

    _parameters = ["array", "kind", "axis", "keepdims", "dtype_", "split_every", "ddof"]
    _defaults = {"keepdims": False, "dtype_": None, "split_every": None, "ddof": 0}

    @property
    def dtype(self):
        return result_dtype(self.operand("kind"), self.operand("array").dtype, self.operand("dtype_"))

    @property
    def chunks(self):
        x, axis = self.operand("array"), self.operand("axis")
        if self.operand("keepdims"):
            return tuple((1,) if d in axis else c for d, c in enumerate(x.chunks))
        return tuple(c for d, c in enumerate(x.chunks) if d not in axis)

    def _tree_label(self):
        return f"{self.operand('kind').capitalize()}(axis={self.operand('axis')})"

    def _simplify_down(self):
        """A reduction over the window axis of a sliding-window view runs on the input's own chunks
        (``SlidingWindowReduction``; the reference's parent rewrite, ``_overlap.py:500-566``)."""
        from ._window import SlidingWindowView, native_window_reduction

        x = self.operand("array")
        if isinstance(x, SlidingWindowView):
            return native_window_reduction(x, self.operand("kind"), self.operand("axis"), self.operand("keepdims"), self.dtype)
        return None

    def _lower(self):
        """``Reduction._lower`` (:154-226) / ``arg_reduction`` (``_arg_reduction.py:89-150``)."""
        x, kind, axis = self.operand("array"), self.operand("kind"), self.operand("axis")
        if kind in ("argmin", "argmax"):
            ravel = len(axis) == x.ndim
            if len(axis) > 1 and not ravel:
                raise TypeError("axis must be either `None` or int for arg reductions")
            tmp = ArgChunk(x, kind, axis, ravel or x.ndim == 1)
        else:
            tmp = ChunkReduce(x, kind, axis, self.dtype)
        return build_tree_reduce(tmp, kind, axis, self.operand("keepdims"), self.dtype,
                                 self.operand("split_every"), self.operand("ddof"))


class ChunkReduce(ArrayExpr):
    """The per-block chunk step: a Blockwise with ``keepdims=True`` and
    ``adjust_chunks={axis: 1}`` (``_reduction.py:181-190``).  Fusable, so the element-wise
    chain that feeds it ends up in the same kernel (``test_general_reduction_names``)."""

    _parameters = ["array", "kind", "axis", "dtype_"]
    _is_blockwise_fusable = True

    @property
    def chunks(self):
        axis = self.operand("axis")
        return tuple(tuple(1 for _ in c) if d in axis else c for d, c in enumerate(self.operand("array").chunks))

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    def _tree_label(self):
        return f"{self.operand('kind')}_chunk(axis={self.operand('axis')})"


class ArgChunk(ArrayExpr):
    """``ArgChunk`` (``_arg_reduction.py:16-86``): NOT a Blockwise, hence never fused with what
    produces its input; per-block offsets (:72-76) become ``arg_offset`` / ravel bookkeeping."""

    _parameters = ["array", "kind", "axis", "ravel"]

    @property
    def chunks(self):
        axis = self.operand("axis")
        return tuple(tuple(1 for _ in c) if d in axis else c for d, c in enumerate(self.operand("array").chunks))

    @property
    def dtype(self):
        return np.dtype(np.intp)

    def _tree_label(self):
        return f"ArgChunk({self.operand('kind')}, axis={self.operand('axis')})"


class CumReduction(ArrayExpr):
    """``CumReduction`` (``reductions/_cumulative.py:100-265``; ``cumsum`` :451, ``cumprod`` :487,
    ``nancumsum`` :523, ``nancumprod`` :560): per-block ``np.cumsum`` plus the running total of the
    preceding blocks along ``axis``.  Not a Blockwise (never fused).  ``kind``: "cumsum" | "cumprod";
    ``nan``: NaNs count as the identity (the ``chunk.nancumsum`` variants)."""

    _parameters = ["array", "kind", "axis", "dtype_", "nan"]

    @property
    def chunks(self):
        return self.operand("array").chunks

    @property
    def dtype(self):
        if self.operand("dtype_") is not None:
            return np.dtype(self.operand("dtype_"))
        fn = np.cumsum if self.operand("kind") == "cumsum" else np.cumprod
        return fn(np.ones((0,), dtype=self.operand("array").dtype)).dtype        # :115-119

    def _tree_label(self):
        return f"CumReduction({'nan' if self.operand('nan') else ''}{self.operand('kind')}, axis={self.operand('axis')})"


class PartialReduce(ArrayExpr):
    """One level of the tree (``PartialReduce`` :900-983): every output block folds a group of
    ``split_every`` input partials in ``lol_tuples`` order."""

    _parameters = ["array", "kind", "axis", "split_every", "keepdims", "dtype_", "final", "ddof"]

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    @property
    def chunks(self):
        x, se = self.operand("array"), self.operand("split_every")
        ch = [tuple(1 for _ in range(-(-len(c) // se[d]))) if d in se else c for d, c in enumerate(x.chunks)]
        if not self.operand("keepdims"):
            ch = [c for d, c in enumerate(ch) if d not in se]
        return tuple(ch)

    def groups(self):
        """[(output key, [input block ids in nesting order])] -- ``_layer`` (:968-983)."""
        import itertools

        x, se = self.operand("array"), self.operand("split_every")
        parts = [[tuple(range(i, min(i + se.get(d, 1), n))) for i in range(0, n, se.get(d, 1))]
                 for d, n in enumerate(x.numblocks)]
        kept = [d for d in range(x.ndim) if d not in se]
        out = []
        for k in itertools.product(*[range(len(p)) for p in parts]):
            p = [parts[d][i] for d, i in enumerate(k)]
            members = list(itertools.product(*p))      # axis-major == lol_tuples nesting order
            key = k if self.operand("keepdims") else tuple(k[d] for d in kept)
            out.append((key, members))
        return out

    def _tree_label(self):
        tag = "aggregate" if self.operand("final") else "partial"
        return f"PartialReduce({self.operand('kind')}-{tag}, split_every={self.operand('split_every')})"


def build_tree_reduce(x, kind, axis, keepdims, dtype, split_every, ddof=0):
    """``_build_tree_reduce_expr`` (:751-806)."""
    se = normalize_split_every(split_every, axis)
    depth = 1
    for d, n in enumerate(x.numblocks):
        if d in se and se[d] != 1 and n > 1:
            depth = int(max(depth, math.ceil(math.log(n, se[d]))))
    for _ in range(depth - 1):
        x = PartialReduce(x, kind, axis, se, True, dtype, False, ddof)
    return PartialReduce(x, kind, axis, se, keepdims, dtype, True, ddof)
