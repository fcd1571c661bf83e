"""Sliding-window views and reductions.

``sliding_window_view(x, w, axis)`` (``dask_array/_overlap.py:1365-1433``) is a lazy node here,
``SlidingWindowView``; on its own it lowers to the overlap pipeline + a zero-copy window view per block.
A reduction over its window axis -- ``sliding_window_view(x, w, axis).sum(axis=-1)``, the rolling
sum / mean / min / max / prod / any / all -- is rewritten to ``SlidingWindowReduction``
(``reductions/_sliding_window.py:405-560``, created by the same parent-rewrite in the reference,
``_overlap.py:500-566``): the windows are never materialised or re-read ``w`` times.  Execution:

  ``WindowHalo``   every output-emitting block with the ``w - 1`` elements that follow it along the sliding
                   axis, gathered from as many following blocks as needed (chunks may be SMALLER than the
                   window -- the case the reference's banded plan exists for) by the rechunk executor: one
                   tiled gather per device, peer stores across GPUs;
  ``WindowReduce`` one ``b2_window_reduce`` launch per block: two-direction segment scans (the reference's
                   suffix scan + prefix scan of ``_sliding_window_banded_reduce`` with segments of exactly
                   ``w``), O(1) operations per element for any window, 2 N bytes of DRAM traffic.
"""
from __future__ import annotations

import itertools
from numbers import Integral

import numpy as np

from ._expr import ArrayExpr

# reducer -> (b2 redop name, divide by the window afterwards, input mapping)
NATIVE_REDUCERS = {"sum": "sum", "prod": "prod", "min": "min", "max": "max", "any": "max", "all": "min", "mean": "sum",
                   "nansum": "sum", "nanprod": "prod"}
_KERNEL_DTYPES = ("float32", "float64", "int32", "int64", "uint8", "bool")


class SlidingWindowView(ArrayExpr):
    """Lazy ``sliding_window_view``: the window axes are appended after the array's own axes."""

    _parameters = ["array", "window_shape", "axes"]

    @property
    def dtype(self):
        return self.operand("array").dtype

    def _depths(self):
        x = self.operand("array")
        d = [0] * x.ndim
        for ax, w in zip(self.operand("axes"), self.operand("window_shape")):
            d[ax] += w - 1
        return d

    def _safe_chunks(self):
        from ._overlap import ensure_minimum_chunksize

        x = self.operand("array")
        return tuple(ensure_minimum_chunksize(d + 1, c) if d else tuple(c) for d, c in zip(self._depths(), x.chunks))

    @property
    def chunks(self):
        safe = self._safe_chunks()
        return tuple(c[:-1] + (c[-1] - d,) for d, c in zip(self._depths(), safe)) + tuple(
            (w,) for w in self.operand("window_shape"))

    def _lower(self):
        from ._collection import Array
        from ._overlap import map_blocks, overlap

        x = Array(self.operand("array")).rechunk(self._safe_chunks())
        over = overlap(x, depth={i: (0, d) for i, d in enumerate(self._depths())}, boundary="none")
        return map_blocks(np.lib.stride_tricks.sliding_window_view, over, tuple(self.operand("window_shape")),
                          tuple(self.operand("axes")), dtype=x.dtype, chunks=self.chunks).expr

    def _tree_label(self):
        return f"SlidingWindowView(window={self.operand('window_shape')}, axis={self.operand('axes')})"


def sliding_window_view(x, window_shape, axis=None, automatic_rechunk=True):
    """``sliding_window_view`` (``_overlap.py:1365-1433``).  ``automatic_rechunk`` only changes how the
    reference re-balances chunk sizes; here chunks are merged just enough to hold a window."""
    from ._collection import Array, asarray

    x = asarray(x)
    window_shape = tuple(window_shape) if np.iterable(window_shape) else (window_shape,)
    if any(w <= 0 for w in window_shape):
        raise ValueError("`window_shape` must contain values > 0")
    if axis is None:
        axis = tuple(range(x.ndim))
        if len(window_shape) != len(axis):
            raise ValueError(f"Since axis is `None`, must provide window_shape for all dimensions of `x`; got "
                             f"{len(window_shape)} window_shape elements and `x.ndim` is {x.ndim}.")
    else:
        axis = tuple(a % x.ndim for a in ((axis,) if isinstance(axis, Integral) else axis))
        if len(window_shape) != len(axis):
            raise ValueError(f"Must provide matching length window_shape and axis; got {len(window_shape)} "
                             f"window_shape elements and {len(axis)} axes elements.")
    for ax, w in zip(axis, window_shape):
        if x.shape[ax] < w:
            raise ValueError("window shape cannot be larger than input array shape")
    return Array(SlidingWindowView(x.expr, tuple(int(w) for w in window_shape), tuple(axis)))


def native_window_reduction(view: SlidingWindowView, kind, axis, keepdims, dtype):
    """The parent rewrite (``_overlap.py:500-566``): a reduction over exactly the window axis of a one-axis
    sliding-window view becomes ``SlidingWindowReduction`` when the kernel covers the reducer and dtype;
    otherwise None (the generic plan -- window views + ordinary reduction -- runs)."""
    x = view.operand("array")
    if len(view.operand("axes")) != 1 or tuple(axis) != (x.ndim,) or kind not in NATIVE_REDUCERS:
        return None
    dtype = np.dtype(dtype)
    work = np.dtype(bool) if kind in ("any", "all") else dtype
    if work.name not in _KERNEL_DTYPES or (kind in ("mean",) and work.kind != "f"):
        return None
    w = int(view.operand("window_shape")[0])
    if int(view.operand("axes")[0]) == x.ndim - 1 and 2 * ((8 * w - 1) | 1) * work.itemsize > 200 * 1024:
        return None        # along the contiguous axis one row of the tile (8 segments, twice) must fit shared memory
    return SlidingWindowReduction(x, int(view.operand("window_shape")[0]), int(view.operand("axes")[0]), x.ndim,
                                  bool(keepdims), kind, dtype.name)


class SlidingWindowReduction(ArrayExpr):
    """``SlidingWindowReduction`` (``reductions/_sliding_window.py:405-560``): output chunks = the input's, trimmed
    by ``window - 1`` at the end of the sliding axis (:431-446)."""

    _parameters = ["array", "window", "sliding_axis", "window_axis", "keepdims", "reducer", "dtype_"]

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    def _trimmed(self):
        x, ax, w = self.operand("array"), self.operand("sliding_axis"), self.operand("window")
        remaining = sum(x.chunks[ax]) - w + 1
        out = []
        for c in x.chunks[ax]:
            if remaining <= 0:
                break
            take = min(c, remaining)
            out.append(take)
            remaining -= take
        return tuple(out)

    @property
    def chunks(self):
        x = self.operand("array")
        ch = list(x.chunks)
        ch[self.operand("sliding_axis")] = self._trimmed()
        if self.operand("keepdims"):
            ch.insert(self.operand("window_axis"), (1,))
        return tuple(ch)

    def _lower(self):
        from ._blockwise import Elemwise

        x, kind = self.operand("array"), self.operand("reducer")
        work = np.dtype(bool) if kind in ("any", "all") else self.dtype
        src = x
        if kind in ("nansum", "nanprod") and x.dtype.kind == "f":
            ident = 0 if kind == "nansum" else 1
            src = Elemwise("where", (Elemwise("isnan", (src,), ()), ident, src), ())
        if kind in ("any", "all"):
            src = Elemwise("not_equal", (src, 0), ()) if src.dtype != np.dtype(bool) else src
        elif src.dtype != work:
            src = Elemwise("astype", (src,), (("dtype", work.name),))
        halo = WindowHalo(src, self.operand("sliding_axis"), self.operand("window") - 1, self._trimmed())
        return WindowReduce(halo, self.operand("window"), self.operand("sliding_axis"), NATIVE_REDUCERS[kind],
                            kind == "mean", self.operand("keepdims"), self.operand("window_axis"), self.dtype.name)

    def _tree_label(self):
        return f"SlidingWindowReduction({self.operand('reducer')}, window={self.operand('window')}, axis={self.operand('sliding_axis')})"


class WindowHalo(ArrayExpr):
    """Block i of the output-emitting blocks plus the ``depth`` elements that follow it along ``axis``."""

    _parameters = ["array", "axis", "depth", "emit"]        # emit: trimmed chunk lengths along the axis

    @property
    def dtype(self):
        return self.operand("array").dtype

    @property
    def chunks(self):
        x = self.operand("array")
        ch = list(x.chunks)
        ch[self.operand("axis")] = tuple(n + self.operand("depth") for n in self.operand("emit"))
        return tuple(ch)

    def pieces(self, new_bid):
        """[(source block id, source slices, destination slices)] of one output block (rechunk executor)."""
        x, ax, depth = self.operand("array"), self.operand("axis"), self.operand("depth")
        ch = x.chunks[ax]
        edges = np.concatenate([[0], np.cumsum(ch)])
        i = new_bid[ax]
        lo = int(edges[i])
        hi = lo + self.operand("emit")[i] + depth
        segs = []
        for j in range(i, len(ch)):
            a, b = max(lo, int(edges[j])), min(hi, int(edges[j + 1]))
            if a < b:
                segs.append((j, slice(a - int(edges[j]), b - int(edges[j])), slice(a - lo, b - lo)))
            if int(edges[j + 1]) >= hi:
                break
        out = []
        for j, ssl, dsl in segs:
            sb, s_sl, d_sl = list(new_bid), [], []
            sb[ax] = j
            for d in range(x.ndim):
                if d == ax:
                    s_sl.append(ssl)
                    d_sl.append(dsl)
                else:
                    n = x.chunks[d][new_bid[d]]
                    s_sl.append(slice(0, n))
                    d_sl.append(slice(0, n))
            out.append((tuple(sb), tuple(s_sl), tuple(d_sl)))
        return out

    def _tree_label(self):
        return f"WindowHalo(axis={self.operand('axis')}, depth={self.operand('depth')})"


class WindowReduce(ArrayExpr):
    """``b2_window_reduce`` per block of a ``WindowHalo``."""

    _parameters = ["array", "window", "axis", "redop", "mean", "keepdims", "window_axis", "dtype_"]

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    @property
    def chunks(self):
        x = self.operand("array")
        ch = list(x.chunks)
        ax = self.operand("axis")
        ch[ax] = tuple(n - (self.operand("window") - 1) for n in ch[ax])
        if self.operand("keepdims"):
            ch.insert(self.operand("window_axis"), (1,))
        return tuple(ch)

    def _tree_label(self):
        return f"WindowReduce({self.operand('redop')}{'/w' if self.operand('mean') else ''}, window={self.operand('window')})"


# ----------------------------------------------------------------------------- trailing windows (bottleneck move_*)
MOVING_REDUCERS = {"move_sum": "nansum", "move_mean": "nanmean", "move_min": "nanmin", "move_max": "nanmax"}


def moving_window(x, window, reducer="move_sum", min_count=None, axis=-1):
    """``MovingWindowReduction`` (``reductions/_sliding_window.py:183-246, 249-400``; bottleneck ``move_*`` semantics, the
    reference's rewrite of ``map_overlap(bottleneck.move_sum, depth={axis: (window - 1, 0)})``): output position ``t``
    reduces the TRAILING window ``[t - window + 1, t]`` clipped at the array's start, skipping NaNs; windows with
    fewer than ``min_count`` (default ``window``) valid values are NaN.  Built from the sliding-window kernel: the
    NaN-free values and the valid flags are padded with ``window - 1`` identities in front and window-reduced on the
    input's own chunks (two ``b2_window_reduce`` launches), then combined in one fused element-wise kernel."""
    from ._collection import asarray, elemwise, full
    from ._views import concatenate

    if reducer not in MOVING_REDUCERS:
        raise ValueError(f"unknown moving-window reducer {reducer!r}")
    x = asarray(x)
    window = int(window)
    if window < 1:
        raise ValueError("window must be >= 1")
    limit = window if min_count is None else int(min_count)
    if not 1 <= limit <= window:
        raise ValueError("min_count must be in [1, window]")
    axis = axis % x.ndim
    dt = x.dtype if x.dtype.kind == "f" else np.dtype(np.float64)
    xf = x.astype(dt)
    ident = {"move_sum": 0.0, "move_mean": 0.0, "move_min": np.inf, "move_max": -np.inf}[reducer]
    valid = elemwise("logical_not", elemwise("isnan", xf))
    vals = elemwise("where", valid, xf, dt.type(ident))
    cnt_in = valid.astype(dt)
    if window > 1:
        pad_shape = tuple(window - 1 if d == axis else n for d, n in enumerate(x.shape))
        pad_chunks = tuple((window - 1,) if d == axis else c for d, c in enumerate(x.chunks))
        vals = concatenate([full(pad_shape, ident, dtype=dt, chunks=pad_chunks), vals], axis=axis)
        cnt_in = concatenate([full(pad_shape, 0.0, dtype=dt, chunks=pad_chunks), cnt_in], axis=axis)
        # put the pad's length on the LAST block: the window reduction trims window - 1 there (:431-446), so the
        # result comes out on the input's own chunks ("same shape and chunks as the input", :253)
        along = x.chunks[axis][:-1] + (x.chunks[axis][-1] + window - 1,)
        if min(along) >= window - 1:
            target = tuple(along if d == axis else c for d, c in enumerate(x.chunks))
            vals, cnt_in = vals.rechunk(target), cnt_in.rechunk(target)
    op = {"move_sum": "sum", "move_mean": "sum", "move_min": "min", "move_max": "max"}[reducer]
    red = getattr(sliding_window_view(vals, window, axis=axis), op)(axis=-1)
    cnt = sliding_window_view(cnt_in, window, axis=axis).sum(axis=-1)
    if reducer == "move_mean":
        red = elemwise("true_divide", red, cnt)
    out = elemwise("where", elemwise("less", cnt, dt.type(limit)), dt.type(np.nan), red).astype(dt)
    # chunks shorter than the window were merged for the halo exchange: restore the input's block structure
    return out if out.chunks == x.chunks else out.rechunk(x.chunks)


def block_ids_of(expr):
    return itertools.product(*[range(len(c)) for c in expr.chunks])
