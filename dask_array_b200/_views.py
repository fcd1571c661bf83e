"""Structural, zero-copy expressions: ``ExpandDims`` / ``Squeeze`` (``manipulation/_expand.py:50-64``),
``BroadcastTo`` (``_broadcast_to.py:37-52``), ``Concatenate`` / ``Stack`` (``stacking/``).  None of
them moves data: every output block is a stride view of one input block (size-1 dims inserted or
dropped, stride-0 broadcast) or simply an input block under a new block id (concatenation along a
chunk boundary).  The kernels downstream read the views through the block descriptors.
"""
from __future__ import annotations

import numpy as np

from ._expr import ArrayExpr, normalize_chunks


class ExpandDims(ArrayExpr):
    _parameters = ["array", "axis"]

    @property
    def chunks(self):
        c = list(self.operand("array").chunks)
        c.insert(self.operand("axis"), (1,))
        return tuple(c)

    @property
    def dtype(self):
        return self.operand("array").dtype

    def source(self, bid):
        ax = self.operand("axis")
        return tuple(b for d, b in enumerate(bid) if d != ax)

    def view(self, chunk):
        ax = self.operand("axis")
        return chunk[tuple(slice(None) for _ in range(ax)) + (None,)]


class Ravel(ArrayExpr):
    """C-order flatten of an array whose trailing axes are single-chunk (``reshape(-1)`` after the rechunk the
    reference's ``reshape`` / ``_prepare_cumulative`` performs, ``reductions/_cumulative.py:77-97``): every block
    (c0, n1, n2, ...) is contiguous in the flattened order, so the output block is a reshaped VIEW of it."""

    _parameters = ["array"]

    @property
    def chunks(self):
        x = self.operand("array")
        if any(len(c) != 1 for c in x.chunks[1:]):
            raise ValueError("Ravel needs single-chunk trailing axes (rechunk first)")
        inner = 1
        for n in x.shape[1:]:
            inner *= n
        return (tuple(c * inner for c in x.chunks[0]),) if x.ndim else ((1,),)

    @property
    def dtype(self):
        return self.operand("array").dtype

    def source(self, bid):
        x = self.operand("array")
        return (bid[0],) + (0,) * (x.ndim - 1) if x.ndim else ()

    def view(self, chunk):
        from ._eager import copy

        return (chunk if chunk.is_contiguous else copy(chunk)).reshape((chunk.size,))


def ravel(x):
    """``Array.ravel`` / ``flatten`` for the layouts the hot path needs: the trailing axes are made single-chunk by a
    rechunk (one tiled gather), then every block is viewed flat."""
    from ._collection import Array, asarray

    x = asarray(x)
    if x.ndim == 1:
        return x
    if x.ndim == 0:
        return x[None] if hasattr(x, "__getitem__") else x
    if any(len(c) != 1 for c in x.chunks[1:]):
        x = x.rechunk((x.chunks[0],) + tuple((n,) for n in x.shape[1:]))
    return Array(Ravel(x.expr))


class Squeeze(ArrayExpr):
    _parameters = ["array", "axes"]

    @property
    def chunks(self):
        ax = self.operand("axes")
        return tuple(c for d, c in enumerate(self.operand("array").chunks) if d not in ax)

    @property
    def dtype(self):
        return self.operand("array").dtype

    def source(self, bid):
        ax, it = self.operand("axes"), iter(bid)
        return tuple(0 if d in ax else next(it) for d in range(self.operand("array").ndim))

    def view(self, chunk):
        ax = self.operand("axes")
        return chunk[tuple(0 if d in ax else slice(None) for d in range(chunk.ndim))]


class BroadcastTo(ArrayExpr):
    _parameters = ["array", "shape_", "chunks_"]

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return self.operand("array").dtype

    def source(self, bid):
        x = self.operand("array")
        off = self.ndim - x.ndim
        return tuple(bid[off + d] if x.numblocks[d] > 1 else 0 for d in range(x.ndim))

    def view(self, chunk, bid):
        return chunk.broadcast_to(self.block_shape(bid))


class Concatenate(ArrayExpr):
    """Inputs must agree on the chunks of every other axis (rechunk first otherwise)."""

    _parameters = ["arrays", "axis"]

    def dependencies(self):
        return list(self.operand("arrays"))

    def map_children(self, fn):
        new = tuple(fn(a) for a in self.operand("arrays"))
        if all(a is b for a, b in zip(new, self.operand("arrays"))):
            return self
        return Concatenate(new, self.operand("axis"))

    @property
    def chunks(self):
        arrs, ax = self.operand("arrays"), self.operand("axis")
        first = arrs[0].chunks
        cat = tuple(c for a in arrs for c in a.chunks[ax])
        return tuple(cat if d == ax else first[d] for d in range(len(first)))

    @property
    def dtype(self):
        return np.result_type(*[a.dtype for a in self.operand("arrays")])

    def source(self, bid):
        """(input index, input block id)"""
        ax, i = self.operand("axis"), bid[self.operand("axis")]
        for k, a in enumerate(self.operand("arrays")):
            n = a.numblocks[ax]
            if i < n:
                return k, tuple(i if d == ax else b for d, b in enumerate(bid))
            i -= n
        raise IndexError(bid)


def expand_dims(a, axis):
    from ._collection import Array

    axis = axis if axis >= 0 else axis + a.ndim + 1
    return Array(ExpandDims(a.expr, axis))


def squeeze(a, axis=None):
    from ._collection import Array

    if axis is None:
        axes = tuple(d for d, n in enumerate(a.shape) if n == 1)
    else:
        axes = tuple(ax % a.ndim for ax in ((axis,) if np.isscalar(axis) else axis))
    for ax in axes:
        if a.shape[ax] != 1:
            raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
    return Array(Squeeze(a.expr, axes)) if axes else a


def broadcast_to(a, shape, chunks=None):
    from ._collection import Array

    shape = tuple(int(s) for s in shape)
    off = len(shape) - a.ndim
    if off < 0 or any(n != 1 and n != m for n, m in zip(a.shape, shape[off:])):
        raise ValueError(f"cannot broadcast shape {a.shape} to shape {shape}")
    if chunks is None:
        chunks = tuple(shape[:off]) + tuple(c if n != 1 else (shape[off + d],) for d, (c, n) in enumerate(zip(a.chunks, a.shape)))
    return Array(BroadcastTo(a.expr, shape, normalize_chunks(chunks, shape)))


def concatenate(arrays, axis=0):
    from ._collection import Array, asarray
    from ._rechunk import Rechunk

    arrays = [asarray(a) for a in arrays]
    nd = arrays[0].ndim
    axis %= nd
    exprs = [a.expr for a in arrays]
    dt = np.result_type(*[a.dtype for a in arrays])
    fixed = []
    for e in exprs:
        if e.dtype != dt:
            e = Array(e).astype(dt).expr
        want = tuple(e.chunks[d] if d == axis else exprs[0].chunks[d] for d in range(nd))
        if tuple(e.shape[d] for d in range(nd) if d != axis) != tuple(exprs[0].shape[d] for d in range(nd) if d != axis):
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
        fixed.append(e if e.chunks == want else Rechunk(e, want))
    return Array(Concatenate(tuple(fixed), axis))


def stack(arrays, axis=0):
    arrays = [expand_dims(a, axis if axis >= 0 else axis + a.ndim + 1) for a in arrays]
    return concatenate(arrays, axis=axis if axis >= 0 else axis + arrays[0].ndim)
