"""Basic slicing (slices -- any step, either sign -- and integers).

Mirrors ``dask_array/slicing/_basic.py:357-493`` (``SliceSlicesIntegers``): every output block
is ``getitem(block, slices)`` of one input block, here a zero-copy strided view of the
device block.  Slices are pushed towards the leaves first (``_expr.py:399-468``,
``Elemwise._accept_slice`` ``_blockwise.py:1096``, ``Transpose._accept_slice``
``_transpose.py:168``, ``BroadcastTrick._accept_slice`` ``_ones_zeros.py:99``) so that
``(x + x.T)[:100, :100]`` computes 100 x 100 elements, as in the README example.
"""
from __future__ import annotations

import numpy as np

from ._blockwise import Elemwise, Transpose
from ._expr import ArrayExpr, BroadcastTrick, FromArray


def normalize_index(index, shape):
    if not isinstance(index, tuple):
        index = (index,)
    if any(i is None or i is Ellipsis for i in index):
        n_real = sum(1 for i in index if i is not None and i is not Ellipsis)
        if any(i is None for i in index):
            raise NotImplementedError("np.newaxis in a dask_array_b200 index")
        k = index.index(Ellipsis)
        index = index[:k] + (slice(None),) * (len(shape) - n_real) + index[k + 1:]
    index = index + (slice(None),) * (len(shape) - len(index))
    if len(index) != len(shape):
        raise IndexError("too many indices for array")
    out = []
    for ix, n in zip(index, shape):
        if isinstance(ix, slice):
            start, stop, step = ix.indices(n)
            if step == 1:
                out.append(slice(start, max(stop, start)))
            else:       # canonical stepped form: start/stop as ``slice.indices`` gives them (stop may be -1)
                out.append(slice(start, stop, step))
        elif isinstance(ix, (int, np.integer)):
            i = int(ix)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError(f"index {ix} is out of bounds for axis with size {n}")
            out.append(i)
        else:
            raise NotImplementedError(f"only basic slicing is on the B200 hot path, got {type(ix).__name__}")
    return tuple(out)


class SliceSlicesIntegers(ArrayExpr):
    _parameters = ["array", "index"]

    @property
    def dtype(self):
        return self.operand("array").dtype

    def _per_dim(self):
        """Per input dim: [(input block, slice or int inside it)] for the selected range."""
        if "per_dim" in self._cache:
            return self._cache["per_dim"]
        x = self.operand("array")
        res = []
        for ix, ch in zip(self.operand("index"), x.chunks):
            edges = np.concatenate([[0], np.cumsum(ch)])
            if isinstance(ix, int):
                b = int(np.searchsorted(edges, ix, side="right") - 1)
                res.append([(b, ix - int(edges[b]))])
                continue
            pcs = []
            step = ix.step or 1
            if step == 1:
                for b in range(len(ch)):
                    a, e = max(ix.start, int(edges[b])), min(ix.stop, int(edges[b + 1]))
                    if a < e:
                        pcs.append((b, slice(a - int(edges[b]), e - int(edges[b]))))
            elif step > 1:
                # ``_slice_1d`` (slicing/_utils.py:279-420): blocks that hold no selected element drop out
                for b in range(len(ch)):
                    lo, hi = int(edges[b]), min(int(edges[b + 1]), ix.stop)
                    k0 = max(0, -(-(lo - ix.start) // step))
                    first = ix.start + k0 * step
                    if first < hi:
                        pcs.append((b, slice(first - lo, hi - lo, step)))
            else:
                # negative step: the blocks come out in reverse order (``_slice_1d`` :411-440)
                for b in range(len(ch) - 1, -1, -1):
                    lo, hi = int(edges[b]), int(edges[b + 1])
                    k0 = max(0, -(-(ix.start - (hi - 1)) // -step))
                    first = ix.start + k0 * step
                    floor = max(ix.stop, lo - 1)                 # selected indices are > floor
                    if first > floor:
                        pcs.append((b, slice(first - lo, (floor - lo) if floor >= lo else None, step)))
            if not pcs:
                pcs = [(0, slice(0, 0))]
            res.append(pcs)
        self._cache["per_dim"] = res
        return res

    @property
    def chunks(self):
        out = []
        for ix, pcs, ch in zip(self.operand("index"), self._per_dim(), self.operand("array").chunks):
            if isinstance(ix, int):
                continue
            out.append(tuple(len(range(*sl.indices(ch[b]))) for b, sl in pcs))
        return tuple(out)

    def source(self, out_bid):
        """(input block id, index tuple) of one output block (``_layer`` :463)."""
        it = iter(out_bid)
        bid, idx = [], []
        for ix, pcs in zip(self.operand("index"), self._per_dim()):
            b, s = pcs[0] if isinstance(ix, int) else pcs[next(it)]
            bid.append(b)
            idx.append(s)
        return tuple(bid), tuple(idx)

    def _simplify_down(self):
        x, index = self.operand("array"), self.operand("index")
        if all(isinstance(i, slice) and i.start == 0 and i.stop == n and (i.step or 1) == 1
               for i, n in zip(index, x.shape)):
            return x
        # the pushdown rules below are written for unit-step windows; stepped slices stay where they are
        # (still zero-copy strided views of the blocks)
        only_slices = all(isinstance(i, slice) and (i.step or 1) == 1 for i in index)
        if isinstance(x, BroadcastTrick) and only_slices:
            shape = tuple(i.stop - i.start for i in index)
            # keep the original chunk size, clipped (``_ones_zeros.py:99-121``)
            chunks = tuple(_clip_chunks(c, i) for c, i in zip(x.chunks, index))
            return BroadcastTrick(x.operand("value"), shape, chunks, x.dtype)
        if isinstance(x, Transpose) and only_slices:
            axes = x.operand("axes")
            inner = [None] * len(axes)
            for out_d, a in enumerate(axes):
                inner[a] = index[out_d]
            return Transpose(SliceSlicesIntegers(x.operand("array"), tuple(inner)), axes)
        if isinstance(x, Elemwise) and only_slices:
            nd = x.ndim

            def push(a):
                off = nd - a.ndim
                # slice(0, 1) only where the operand is BROADCAST along the dim; a genuine size-1 dim of the
                # output takes the real index (it may be empty: (a + b)[0:0])
                sub = tuple(slice(0, 1) if a.shape[d] == 1 and x.shape[off + d] != 1 else index[off + d]
                            for d in range(a.ndim))
                return SliceSlicesIntegers(a, sub)
            return x._map_args(push)
        if isinstance(x, SliceSlicesIntegers) and only_slices and all(
                isinstance(i, slice) and (i.step or 1) == 1 for i in x.operand("index")):
            inner = x.operand("index")
            merged = tuple(slice(a.start + b.start, a.start + b.stop) for a, b in zip(inner, index))
            return SliceSlicesIntegers(x.operand("array"), merged)
        if isinstance(x, FromArray) and only_slices:
            arr = x.operand("array")[index]
            chunks = tuple(_clip_chunks(c, i) for c, i in zip(x.chunks, index))
            return FromArray(arr, chunks)
        return None

    def _tree_label(self):
        return f"Slice{self.operand('index')}"


def _clip_chunks(chunks, sl):
    edges = np.concatenate([[0], np.cumsum(chunks)])
    out = [int(min(sl.stop, hi) - max(sl.start, lo)) for lo, hi in zip(edges[:-1], edges[1:])]
    return tuple(c for c in out if c > 0) or (0,)
