"""Top-k selection: ``topk`` / ``argtopk`` (``routines/_topk.py:14-80``; chunk functions ``chunk.topk``,
``topk_aggregate``, ``argtopk``, ``argtopk_aggregate`` ``_chunk.py:200-290``).

The reference keeps ``k`` unsorted candidates per block (``np.partition``), concatenates them along the
axis level by level (``reduction(..., chunk=topk, combine=topk, aggregate=topk_aggregate)``) and sorts at
the end.  Here every level is one launch of ``b2_topk_rows``: segments of each row are sorted in shared
memory (bitonic network on (value, index) pairs) and their best ``k`` entries -- already in final order --
form the next level's rows, until one segment remains.  Values are the reference's values; the order of
EQUAL values among the returned indices is unspecified in both.
"""
from __future__ import annotations

import numpy as np

from ._expr import ArrayExpr


class TopK(ArrayExpr):
    _parameters = ["array", "k", "axis", "arg"]

    @property
    def chunks(self):
        x = self.operand("array")
        ax = self.operand("axis")
        keep = min(abs(self.operand("k")), x.shape[ax])
        return tuple((keep,) if d == ax else c for d, c in enumerate(x.chunks))

    @property
    def dtype(self):
        return np.dtype(np.intp) if self.operand("arg") else self.operand("array").dtype

    def _tree_label(self):
        return f"{'ArgTopK' if self.operand('arg') else 'TopK'}(k={self.operand('k')}, axis={self.operand('axis')})"


def _topk(a, k, axis, arg):
    from ._collection import Array, asarray
    from ._reductions import validate_axis

    a = asarray(a)
    if not isinstance(k, (int, np.integer)) or k == 0:
        raise ValueError("k must be a non-zero integer")
    axis = validate_axis(axis, a.ndim)[0]
    return Array(TopK(a.expr, int(k), axis, bool(arg)))


def topk(a, k, axis=-1, split_every=None):
    """The ``k`` largest elements along ``axis``, largest first (``k < 0``: the ``-k`` smallest, smallest
    first).  ``split_every`` shapes the reference's tree only; the values do not depend on it."""
    return _topk(a, k, axis, False)


def argtopk(a, k, axis=-1, split_every=None):
    """Indices (along ``axis``) of the elements ``topk`` returns."""
    return _topk(a, k, axis, True)
