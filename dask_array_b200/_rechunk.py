"""Rechunk: pure data movement between two block grids.

Mirrors ``dask_array/_rechunk.py``: ``Rechunk`` (:646) with its no-op / double-rechunk
rules (:744-805), lowered to ``TasksRechunk`` (:1157-1323).  The reference's planner
(``plan_rechunk`` :442-516) inserts intermediate chunkings to bound the number of Python
tasks -- for BASELINE config 4 it goes (16384,256)x64 -> (16384,1024)x16 -> (256,16384)x64 --
and every stage is a full copy (``getitem`` + ``concatenate3``).  A kernel has no per-task
overhead to bound, so here EVERY rechunk is a single tiled gather old -> new
(``b2_gather_launch``): each new block is assembled from the intersecting rectangles of
the old blocks (``old_to_new`` / ``intersect_chunks`` :130-198).  Values are identical by
construction; traffic is the algorithmic 2 x N x itemsize instead of >= 4 x.
"""
from __future__ import annotations

import itertools

import numpy as np

from ._expr import ArrayExpr, normalize_chunks


def old_to_new(old_chunks, new_chunks):
    """Per dim, per new block: [(old block index, slice inside it)] (``_rechunk.py:130-175``)."""
    out = []
    for oc, nc in zip(old_chunks, new_chunks):
        edges = np.concatenate([[0], np.cumsum(oc)]).astype(np.int64)
        dim, pos = [], 0
        for n in nc:
            lo, hi = pos, pos + n
            first = int(np.searchsorted(edges, lo, side="right") - 1)
            pieces = []
            i = max(first, 0)
            while i < len(oc) and edges[i] < hi:
                a, b = max(lo, int(edges[i])), min(hi, int(edges[i + 1]))
                if a < b:
                    pieces.append((i, slice(a - int(edges[i]), b - int(edges[i]))))
                i += 1
            dim.append(pieces)
            pos = hi
        out.append(dim)
    return out


class Rechunk(ArrayExpr):
    _parameters = ["array", "chunks_"]

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return self.operand("array").dtype

    def _simplify_down(self):
        x = self.operand("array")
        if x.chunks == self.chunks:                        # no-op removal (:744)
            return x
        if isinstance(x, Rechunk):                         # rechunk(rechunk(x)) (:755)
            return Rechunk(x.operand("array"), self.chunks)
        return None

    def _lower(self):
        return TasksRechunk(self.operand("array"), self.chunks)

    def _tree_label(self):
        return f"Rechunk(chunks={_short(self.chunks)})"


class TasksRechunk(Rechunk):
    """The executable form (``TasksRechunk._layer`` :1171-1187): one gather launch."""

    def _simplify_down(self):
        return None

    def _lower(self):
        return None

    def pieces(self, new_bid):
        """[(old block id, source slices, destination slices)] of one new block."""
        o2n = self._cache.get("o2n")
        if o2n is None:
            o2n = self._cache["o2n"] = old_to_new(self.operand("array").chunks, self.chunks)
        per_dim = [o2n[d][i] for d, i in enumerate(new_bid)]
        starts = []
        for pcs in per_dim:
            acc, st = 0, []
            for _, s in pcs:
                st.append(acc)
                acc += s.stop - s.start
            starts.append(st)
        out = []
        for combo in itertools.product(*[range(len(p)) for p in per_dim]):
            obid = tuple(per_dim[d][k][0] for d, k in enumerate(combo))
            src = tuple(per_dim[d][k][1] for d, k in enumerate(combo))
            dst = tuple(slice(starts[d][k], starts[d][k] + (src[d].stop - src[d].start)) for d, k in enumerate(combo))
            out.append((obid, src, dst))
        return out

    def _tree_label(self):
        return f"TasksRechunk(chunks={_short(self.chunks)})"


def _short(chunks):
    return tuple((c[0],) if len(set(c)) == 1 else c for c in chunks)


def rechunk(x_expr, chunks):
    """``rechunk()`` (:1452): normalise the request ({axis: size} dicts, -1, ints)."""
    if isinstance(chunks, dict):
        chunks = tuple(chunks.get(d, x_expr.chunks[d]) for d in range(x_expr.ndim))
    return Rechunk(x_expr, normalize_chunks(chunks, x_expr.shape))
