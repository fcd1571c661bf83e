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
        return None

    def _lower(self):
        return None                                        # deferred: lower-inserted rechunks may still push down

    def _lower_rechunk(self):
        return TasksRechunk(self.operand("array"), self.chunks)

    def _tree_label(self):
        return f"Rechunk(chunks={_short(self.chunks)})"


class TasksRechunk(Rechunk):
    """The executable form (``TasksRechunk._layer`` :1171-1187): one gather launch."""

    def _simplify_down(self):
        return None

    def _lower(self):
        return None

    def _lower_rechunk(self):
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


def balance_chunksizes(chunks):
    """``_balance_chunksizes`` (:529-560): among the regular blockings whose edge lies within half a median of the
    median block length and that give the same number of blocks (one fewer when the smallest block is at most half
    the largest), take the one with the smallest spread; unchanged when none exists or a block is empty."""
    if min(chunks) == 0:
        return tuple(chunks)
    total = sum(chunks)
    median = int(np.median(chunks))
    want = len(chunks) - (1 if min(chunks) <= 0.5 * max(chunks) else 0)
    best = None
    for edge in range(median - median // 2, median + median // 2 + 1):
        full, rest = divmod(total, edge)
        cand = (edge,) * full + ((rest,) if rest else ())
        if len(cand) == want and (best is None or max(cand) - min(cand) < max(best) - min(best)):
            best = cand
    if best is None:
        import warnings

        warnings.warn("chunk size balancing not possible with given chunks. Try increasing the chunk size.")
        return tuple(chunks)
    return best


def rechunk(x_expr, chunks="auto", block_size_limit=None, balance=False):
    """``rechunk()`` (:1452) / ``Rechunk.chunks`` (:690-718): ``{axis: size}`` dicts (negative axes, missing / None
    entries keep the current blocks), None entries in tuples, -1, ints, explicit blocks, and "auto" / byte sizes,
    which scale the CURRENT blocks (``previous_chunks=x.chunks``) up to ``block_size_limit`` (default 128 MiB)."""
    nd = x_expr.ndim
    if isinstance(chunks, dict):
        given = {}
        for ax, v in chunks.items():
            if not -nd <= ax < nd:
                raise ValueError(f"axis {ax} is out of bounds for array of dimension {nd}")
            given[ax % nd] = v
        chunks = tuple(x_expr.chunks[d] if given.get(d) is None else given[d] for d in range(nd))
    if isinstance(chunks, (tuple, list)):
        chunks = tuple(cur if c is None else c for c, cur in zip(chunks, x_expr.chunks)) if len(chunks) == nd else tuple(chunks)
    new = normalize_chunks(chunks, x_expr.shape, dtype=x_expr.dtype, limit=block_size_limit,
                           previous_chunks=x_expr.chunks)
    if len(new) != nd:
        raise ValueError("Provided chunks are not consistent with shape")
    if balance:
        new = tuple(balance_chunksizes(c) for c in new)
    return Rechunk(x_expr, new)


# ----------------------------------------------------------------------------- pushdown
def _parent_counts(root) -> dict:
    """How many distinct parents consume each node (the ``dependents`` the reference's
    ``_simplify_up`` receives, ``_rechunk.py:104-108``)."""
    counts, seen, stack = {root._name: 0}, set(), [root]
    while stack:
        node = stack.pop()
        if node._name in seen:
            continue
        seen.add(node._name)
        for name in {d._name: d for d in node.dependencies()}:
            counts[name] = counts.get(name, 0) + 1
        stack.extend(node.dependencies())
    return counts


def _push(node: "Rechunk"):
    """``Rechunk._pushdown`` (``_rechunk.py:755-806``) for the children of the hot path; None = keep.
    On the GPU a pushed rechunk that reaches a host leaf costs nothing (the upload simply cuts the
    host array at the new grid) and one that reaches an element-wise chain lets the chain fuse with
    whatever consumes the rechunked result."""
    from ._blockwise import Elemwise, Transpose
    from ._expr import BroadcastTrick, FromArray

    child, target = node.operand("array"), node.chunks
    if type(child) is Rechunk:                                              # :776-784
        return Rechunk(child.operand("array"), target)
    if isinstance(child, Transpose):                                        # :787 / _pushdown_through_transpose
        axes = tuple(child.operand("axes"))
        inner = [None] * len(axes)
        for out_dim, in_dim in enumerate(axes):
            inner[in_dim] = target[out_dim]
        return Transpose(Rechunk(child.operand("array"), tuple(inner)), axes)
    if isinstance(child, Elemwise):                                         # :795 / _pushdown_through_elemwise :970-1032
        nd = child.ndim

        def fix(a):
            off = nd - a.ndim
            new = tuple((1,) if a.shape[d] == 1 else target[d + off] for d in range(a.ndim))
            return a if new == a.chunks else Rechunk(a, new)
        return child._map_args(fix)
    if type(child) is FromArray:                                            # _pushdown_into_io :808-823
        return FromArray(child.operand("array"), target)
    if type(child) is BroadcastTrick:
        return BroadcastTrick(child.operand("value"), child.operand("shape_"), target, child.operand("dtype_"))
    return None


def pushdown_rechunks(root):
    """Push every ``Rechunk`` as far towards the leaves as the reference does, never into a node that
    something else also consumes (``tests/test_rechunk_pushdown.py:605-680``: a pushed copy would
    re-derive the shared node -- a second read, a duplicated chain)."""
    for _ in range(256):
        counts = _parent_counts(root)
        memo, changed = {}, False

        def visit(node):
            nonlocal changed
            if node._name in memo:
                return memo[node._name]
            out = node
            if type(node) is Rechunk and not changed:
                child = node.operand("array")
                if child.chunks == node.chunks:
                    out, changed = child, True
                elif counts.get(child._name, 0) <= 1:
                    pushed = _push(node)
                    if pushed is not None:
                        out, changed = pushed, True
            if out is node:
                out = node.map_children(visit)
            memo[node._name] = out
            return out

        root = visit(root)
        if not changed:
            return root
    return root
