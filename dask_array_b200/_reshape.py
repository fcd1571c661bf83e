"""``reshape`` (``manipulation/_reshape.py``): a C-order reshape that only merges or splits neighbouring axes is a
re-blocking (one tiled gather, ``TasksRechunk``) after which the k-th input block, viewed with the k-th output block's
shape, IS the k-th output block (``ReshapeLowered._layer`` :431-443) -- a zero-copy view per block here.

``reshape_plan`` restates the chunk planner ``reshape_rechunk`` (:38-127): which blocks the input must be cut into,
and which blocks the output then has.  It is pinned to the reference's own function by randomised equivalence
(``tests/test_reshape_plan.py``) and by ``tests/golden/reshape.json``.
"""
from __future__ import annotations

import itertools
import math

from ._expr import ArrayExpr

_UNSUPPORTED = """
reshape only supports operations that merge or split existing dimensions evenly, e.g. for
x = ones((6, 5, 4), chunks=(3, 2, 2)):  x.reshape((3, 2, 5, 4)) splits 6 into 3 and 2, x.reshape((30, 4)) merges 6 and 5,
but x.reshape((4, 5, 6)) would split existing dimensions unevenly.  Reshape in several passes instead.
"""


def _split_blocks(blocks, parts):
    """Cut every block into about ``parts`` pieces (``expand_tuple`` :203-231): ``(7, 4)`` by 3 -> (2, 2, 3, 1, 1, 2)."""
    if parts == 1:
        return tuple(blocks)
    out = []
    for c in blocks:
        piece = max(c / parts, 1)
        left = c
        while left >= 2 * piece:
            out.append(int(piece))
            left -= int(piece)
        if left:
            out.append(left)
    return tuple(out)


def _merge_to_multiples(blocks, unit):
    """Merge neighbouring blocks until ``unit`` divides every block (``contract_tuple`` :234-254)."""
    out, carry = [], 0
    for c in blocks:
        c += carry
        whole, carry = divmod(c, unit)
        if whole:
            out.append(whole * unit)
    return tuple(out)


def _largest_block(chunks, lo, hi):
    return int(math.prod(max(chunks[a]) for a in range(lo, hi + 1)))


def _near_even(n, parts):
    size = math.ceil(n / parts)
    out = [size] * parts
    for k in range(size * parts - n):
        out[k] -= 1
    return out


def _cap_group_blocks(lo, hi, cap, chunks):
    """``_smooth_chunks`` (:142-200).  Axes ``lo..hi`` are about to be merged into (or were split from) one axis and
    ``lo+1..hi`` were made single-block, which can blow the block size up.  Split the first axis of the group that is
    not all-ones -- the only one that can be cut without breaking C order -- until the largest block is back at
    ``cap`` elements."""
    first = lo
    while True:
        biggest = _largest_block(chunks, first, hi)
        if cap == biggest:
            return chunks
        lo = first
        while all(c == 1 for c in chunks[lo]):
            lo += 1
        if lo > hi:
            return chunks
        blocks = chunks[lo]
        if len(blocks) == 1:
            n = blocks[0]
            cut = _near_even(n, min(math.ceil(biggest / cap), n))
            chunks[lo] = tuple(cut)
            if all(c == 1 for c in cut) and lo < hi:
                continue                      # this axis dissolved into ones: the next one may need cutting too
            return chunks
        others = biggest // max(blocks)
        cut = []
        for c in blocks:
            if c * others <= cap:
                cut.append(c)
            else:
                cut.extend(_near_even(c, math.ceil(c * others / cap)))
        chunks[lo] = tuple(cut)
        return chunks


def _products(chunks, lo, hi):
    return tuple(math.prod(combo) for combo in itertools.product(*chunks[lo:hi + 1]))


def reshape_plan(inshape, outshape, inchunks):
    """``(input blocks to re-block to, output blocks)`` for a C-order reshape ``inshape -> outshape``.  The axes are
    matched from the right: equal lengths carry their blocks over, length-1 axes appear / vanish, a run of input axes
    whose product is one output axis is MERGED (all but the first become single-block, the first is cut finer), one
    input axis that is the product of a run of output axes is SPLIT (its blocks are merged to multiples of the inner
    product).  Anything else -- an uneven split -- raises ``NotImplementedError`` like the reference."""
    new_in, new_out = [None] * len(inshape), [None] * len(outshape)
    i, o = len(inshape) - 1, len(outshape) - 1
    while i >= 0 or o >= 0:
        if i < 0 or o < 0:
            # one side is used up: what is left on the other are length-1 axes
            if i < 0:
                new_out[o] = (1,)
                o -= 1
            else:
                new_in[i] = (1,)
                i -= 1
            continue
        n_in, n_out = inshape[i], outshape[o]
        if n_in == n_out:
            new_in[i] = new_out[o] = inchunks[i]
            i, o = i - 1, o - 1
        elif n_in == 1:
            new_in[i] = (1,)
            i -= 1
        elif n_out == 1:
            new_out[o] = (1,)
            o -= 1
        elif n_in < n_out:
            lo = i - 1
            while lo >= 0 and math.prod(inshape[lo:i + 1]) < n_out:
                lo -= 1
            if math.prod(inshape[lo:i + 1]) != n_out:
                raise NotImplementedError(_UNSUPPORTED)
            if all(len(inchunks[a]) == inshape[a] for a in range(i)):
                # every earlier axis is cut into single elements: blocks only move, nothing is re-blocked
                for a in range(i + 1):
                    new_in[a] = inchunks[a]
                new_out[o] = inchunks[i] * math.prod(len(inchunks[a]) for a in range(lo, i))
            else:
                for a in range(lo + 1, i + 1):
                    new_in[a] = (inshape[a],)
                new_in[lo] = _split_blocks(inchunks[lo], math.prod(len(inchunks[a]) for a in range(lo + 1, i + 1)))
                new_in = _cap_group_blocks(lo, i, _largest_block(inchunks, lo, i), new_in)
                new_out[o] = _products(new_in, lo, i)
            o, i = o - 1, lo - 1
        else:
            lo = o - 1
            while lo >= 0 and math.prod(outshape[lo:o + 1]) < n_in:
                lo -= 1
            if math.prod(outshape[lo:o + 1]) != n_in:
                raise NotImplementedError(_UNSUPPORTED)
            inner = math.prod(outshape[lo + 1:o + 1])
            new_in[i] = _merge_to_multiples(inchunks[i], inner)
            for a in range(lo + 1, o + 1):
                new_out[a] = (outshape[a],)
            new_out[lo] = tuple(c // inner for c in new_in[i])
            new_out = _cap_group_blocks(lo, o, _largest_block(inchunks, i, i), new_out)
            new_in[i] = _products(new_out, lo, o)
            o, i = lo - 1, i - 1
    return tuple(new_in), tuple(new_out)


class Reshape(ArrayExpr):
    """``ReshapeLowered`` (:417-456): the input is already blocked so that input block k (C order over the block grid)
    holds exactly the elements of output block k; ``chunks_`` are the output blocks."""

    _parameters = ["array", "shape_", "chunks_"]

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return self.operand("array").dtype

    def source(self, bid):
        """The input block holding output block ``bid``: same rank in C order over the two block grids."""
        x = self.operand("array")
        rank = 0
        for k, n in zip(bid, self.numblocks):
            rank = rank * n + k
        out = []
        for n in reversed(x.numblocks):
            rank, k = divmod(rank, n)
            out.append(k)
        return tuple(reversed(out))

    def view(self, chunk, bid):
        from ._eager import copy

        return (chunk if chunk.is_contiguous else copy(chunk)).reshape(self.block_shape(bid))

    def _tree_label(self):
        return f"Reshape{tuple(self.operand('shape_'))}"


def reshape(x, shape, merge_chunks=True, limit=None):
    """``reshape`` (:460-522): ``-1`` is inferred, the identity returns ``x``, a one-block array is viewed directly,
    ``merge_chunks=False`` first cuts the leading axes that disappear into single elements (so blocks only move)."""
    from ._collection import Array, asarray

    x = asarray(x)
    shape = (shape,) if isinstance(shape, (int,)) or hasattr(shape, "__index__") else tuple(shape)
    shape = tuple(int(s) for s in shape)
    unknown = [k for k, s in enumerate(shape) if s == -1]
    if unknown:
        if len(unknown) > 1:
            raise ValueError("can only specify one unknown dimension")
        if len(shape) == 1 and x.ndim == 1:
            return Array(x.expr)
        known = math.prod(s for s in shape if s != -1)
        missing = x.size / known if known else 0
        if missing != int(missing):
            raise ValueError("total size of new array must be unchanged")
        shape = tuple(int(missing) if s == -1 else s for s in shape)
    if math.prod(shape) != x.size:
        raise ValueError("total size of new array must be unchanged")
    if x.shape == shape:
        return x
    expr = x.expr
    if math.prod(expr.numblocks) == 1:
        return Array(Reshape(expr, shape, tuple((n,) for n in shape)))
    if not merge_chunks and x.ndim > len(shape):
        expr = x.rechunk({a: 1 for a in range(x.ndim - len(shape))}).expr
    cut, blocks = reshape_plan(expr.shape, shape, expr.chunks)
    if cut != expr.chunks:
        from ._rechunk import Rechunk

        expr = Rechunk(expr, cut)
    return Array(Reshape(expr, shape, blocks))
