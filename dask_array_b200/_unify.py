"""Chunk unification of element-wise operands -- the step in front of every fused chain whose
operands disagree on their block grid (``unify_chunks_expr``, ``_expr.py:723-905``; helpers
``coarse_blockdim`` :586-660, ``moved_fraction`` :672-720, ``common_blockdim``
``_core_utils.py:893-960``; defaults ``dask_array/__init__.py:14-29``).

Same decisions as the reference, axis by axis:

1. *nest*: when every operand's boundaries contain the boundaries of the operand with the fewest
   blocks, everybody is merged up to that coarsest grid -- unless merging would move more than
   ``MERGE_COST_RATIO`` x the bytes that already sit on that grid (a light, coarse operand must not
   inflate a heavy, fine one); then the axis is *refined* (cut at the union of all boundaries);
2. *interleaved* grids (no nesting) are refined, then -- if some operand's existing grid can adopt the
   others for a proportionate number of moved bytes -- *realigned* to the cheapest such grid with the
   fewest blocks;
3. a merge that would manufacture a chunk above ``CHUNK_LIMIT`` bytes refines the merged axes instead.

On the GPU a "rechunk" of an operand is one tiled gather launch (``_rechunk.py``) -- or the same
launch storing into peer memory when the operand is spread over several GPUs -- so the policy's
notion of "bytes moved" is exactly the traffic of that launch.
"""
from __future__ import annotations

import math
from itertools import accumulate

MERGE_COST_RATIO = 4
CHUNK_LIMIT = 512 * 2**20


def _cuts(layout) -> set:
    """Interior boundaries of a chunking."""
    return set(accumulate(layout[:-1]))


def refine(layouts) -> tuple:
    """Cut one axis at the union of all boundaries (the reference's ``common_blockdim``)."""
    layouts = {tuple(c) for c in layouts}
    if not any(layouts):
        return ()
    chunked = [c for c in layouts if len(c) > 1]
    if not chunked:
        return max(layouts, key=lambda c: c[0])
    if len(chunked) == 1:
        return chunked[0]
    totals = {sum(c) for c in chunked}
    if len(totals) > 1:
        raise ValueError("Chunks do not add up to same value", layouts)
    edges = sorted(set().union(*(_cuts(c) for c in chunked)) | {0, totals.pop()})
    return tuple(b - a for a, b in zip(edges, edges[1:]))


def coarsest_nested(layouts) -> tuple:
    """The grid with the fewest blocks if all others nest inside it, else ``refine``
    (the reference's ``coarse_blockdim``)."""
    layouts = {tuple(c) for c in layouts}
    if not any(layouts):
        return ()
    chunked = [c for c in layouts if len(c) > 1]
    if not chunked:
        return max(layouts, key=lambda c: c[0])
    if len(chunked) == 1:
        return chunked[0]
    if len({sum(c) for c in chunked}) > 1:
        raise ValueError("Chunks do not add up to same value", layouts)
    top = min(chunked, key=len)
    need = _cuts(top)
    if all(need <= _cuts(c) for c in chunked):
        return top
    return refine(layouts)


def moved_share(src, dst) -> float:
    """Share of an axis a rechunk ``src -> dst`` moves: each new chunk is built where its largest old
    piece already lives, only the rest travels (the reference's ``moved_fraction``)."""
    src, dst = tuple(src), tuple(dst)
    total = sum(src)
    if not total or src == dst or sum(dst) != total:
        return 0.0
    moved = 0.0
    k, lo_s, lo_d = 0, 0.0, 0.0
    for width in dst:
        hi_d = lo_d + width
        keep = 0.0
        while True:
            hi_s = lo_s + src[k]
            keep = max(keep, min(hi_s, hi_d) - max(lo_s, lo_d))
            if hi_s > hi_d or k + 1 == len(src):
                break
            k, lo_s = k + 1, hi_s
        moved += width - keep
        lo_d = hi_d
    return moved / total


def unify(operands, policy: str = "auto", limit: int | None = CHUNK_LIMIT):
    """``operands``: [(shape, chunks, itemsize)] of the array operands of an element-wise op, NumPy
    right-aligned.  Returns ``(out_chunks, [target chunks per operand])`` -- ``out_chunks`` for the
    broadcast result, targets equal to the operand's own chunks where nothing has to move."""
    nd = max((len(s) for s, _, _ in operands), default=0)
    # axis position `a` counts from the RIGHT (NumPy broadcasting), so operands of different rank line up
    votes = [[] for _ in range(nd)]          # per axis: (layout, extent, nbytes)
    for shape, chunks, item in operands:
        nbytes = float(math.prod(shape) * item)
        for n in range(len(shape)):
            votes[len(shape) - 1 - n].append((tuple(chunks[n]), shape[n], nbytes))

    def pick(axis_votes, fn):
        seen = {lay for lay, _, _ in axis_votes}
        if len(seen) > 1:
            seen -= {(1,)}                   # extent-1 operands broadcast: no opinion
        return fn(seen)

    merge = policy != "refine"
    chosen = [pick(v, coarsest_nested if merge else refine) for v in votes]
    fine = None

    def fine_grid():
        nonlocal fine
        if fine is None:
            fine = [pick(v, refine) for v in votes]
        return fine

    if merge and policy != "coarse":
        for a, v in enumerate(votes):
            opinions = [(lay, nb) for lay, extent, nb in v if extent > 1 and len(lay) > 1]
            if not opinions:
                continue
            target = chosen[a]
            anchored = any(lay == target for lay, _ in opinions)
            at_target = sum(nb for lay, nb in opinions if lay == target)
            merge_cost = sum(nb * moved_share(lay, target) for lay, nb in opinions
                             if lay != target and len(target) < len(lay))
            refused = merge_cost > MERGE_COST_RATIO * at_target
            if refused:
                chosen[a] = target = fine_grid()[a]
            if (anchored and not refused) or any(lay == target for lay, _ in opinions):
                continue
            # nobody holds the chosen grid (interleaved layouts, or a refused merge): adopt an existing
            # grid if the others can join it for a proportionate cost -- fewest blocks, then cheapest
            weight = {}
            for lay, nb in opinions:
                weight[lay] = weight.get(lay, 0.0) + nb
            best = None
            for lay, anchor in weight.items():
                cost = sum(nb * moved_share(src, lay) for src, nb in opinions if src != lay)
                if cost <= MERGE_COST_RATIO * anchor:
                    cand = (len(lay), cost, -anchor, lay)
                    best = cand if best is None or cand < best else best
            if best is not None:
                chosen[a] = best[3]

    if limit and merge:
        worst = 0
        for shape, chunks, item in operands:
            r = len(shape)
            new = item * math.prod(max(chosen[r - 1 - n]) for n in range(r) if shape[n] > 1)
            old = item * math.prod(max(chunks[n]) for n in range(r) if shape[n] > 1)
            if new > old:
                worst = max(worst, new)
        if worst > limit:
            f = fine_grid()
            chosen = [f[a] if len(f[a]) > len(chosen[a]) else chosen[a] for a in range(nd)]

    targets = []
    for shape, chunks, _ in operands:
        r = len(shape)
        targets.append(tuple(chosen[r - 1 - n] if (shape[n] > 1 or shape[n] == 0) else (shape[n],)
                             for n in range(r)))
    return tuple(chosen[nd - 1 - d] for d in range(nd)), targets
