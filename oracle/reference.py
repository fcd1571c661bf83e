"""CPU oracle: a NumPy restatement of dask-array's data-parallel hot path.

TEST INFRASTRUCTURE ONLY.  Imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker and the CPU
baseline, never by the product (``dask_array_b200`` does not import this package and fails
loudly without its CUDA library).

Why a restatement: the reference (``/root/reference/dask_array``) is pure Python on top of
the third-party packages ``dask`` (>=2025.4.0, uv.lock pins 2025.12.0) and ``toolz`` (1.1.0),
which are not installed here and cannot be (no network), so ``import dask_array`` fails.  All
arithmetic of the path lives in NumPy (installed, 2.3.5).  This module applies, per block and
in the reference's tree order, the same NumPy calls the reference's chunk / combine /
aggregate functions make; every function cites the reference lines it follows.

PARITY PINNED: ``tests/golden/generate.py`` imports the reference's OWN functions
(``reductions/_common.py``, ``_core_utils.py``, ``_rechunk.py``, ``linalg/_tensordot.py`` ...)
unmodified through a stub of the two missing packages (``tests/golden/_refshim.py``), runs
them on seeded inputs and stores inputs/outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this module against those fixtures bit for bit.  What
is restated without a runnable original is only the orchestration that lives inside ``dask``
itself (task ordering of the threaded scheduler), which does not affect values.
"""
from __future__ import annotations

import itertools
import math
from concurrent.futures import ThreadPoolExecutor
from functools import reduce as _fold

import numpy as np


# =============================================================================== blocked arrays
def normalize_chunks(chunks, shape):
    """(100, 100) -> ((100,)*10, (100,)*10); tuples of tuples pass through
    (the regular-chunk subset of ``_core_utils.py:731 normalize_chunks``)."""
    out = []
    for c, n in zip(chunks, shape):
        if isinstance(c, (tuple, list)):
            assert sum(c) == n, (c, n)
            out.append(tuple(int(v) for v in c))
        else:
            c = int(c)
            full, rest = divmod(n, c) if c else (0, 0)
            out.append((c,) * full + ((rest,) if rest else ()) if n else (0,))
    return tuple(out)


def block_slices(chunks):
    """slices of every block, keyed by block id (``slices_from_chunks``)."""
    starts = [np.concatenate([[0], np.cumsum(c)]) for c in chunks]
    out = {}
    for bid in itertools.product(*[range(len(c)) for c in chunks]):
        out[bid] = tuple(slice(int(starts[d][i]), int(starts[d][i + 1])) for d, i in enumerate(bid))
    return out


class Blocked:
    """A chunked array held as {block id: ndarray} -- what the scheduler's result dict
    holds after every task of a layer ran."""

    def __init__(self, blocks, chunks):
        self.blocks = blocks
        self.chunks = tuple(tuple(c) for c in chunks)

    @property
    def numblocks(self):
        return tuple(len(c) for c in self.chunks)

    @property
    def shape(self):
        return tuple(sum(c) for c in self.chunks)

    @property
    def ndim(self):
        return len(self.chunks)

    @classmethod
    def from_array(cls, x, chunks):
        """``io/_from_array.py:60`` -- blocks are slices (views) of ``x``."""
        x = np.asarray(x)
        chunks = normalize_chunks(chunks, x.shape)
        return cls({bid: x[sl] for bid, sl in block_slices(chunks).items()}, chunks)

    def to_array(self):
        """finalize -> concatenate3 (``_core_utils.py:1426-1448, 1182-1248``): allocate the
        result once and assign every block into its slot."""
        first = next(iter(self.blocks.values()))
        out = np.empty(self.shape, dtype=first.dtype)
        for bid, sl in block_slices(self.chunks).items():
            out[sl] = self.blocks[bid]
        return out


def _pmap(fn, items, workers):
    if not workers or workers <= 1:
        return [fn(i) for i in items]
    with ThreadPoolExecutor(max_workers=workers) as ex:      # dask.threaded's model
        return list(ex.map(fn, items))


# =============================================================================== blockwise
def broadcast_block_id(out_bid, dep_numblocks):
    """``_blockwise.py:1243 _broadcast_block_id``: right-aligned; 1-block dims use block 0."""
    off = len(out_bid) - len(dep_numblocks)
    return tuple(out_bid[off + d] if nb > 1 else 0 for d, nb in enumerate(dep_numblocks))


def elemwise(func, *args, workers=0):
    """Elemwise (``_blockwise.py:837, 1030-1074``): ``func(*blocks)`` per output block with
    NumPy broadcasting; non-Blocked args (Python scalars) are passed through unchanged, which
    keeps NEP-50 weak promotion (``f4 * 2 -> f4``)."""
    arrs = [a for a in args if isinstance(a, Blocked)]
    nd = max(a.ndim for a in arrs)
    out_chunks = []
    for d in range(nd):
        cands = [a.chunks[d - (nd - a.ndim)] for a in arrs if d - (nd - a.ndim) >= 0]
        best = max(cands, key=lambda c: (sum(c), len(c)))
        for c in cands:
            assert c == best or c == (1,), f"operands must be chunk-aligned: {c} vs {best}"
        out_chunks.append(best)
    out_chunks = tuple(out_chunks)

    def one(bid):
        vals = [a.blocks[broadcast_block_id(bid, a.numblocks)] if isinstance(a, Blocked) else a for a in args]
        return bid, func(*vals)

    bids = list(itertools.product(*[range(len(c)) for c in out_chunks]))
    return Blocked(dict(_pmap(one, bids, workers)), out_chunks)


def transpose(x, axes=None):
    """Transpose (``manipulation/_transpose.py:14-75``): ``np.transpose`` per block (a view)
    and the block grid permuted the same way (``_input_block_id`` :73)."""
    axes = tuple(reversed(range(x.ndim))) if axes is None else tuple(axes)
    blocks = {tuple(bid[a] for a in axes): np.transpose(b, axes) for bid, b in x.blocks.items()}
    return Blocked(blocks, tuple(x.chunks[a] for a in axes))


def broadcast_trick(value, shape, chunks, dtype):
    """Ones/Zeros/Full (``creation/_ones_zeros.py:17-137``, ``creation/_utils.py:65-72``):
    every block is a zero-stride broadcast view of one element."""
    chunks = normalize_chunks(chunks, shape)
    one = np.full((1,) * len(shape), value, dtype=dtype)
    blocks = {bid: np.broadcast_to(one, tuple(s.stop - s.start for s in sl))
              for bid, sl in block_slices(chunks).items()}
    return Blocked(blocks, chunks)


def getitem_slices(x, index):
    """Basic slicing (``slicing/_basic.py:357-493``): equals NumPy slicing of the whole array,
    re-blocked on the surviving part of each chunk."""
    full = x.to_array()[index]
    new_chunks = []
    for d, ix in enumerate(index):
        if isinstance(ix, slice):
            start, stop, step = ix.indices(x.shape[d])
            assert step == 1
            edges = np.concatenate([[0], np.cumsum(x.chunks[d])])
            cs = [int(min(stop, hi) - max(start, lo)) for lo, hi in zip(edges[:-1], edges[1:])]
            new_chunks.append(tuple(c for c in cs if c > 0) or (0,))
    return Blocked.from_array(full, tuple(new_chunks))


# =============================================================================== chunk kernels
def numel(x, axis=None, keepdims=False, dtype=np.float64):
    """``_dispatch.py:209-238``: element count as a broadcast (zero-stride) array."""
    shape = x.shape
    if axis is None:
        prod = np.prod(shape, dtype=dtype)
        return np.full((1,) * len(shape), prod, dtype=dtype) if keepdims else prod
    axis = [axis] if not isinstance(axis, (tuple, list)) else axis
    prod = math.prod(shape[d] for d in axis)
    new_shape = tuple(1 if d in axis else shape[d] for d in range(len(shape))) if keepdims else \
        tuple(shape[d] for d in range(len(shape)) if d not in axis)
    return np.broadcast_to(np.array(prod, dtype=dtype), new_shape)


def concatenate2(arrays, axes):
    """``_core_utils.py:191-252``: nested lists -> one array, one nesting level per axis;
    dicts are concatenated field by field."""
    if axes == () or axes == []:
        return arrays[0] if isinstance(arrays, list) else arrays
    if not isinstance(arrays, (list, tuple)):
        return arrays
    if len(axes) > 1:
        arrays = [concatenate2(a, axes[1:]) for a in arrays]
    if isinstance(arrays[0], dict):
        return {k: np.concatenate([a[k] for a in arrays], axis=axes[0]) for k in arrays[0]}
    return np.concatenate(arrays, axis=axes[0])


def deepmap(fn, seq):
    return [deepmap(fn, s) for s in seq] if isinstance(seq, list) else fn(seq)


def mean_chunk(x, dtype, axis, keepdims=True):
    """``reductions/_common.py:270-281``."""
    n = numel(x, dtype=dtype, axis=axis, keepdims=keepdims)
    if 0 in n.strides:
        n = np.full((1,) * x.ndim, n.flat[0] if n.size else 0, dtype=n.dtype)
    return {"n": n, "total": np.sum(x, dtype=dtype, axis=axis, keepdims=keepdims)}


def mean_combine(pairs, dtype, axis, keepdims=True):
    """``_common.py:284-305``."""
    pairs = pairs if isinstance(pairs, list) else [pairs]
    n = concatenate2(deepmap(lambda p: p["n"], pairs), axes=axis).sum(axis=axis, keepdims=keepdims)
    total = concatenate2(deepmap(lambda p: p["total"], pairs), axes=axis).sum(axis=axis, keepdims=keepdims)
    return {"n": n, "total": total}


def mean_agg(pairs, dtype, axis, keepdims=False):
    """``_common.py:308-320``; ``divide`` -> ``np.true_divide(.., dtype=dtype)`` (``_dispatch.py:157``)."""
    pairs = pairs if isinstance(pairs, list) else [pairs]
    n = np.sum(concatenate2(deepmap(lambda p: p["n"], pairs), axes=axis), axis=axis, dtype=dtype, keepdims=keepdims)
    total = concatenate2(deepmap(lambda p: p["total"], pairs), axes=axis).sum(axis=axis, dtype=dtype, keepdims=keepdims)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.true_divide(total, n, dtype=dtype)


def moment_chunk(A, dtype, axis, keepdims=True, order=2):
    """``_common.py:368-404``: two passes per block -- total, then sum((A - total/n)**k)."""
    n = numel(A, axis=axis, keepdims=keepdims)
    n = np.full((1,) * A.ndim, n.flat[0] if n.size else 0, dtype=np.int64) if 0 in n.strides else n.astype(np.int64)
    total = np.sum(A, dtype=dtype, axis=axis, keepdims=keepdims)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = total / n
    d = A - u
    xs = [np.sum(d**i, dtype=dtype, axis=axis, keepdims=keepdims) for i in range(2, order + 1)]
    return {"total": total, "n": n, "M": np.stack(xs, axis=-1)}


def _moment_helper(Ms, ns, inner, order, axis, kw):
    """``_common.py:407-412`` (binomial cross terms for order > 2)."""
    M = Ms[..., order - 2].sum(axis=axis, **kw) + np.sum(ns * inner**order, axis=axis, **kw)
    for k in range(1, order - 1):
        coeff = math.factorial(order) / (math.factorial(k) * math.factorial(order - k))
        M += coeff * np.sum(Ms[..., order - k - 2] * inner**k, axis=axis, **kw)
    return M


def moment_combine(pairs, dtype, axis, order=2):
    """``_common.py:415-453`` (Chan et al. merge of block moments)."""
    pairs = pairs if isinstance(pairs, list) else [pairs]
    kw = dict(dtype=None, keepdims=True)
    ns = concatenate2(deepmap(lambda p: p["n"], pairs), axes=axis)
    n = ns.sum(axis=axis, **kw)
    totals = concatenate2(deepmap(lambda p: p["total"], pairs), axes=axis)
    Ms = concatenate2(deepmap(lambda p: p["M"], pairs), axes=axis)
    total = totals.sum(axis=axis, **kw)
    with np.errstate(divide="ignore", invalid="ignore"):
        mu = np.true_divide(total, n, dtype=dtype)
        inner = np.true_divide(totals, ns, dtype=dtype) - mu
    xs = [_moment_helper(Ms, ns, inner, o, axis, kw) for o in range(2, order + 1)]
    return {"total": total, "n": n, "M": np.stack(xs, axis=-1)}


def moment_agg(pairs, dtype, axis, keepdims=False, order=2, ddof=0):
    """``_common.py:456-505``."""
    pairs = pairs if isinstance(pairs, list) else [pairs]
    kw = dict(dtype=dtype, keepdims=keepdims)
    kd = dict(dtype=None, keepdims=True)
    ns = concatenate2(deepmap(lambda p: p["n"], pairs), axes=axis)
    n = ns.sum(axis=axis, **kd)
    totals = concatenate2(deepmap(lambda p: p["total"], pairs), axes=axis)
    Ms = concatenate2(deepmap(lambda p: p["M"], pairs), axes=axis)
    with np.errstate(divide="ignore", invalid="ignore"):
        mu = np.true_divide(totals.sum(axis=axis, **kd), n)
        inner = np.true_divide(totals, ns, dtype=dtype) - mu
    inner = np.where(ns == 0, 0, inner)
    M = _moment_helper(Ms, ns, inner, order, axis, kw)
    den = n.sum(axis=axis, **kw) - ddof
    if np.ndim(den) == 0:
        den = np.nan if den < 0 else den
    else:
        den = den.astype(np.float64) if den.dtype.kind in "iu" else den
        den[den < 0] = np.nan
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.true_divide(M, den, dtype=dtype)


def chunk_min(x, axis=None, keepdims=None):
    """``_common.py:92-97``."""
    return np.array([], dtype=x.dtype, ndmin=x.ndim) if x.size == 0 else np.min(x, axis=axis, keepdims=keepdims)


def chunk_max(x, axis=None, keepdims=None):
    """``_common.py:100-105``."""
    return np.array([], dtype=x.dtype, ndmin=x.ndim) if x.size == 0 else np.max(x, axis=axis, keepdims=keepdims)


def _kd(fn):
    """``_chunk.py:137-168 keepdims_wrapper`` for np.argmin/np.argmax (the result is indexed as it is --
    NumPy scalars accept ``r[None, ...]`` -- so a duck-array result stays a duck array)."""
    def wrapped(x, axis=None, keepdims=None):
        r = fn(x, axis=axis)
        if not keepdims:
            return r
        axes = range(x.ndim) if axis is None else ([axis] if np.isscalar(axis) else axis)
        return r[tuple(None if d in axes else slice(None) for d in range(x.ndim))]
    wrapped.__name__ = fn.__name__
    return wrapped


argmin_kd, argmax_kd = _kd(np.argmin), _kd(np.argmax)


def _arg_result(vals, arg):
    """``_common.py:722-731`` / ``:738-746``: a structured array, or -- for chunk types without
    structured dtypes (``np.empty_like`` raises TypeError) -- a dict."""
    try:
        out = np.empty_like(vals, shape=vals.shape, dtype=[("vals", vals.dtype), ("arg", arg.dtype)])
    except TypeError:
        out = dict()
    out["vals"], out["arg"] = vals, arg
    return out


def arg_chunk(func, argfunc, x, axis, offset_info):
    """``_common.py:704-732``."""
    arg_axis = None if len(axis) == x.ndim or x.ndim == 1 else axis[0]
    vals = func(x, axis=arg_axis, keepdims=True)
    arg = argfunc(x, axis=arg_axis, keepdims=True)
    if x.ndim > 0:
        if arg_axis is None:
            offset, total_shape = offset_info
            ind = np.unravel_index(arg.ravel()[0], x.shape)
            total_ind = tuple(o + i for (o, i) in zip(offset, ind))
            arg[:] = np.ravel_multi_index(total_ind, total_shape)
        else:
            arg += offset_info
    return _arg_result(vals, arg)


def _arg_combine(data, axis, argfunc, keepdims=False):
    """``_common.py:675-701`` (both the structured-array and the dict representation)."""
    ndim = data["vals"].ndim if isinstance(data, dict) else data.ndim
    axis = None if len(axis) == ndim or ndim == 1 else axis[0]
    vals, arg = data["vals"], data["arg"]
    if axis is None:
        local = argfunc(vals, axis=axis, keepdims=keepdims)
        return arg.ravel()[local], vals.ravel()[local]
    local = argfunc(vals, axis=axis)
    inds = list(np.ogrid[tuple(map(slice, local.shape))])
    inds.insert(axis, local)
    vals, arg = vals[tuple(inds)], arg[tuple(inds)]
    if keepdims:
        vals, arg = np.expand_dims(vals, axis), np.expand_dims(arg, axis)
    return arg, vals


def arg_combine(argfunc, data, axis):
    """``_common.py:735-746``."""
    arg, vals = _arg_combine(data, axis, argfunc, keepdims=True)
    return _arg_result(vals, arg)


def arg_agg(argfunc, data, axis, keepdims=False):
    """``_common.py:749-750``."""
    return _arg_combine(data, axis, argfunc, keepdims=keepdims)[0]


# =============================================================================== tree reduction
def normalize_split_every(split_every, axis):
    """``reductions/_reduction.py:715-725`` (config default ``split_every`` = 16)."""
    split_every = split_every or 16
    if isinstance(split_every, dict):
        return {k: split_every.get(k, 2) for k in axis}
    n = max(int(split_every ** (1 / (len(axis) or 1))), 2)
    return dict.fromkeys(axis, n)


def partition_all(n, seq):
    seq = list(seq)
    return [tuple(seq[i:i + n]) for i in range(0, len(seq), n)]


def _lol(blocks, ndim, fixed, groups, d=0, prefix=()):
    """``dask.blockwise.lol_tuples``: one nesting level per reduced axis, axis order."""
    if d == ndim:
        return blocks[prefix]
    if d in fixed:
        return _lol(blocks, ndim, fixed, groups, d + 1, prefix + (fixed[d],))
    return [_lol(blocks, ndim, fixed, groups, d + 1, prefix + (i,)) for i in groups[d]]


def partial_reduce(x, func, split_every, keepdims, workers=0):
    """One ``PartialReduce`` level (``_reduction.py:900-983``)."""
    nb = x.numblocks
    parts = [partition_all(split_every.get(d, 1), range(n)) for d, n in enumerate(nb)]
    out_chunks = [tuple(1 for _ in p) if d in split_every else x.chunks[d] for d, p in enumerate(parts)]
    kept = [d for d in range(x.ndim) if d not in split_every]

    def one(k):
        p = [parts[d][i] for d, i in enumerate(k)]
        fixed = {d: g[0] for d, g in enumerate(p) if len(g) == 1 and d not in split_every}
        groups = {d: g for d, g in enumerate(p) if d in split_every}
        key = k if keepdims else tuple(k[d] for d in kept)
        return key, func(_lol(x.blocks, x.ndim, fixed, groups))

    keys = list(itertools.product(*[range(len(p)) for p in parts]))
    blocks = dict(_pmap(one, keys, workers))
    if not keepdims:
        out_chunks = [out_chunks[d] for d in kept]
    return Blocked(blocks, tuple(out_chunks))


def tree_reduce(x, aggregate, axis, keepdims, split_every=None, combine=None, concatenate=True, workers=0):
    """``_build_tree_reduce_expr`` (``_reduction.py:751-806``): depth-1 combine levels, then
    one aggregate level."""
    split_every = normalize_split_every(split_every, axis)
    depth = 1
    for d, n in enumerate(x.numblocks):
        if d in split_every and split_every[d] != 1:
            depth = int(max(depth, math.ceil(math.log(n, split_every[d])))) if n > 1 else depth
    comb = combine or aggregate
    srt = sorted(axis)
    func = (lambda lst: comb(concatenate2(lst, axes=srt), axis=axis, keepdims=True)) if concatenate else \
        (lambda lst: comb(lst, axis=axis, keepdims=True))
    for _ in range(depth - 1):
        x = partial_reduce(x, func, split_every, True, workers)
    agg = (lambda lst: aggregate(concatenate2(lst, axes=srt), axis=axis, keepdims=keepdims)) if concatenate else \
        (lambda lst: aggregate(lst, axis=axis, keepdims=keepdims))
    return partial_reduce(x, agg, split_every, keepdims, workers)


def _norm_axis(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    axis = (axis,) if np.isscalar(axis) else tuple(axis)
    return tuple(a % ndim for a in axis)


def reduction(x, chunk, aggregate, axis=None, keepdims=False, split_every=None, combine=None,
              concatenate=True, workers=0):
    """``Reduction._lower`` (``_reduction.py:154-226``): chunk step per block with
    ``keepdims=True`` (fused with whatever produced the block), then the tree."""
    axis = _norm_axis(axis, x.ndim)

    def one(item):
        bid, b = item
        return bid, chunk(b, axis=axis, keepdims=True)

    blocks = dict(_pmap(one, list(x.blocks.items()), workers))
    chunks = tuple(tuple(1 for _ in c) if d in axis else c for d, c in enumerate(x.chunks))
    return tree_reduce(Blocked(blocks, chunks), aggregate, axis, keepdims, split_every, combine, concatenate, workers)


def _finish(b):
    return b.to_array() if b.ndim else b.blocks[()]


def da_sum(x, axis=None, keepdims=False, dtype=None, split_every=None, workers=0):
    """``_common.py:57-70``: dtype = what np.sum gives (ints widen to 64 bit)."""
    first = next(iter(x.blocks.values()))
    dt = dtype or np.zeros(1, dtype=first.dtype).sum().dtype
    ch = lambda b, axis, keepdims: np.sum(b, axis=axis, keepdims=keepdims, dtype=dt)
    return _finish(reduction(x, ch, ch, axis, keepdims, split_every, workers=workers))


def da_min(x, axis=None, keepdims=False, split_every=None, workers=0):
    """``_common.py:109-122``."""
    agg = lambda b, axis, keepdims: np.min(b, axis=axis, keepdims=keepdims)
    return _finish(reduction(x, chunk_min, agg, axis, keepdims, split_every, combine=chunk_min, workers=workers))


def da_max(x, axis=None, keepdims=False, split_every=None, workers=0):
    """``_common.py:125-138``."""
    agg = lambda b, axis, keepdims: np.max(b, axis=axis, keepdims=keepdims)
    return _finish(reduction(x, chunk_max, agg, axis, keepdims, split_every, combine=chunk_max, workers=workers))


def _mean_dtype(x, dtype):
    first = next(iter(x.blocks.values()))
    return np.dtype(dtype) if dtype is not None else np.mean(np.zeros((1,), dtype=first.dtype)).dtype


def da_mean(x, axis=None, keepdims=False, dtype=None, split_every=None, workers=0):
    """``_common.py:323-343``."""
    dt = _mean_dtype(x, dtype)
    return _finish(reduction(
        x, lambda b, axis, keepdims: mean_chunk(b, dt, axis, keepdims),
        lambda p, axis, keepdims: mean_agg(p, dt, axis, keepdims), axis, keepdims, split_every,
        combine=lambda p, axis, keepdims: mean_combine(p, dt, axis, keepdims), concatenate=False, workers=workers))


def da_var(x, axis=None, keepdims=False, dtype=None, ddof=0, split_every=None, workers=0):
    """``_common.py:572-593``."""
    first = next(iter(x.blocks.values()))
    dt = np.dtype(dtype) if dtype is not None else np.var(np.ones((1,), dtype=first.dtype)).dtype
    return _finish(reduction(
        x, lambda b, axis, keepdims: moment_chunk(b, dt, axis, keepdims),
        lambda p, axis, keepdims: moment_agg(p, dt, axis, keepdims, ddof=ddof), axis, keepdims, split_every,
        combine=lambda p, axis, keepdims: moment_combine(p, dt, axis), concatenate=False, workers=workers))


def da_std(x, axis=None, keepdims=False, dtype=None, ddof=0, split_every=None, workers=0):
    """``_common.py:625-653``: ``sqrt(var(...))``."""
    return np.sqrt(da_var(x, axis, keepdims, dtype, ddof, split_every, workers))


def _arg_reduction(x, func, argfunc, axis, keepdims, split_every, workers):
    """``reductions/_arg_reduction.py:66-150``: ArgChunk with per-block offsets, then the tree
    with arg_combine / arg_agg (concatenate=True)."""
    if axis is None:
        ax, ravel = tuple(range(x.ndim)), True
    else:
        ax, ravel = (axis % x.ndim,), x.ndim == 1
    starts = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in x.chunks]

    def one(item):
        bid, b = item
        off = tuple(int(starts[d][i]) for d, i in enumerate(bid))
        info = (off, x.shape) if ravel else off[ax[0]]
        return bid, arg_chunk(func, argfunc, b, ax, info)

    blocks = dict(_pmap(one, list(x.blocks.items()), workers))
    chunks = tuple(tuple(1 for _ in c) if d in ax else c for d, c in enumerate(x.chunks))
    res = tree_reduce(Blocked(blocks, chunks), lambda d, axis, keepdims: arg_agg(argfunc, d, axis, keepdims),
                      ax, keepdims, split_every, combine=lambda d, axis, keepdims: arg_combine(argfunc, d, axis),
                      workers=workers)
    return _finish(res)


def da_argmax(x, axis=None, keepdims=False, split_every=None, workers=0):
    """``_common.py:775-786``."""
    return _arg_reduction(x, np.max, argmax_kd, axis, keepdims, split_every, workers)


def da_argmin(x, axis=None, keepdims=False, split_every=None, workers=0):
    """``_common.py:789-799``."""
    return _arg_reduction(x, np.min, argmin_kd, axis, keepdims, split_every, workers)


# =============================================================================== rechunk
def old_to_new(old, new):
    """Per new block, the (old block, slice) pieces that make it up
    (``_rechunk.py:130-175 old_to_new`` for known chunk sizes)."""
    out = []
    for oc, nc in zip(old, new):
        edges = np.concatenate([[0], np.cumsum(oc)])
        dim, pos = [], 0
        for n in nc:
            lo, hi, pieces = pos, pos + n, []
            if n == 0:
                pieces.append((0, slice(0, 0)))
            for i in range(len(oc)):
                a, b = max(lo, int(edges[i])), min(hi, int(edges[i + 1]))
                if a < b:
                    pieces.append((i, slice(a - int(edges[i]), b - int(edges[i]))))
            dim.append(pieces)
            pos = hi
        out.append(dim)
    return out


def rechunk(x, new_chunks, workers=0):
    """TasksRechunk (``_rechunk.py:1171-1187, 1252-1323``): every new block is the
    concatenation (``concatenate3``) of ``getitem`` slices of old blocks.  The planner's
    intermediate stages (``plan_rechunk`` :442-516) change cost, not values, so one direct
    stage gives the identical array."""
    new_chunks = normalize_chunks(new_chunks, x.shape)
    o2n = old_to_new(x.chunks, new_chunks)

    def one(nbid):
        per_dim = [o2n[d][i] for d, i in enumerate(nbid)]
        shape = tuple(new_chunks[d][i] for d, i in enumerate(nbid))
        first = next(iter(x.blocks.values()))
        out = np.empty(shape, dtype=first.dtype)
        pos = [np.concatenate([[0], np.cumsum([p[1].stop - p[1].start for p in pieces])]) for pieces in per_dim]
        for combo in itertools.product(*[range(len(p)) for p in per_dim]):
            obid = tuple(per_dim[d][k][0] for d, k in enumerate(combo))
            sl = tuple(per_dim[d][k][1] for d, k in enumerate(combo))
            dst = tuple(slice(int(pos[d][k]), int(pos[d][k + 1])) for d, k in enumerate(combo))
            out[dst] = x.blocks[obid][sl]
        return nbid, out

    bids = list(itertools.product(*[range(len(c)) for c in new_chunks]))
    return Blocked(dict(_pmap(one, bids, workers)), new_chunks)


# =============================================================================== matmul
def matmul(a, b, workers=0):
    """``linalg/_tensordot.py:253-334``: one ``np.matmul`` per (i, k, j) block triple keeping
    the contracted axis as size 1 (``_matmul`` :194-213), then ``_sum_wo_cat`` (:216-249):
    ``reduce(np.add, partials)`` in ascending k through the split_every tree."""
    assert a.ndim == 2 and b.ndim == 2 and a.chunks[1] == b.chunks[0]
    ni, nk, nj = len(a.chunks[0]), len(a.chunks[1]), len(b.chunks[1])
    dt = np.result_type(next(iter(a.blocks.values())).dtype, next(iter(b.blocks.values())).dtype)

    def one(ikj):
        i, k, j = ikj
        return ikj, np.matmul(a.blocks[(i, k)], b.blocks[(k, j)])[..., np.newaxis, :]

    trip = dict(_pmap(one, list(itertools.product(range(ni), range(nk), range(nj))), workers))
    part = Blocked(trip, (a.chunks[0], (1,) * nk, b.chunks[1]))
    sdt = np.zeros(1, dtype=dt).sum().dtype

    def chunk_sum(lst, axis, keepdims):
        out = _fold(lambda p, q: np.add(p, q, dtype=sdt), lst) if isinstance(lst, list) else lst
        return out if keepdims else out.squeeze(axis[0])

    if nk == 1:
        return Blocked({(i, j): v.squeeze(1) for (i, _, j), v in trip.items()}, (a.chunks[0], b.chunks[1]))
    return tree_reduce(part, chunk_sum, (1,), False, None, None, concatenate=False, workers=workers)


# =============================================================================== config helpers
def fused_chain(x):
    """The BASELINE config-2 chain on one block, as the reference evaluates it: four NumPy
    calls, four temporaries (``sin``, ``*2``, ``**2``, ``+``)."""
    return np.sin(x) * 2 + x**2


def fused_chain_mean_std(xh, chunks, workers=0):
    """``(sin(x)*2 + x**2).mean(axis=0)`` and ``.std()`` in the reference's block/tree order."""
    x = Blocked.from_array(xh, chunks)
    y = elemwise(fused_chain, x, workers=workers)
    return da_mean(y, axis=0, workers=workers), da_std(y, workers=workers)


# ----------------------------------------------------------------------------- chunk unification
# Restatement of ``unify_chunks_expr`` and its helpers (``_expr.py:586-905``,
# ``_core_utils.py:893-960``) for element-wise operands.  Pinned by tests/golden/unify.json, which
# tests/golden/generate_unify.py records from the reference's own functions.
MERGE_COST_RATIO = 4                    # _expr.py:669
UNIFY_CHUNKS_LIMIT = 512 * 2**20        # dask_array/__init__.py:25 ("512 MiB")


def common_blockdim(blockdims):
    """Finest common refinement of several chunkings of one axis (``_core_utils.py:893-960``)."""
    blockdims = set(map(tuple, blockdims))
    if not any(blockdims):
        return ()
    multi = {d for d in blockdims if len(d) > 1}
    if len(multi) == 1:
        return next(iter(multi))
    if not multi:
        return max(blockdims, key=lambda d: d[0])
    if len({sum(d) for d in multi}) > 1:
        raise ValueError("Chunks do not add up to same value", blockdims)
    stacks = [list(d)[::-1] for d in multi]
    total, done, out = sum(next(iter(multi))), 0, []
    while done < total:
        m = min(s[-1] for s in stacks)
        out.append(m)
        for s in stacks:
            s[-1] -= m
            if s[-1] == 0:
                s.pop()
        done += m
    return tuple(out)


def coarse_blockdim(blockdims):
    """Coarsest chunking when every other one nests inside it, else ``common_blockdim``
    (``_expr.py:586-660``)."""
    blockdims = set(map(tuple, blockdims))
    if not any(blockdims):
        return ()
    multi = {d for d in blockdims if len(d) > 1}
    if not multi:
        return max(blockdims, key=lambda d: d[0])
    if len(multi) == 1:
        return next(iter(multi))
    if len({sum(d) for d in multi}) > 1:
        raise ValueError("Chunks do not add up to same value", blockdims)
    coarsest = min(multi, key=len)
    edges = set(np.cumsum(coarsest[:-1]).tolist())
    for d in multi:
        if d != coarsest and not edges.issubset(set(np.cumsum(d[:-1]).tolist())):
            return common_blockdim(blockdims)
    return coarsest


def moved_fraction(src, dst):
    """Fraction of an axis's bytes a rechunk ``src -> dst`` moves when every ``dst`` chunk is
    assembled where its largest ``src`` piece already lives (``_expr.py:672-720``)."""
    total = sum(src)
    if not total or tuple(src) == tuple(dst) or sum(dst) != total:
        return 0.0
    moved, i, s0, d0 = 0.0, 0, 0.0, 0.0
    for t in dst:
        d1, best = d0 + t, 0.0
        while True:
            s1 = s0 + src[i]
            best = max(best, min(s1, d1) - max(s0, d0))
            if s1 <= d1 and i + 1 < len(src):
                i, s0 = i + 1, s1
            else:
                break
        moved += t - best
        d0 = d1
    return moved / total


def _broadcast_dimensions(pairs, consolidate):
    """``dask.blockwise.broadcast_dimensions``: per index label the set of operand chunkings, the
    broadcast sentinel ``(1,)`` dropped when something else is present."""
    g = {}
    for chunks, ind in pairs:
        for j, c in zip(ind, chunks):
            g.setdefault(j, set()).add(tuple(c))
    return {j: consolidate(v - {(1,)} if len(v) > 1 else v) for j, v in g.items()}


def unify_chunks(operands, policy="auto", limit=UNIFY_CHUNKS_LIMIT):
    """``unify_chunks_expr`` (``_expr.py:723-905``) for element-wise operands.

    ``operands``: [(shape, chunks, itemsize)], NumPy right-aligned; index labels as
    ``Elemwise.args`` builds them (``_blockwise.py:985-1001``): axis n of an operand carries label
    ``ndim - 1 - n``.  Returns ``(chunkss by label, [target chunks per operand], changed)``.
    """
    ops = []
    for shape, chunks, itemsize in operands:
        nd = len(shape)
        if nd == 0:
            ops.append(None)                       # scalars carry no layout (:741)
            continue
        ops.append((tuple(shape), tuple(map(tuple, chunks)), tuple(range(nd))[::-1],
                    float(math.prod(shape) * itemsize), itemsize))
    live = [o for o in ops if o is not None]
    inds = [o[2] for o in live]
    if live and all(i == inds[0] for i in inds) and all(o[1] == live[0][1] for o in live):
        return dict(zip(inds[0], live[0][1])), [o[1] if o else () for o in ops], False      # :733-734
    pairs = [(o[1], o[2]) for o in live]
    consolidate = common_blockdim if policy == "refine" else coarse_blockdim
    chunkss = _broadcast_dimensions(pairs, consolidate)
    fine = None
    if consolidate is coarse_blockdim and policy != "coarse":
        moved, anchored, layouts, seen = {}, {}, {}, set()
        for k, o in enumerate(live):
            shape, chunks, ind, nbytes, _ = o
            for n, j in enumerate(ind):
                src, target = chunks[n], chunkss[j]
                if shape[n] <= 1 or len(src) <= 1:
                    continue
                layouts.setdefault(j, []).append((src, nbytes))
                if src == target:
                    anchored[j] = anchored.get(j, 0.0) + nbytes
                elif len(target) < len(src):
                    moved[j] = moved.get(j, 0.0) + nbytes * moved_fraction(src, target)
        refused = {j for j, cost in moved.items() if cost > MERGE_COST_RATIO * anchored.get(j, 0.0)}
        if refused:
            fine = _broadcast_dimensions(pairs, common_blockdim)
            chunkss = {j: fine[j] if j in refused else c for j, c in chunkss.items()}
        for j, lay in layouts.items():                                                         # :815-838
            if j in anchored and j not in refused:
                continue
            target = chunkss[j]
            if any(src == target for src, _ in lay):
                continue
            cands = {}
            for src, nb in lay:
                cands[src] = cands.get(src, 0.0) + nb
            feasible = []
            for layout, anchor in cands.items():
                cost = sum(nb * moved_fraction(src, layout) for src, nb in lay if src != layout)
                if cost <= MERGE_COST_RATIO * anchor:
                    feasible.append((len(layout), cost, -anchor, layout))
            if feasible:
                chunkss[j] = min(feasible)[3]
    if limit and consolidate is coarse_blockdim:                                               # :840-872
        worst = 0
        for shape, chunks, ind, _, itemsize in live:
            target = itemsize * math.prod(max(chunkss[j]) for n, j in enumerate(ind) if shape[n] > 1)
            current = itemsize * math.prod(max(c) for n, c in enumerate(chunks) if shape[n] > 1)
            if target > current:
                worst = max(worst, target)
        if worst > limit:
            if fine is None:
                fine = _broadcast_dimensions(pairs, common_blockdim)
            coarsened = {j for j, c in chunkss.items() if len(fine[j]) > len(c)}
            chunkss = {j: fine[j] if j in coarsened else c for j, c in chunkss.items()}
    out, changed = [], False
    for o in ops:
        if o is None:
            out.append(())
            continue
        shape, chunks, ind, _, _ = o
        tgt = tuple(chunkss[j] if (shape[n] > 1 or shape[n] == 0) else (shape[n],) for n, j in enumerate(ind))
        changed |= tgt != chunks
        out.append(tgt)
    return chunkss, out, changed


# ----------------------------------------------------------------------------- cumulative scans
def da_cumulative(x, axis, kind="cumsum", nan=False, dtype=None):
    """``CumReduction._layer`` (``reductions/_cumulative.py:174-264``, sequential method): scan every
    block (``np.cumsum`` / ``np.cumprod`` or their nan variants, ``_chunk.py`` nancumsum/nancumprod),
    then walk the blocks along ``axis``: ``extra_i = extra_{i-1} (+|*) last hyperplane of block i-1``
    (identity for an empty block, ``_cum_tail`` :28-39) and ``result_i = extra_i (+|*) scanned_i``;
    the first block is its own scan.  Pinned bit for bit by tests/golden/cumulative.npz."""
    func = {("cumsum", False): np.cumsum, ("cumsum", True): np.nancumsum,
            ("cumprod", False): np.cumprod, ("cumprod", True): np.nancumprod}[(kind, bool(nan))]
    binop = np.add if kind == "cumsum" else np.multiply
    ident = 0 if kind == "cumsum" else 1
    any_block = next(iter(x.blocks.values()))
    dtype = np.dtype(dtype) if dtype is not None else func(np.ones((0,), dtype=any_block.dtype), axis=0).dtype
    nb = x.numblocks
    out = {}
    tail = (slice(None),) * axis + (slice(-1, None),)
    for cid in itertools.product(*[range(n) if d != axis else [0] for d, n in enumerate(nb)]):
        extra = None
        prev = None
        for i in range(nb[axis]):
            bid = cid[:axis] + (i,) + cid[axis + 1:]
            scanned = func(x.blocks[bid], axis=axis, dtype=dtype)
            if i == 0:
                shape = tuple(1 if d == axis else n for d, n in enumerate(scanned.shape))
                extra = np.full(shape, ident, dtype=dtype)
                out[bid] = scanned
            else:
                if prev.shape[axis] == 0:
                    t = np.full(tuple(1 if d == axis else n for d, n in enumerate(prev.shape)), ident, dtype=prev.dtype)
                else:
                    t = prev[tail]
                extra = binop(extra, t)
                out[bid] = binop(extra, scanned)
            prev = scanned
    return Blocked(out, x.chunks)


# ----------------------------------------------------------------------------- overlap (halo exchange)
_PAD_MODE = {"periodic": "wrap", "reflect": "symmetric", "nearest": "edge"}


def overlap(x, depth, boundary):
    """``overlap`` (``_overlap.py:906-987``) restated on the whole array: pad every axis as its boundary
    condition says (``periodic`` :715 = wrap, ``reflect`` :733 = symmetric -- the edge cell is repeated --,
    ``nearest`` :759 = edge, a value = constant, ``"none"`` = no pad), then every block of the ORIGINAL
    grid is cut out together with ``depth`` cells on each side (clipped at un-padded edges).  ``x``:
    Blocked whose chunks already hold the depth; ``depth`` / ``boundary``: {axis: value}.  The docstring
    example of the reference (:935-962) is asserted in tests/test_gpu_overlap.py."""
    full = x.to_array()
    nd = full.ndim
    padded = full
    for ax in range(nd):
        d, kind = depth.get(ax, 0), boundary.get(ax, "none")
        if d == 0 or (isinstance(kind, str) and kind == "none"):
            continue
        width = [(0, 0)] * nd
        width[ax] = (d, d)
        if isinstance(kind, str):
            padded = np.pad(padded, width, mode=_PAD_MODE[kind])
        else:
            padded = np.pad(padded, width, mode="constant", constant_values=kind)
    blocks, chunks = {}, []
    for ax in range(nd):
        d, kind = depth.get(ax, 0), boundary.get(ax, "none")
        has_pad = d > 0 and not (isinstance(kind, str) and kind == "none")
        n = len(x.chunks[ax])
        chunks.append(tuple(c + (d if (j > 0 or has_pad) else 0) + (d if (j < n - 1 or has_pad) else 0)
                            for j, c in enumerate(x.chunks[ax])))
    for bid in itertools.product(*[range(len(c)) for c in x.chunks]):
        sl = []
        for ax, j in enumerate(bid):
            d, kind = depth.get(ax, 0), boundary.get(ax, "none")
            has_pad = d > 0 and not (isinstance(kind, str) and kind == "none")
            start = sum(x.chunks[ax][:j]) + (d if has_pad else 0)          # position in the padded array
            lo = start - (d if (j > 0 or has_pad) else 0)
            hi = start + x.chunks[ax][j] + (d if (j < len(x.chunks[ax]) - 1 or has_pad) else 0)
            sl.append(slice(lo, hi))
        blocks[bid] = padded[tuple(sl)]
    return Blocked(blocks, tuple(chunks))


# ----------------------------------------------------------------------------- topk / argtopk
def _chunk_topk(a, k, axis):
    """``chunk.topk`` (``_chunk.py:200-215``): the k largest (k < 0: -k smallest) along axis, unsorted."""
    if abs(k) >= a.shape[axis]:
        return a
    a = np.partition(a, -k, axis=axis)
    k_slice = slice(-k, None) if k > 0 else slice(-k)
    return a[tuple(k_slice if i == axis else slice(None) for i in range(a.ndim))]


def da_topk(x, k, axis=-1, split_every=4):
    """``topk`` (``routines/_topk.py:14-40``): ``reduction`` with chunk = combine = ``chunk.topk`` and
    aggregate = ``chunk.topk_aggregate`` (``_chunk.py:218-229``: topk once more, then sort -- descending for
    k > 0).  Pinned by tests/golden/topk.npz."""
    axis %= x.ndim
    nb = x.numblocks
    out = {}
    for cid in itertools.product(*[range(n) if d != axis else [0] for d, n in enumerate(nb)]):
        parts = [_chunk_topk(x.blocks[cid[:axis] + (i,) + cid[axis + 1:]], k, axis) for i in range(nb[axis])]
        while len(parts) > split_every:
            parts = [_chunk_topk(np.concatenate(parts[i:i + split_every], axis=axis), k, axis)
                     for i in range(0, len(parts), split_every)]
        a = np.sort(_chunk_topk(np.concatenate(parts, axis=axis), k, axis), axis=axis)
        if k > 0:
            a = a[tuple(slice(None, None, -1) if i == axis else slice(None) for i in range(a.ndim))]
        out[cid] = a
    keep = min(abs(k), x.shape[axis])
    return Blocked(out, tuple((keep,) if d == axis else c for d, c in enumerate(x.chunks)))


def da_argtopk(x, k, axis=-1):
    """``argtopk`` (``routines/_topk.py:43-80``; ``chunk.argtopk`` / ``argtopk_aggregate`` ``_chunk.py:240-281``):
    the positions along ``axis`` of the elements ``topk`` returns, in the same order.  Restated on the
    whole chain (for distinct values the tree shape cannot change the answer)."""
    axis %= x.ndim
    nb = x.numblocks
    out = {}
    for cid in itertools.product(*[range(n) if d != axis else [0] for d, n in enumerate(nb)]):
        a = np.concatenate([x.blocks[cid[:axis] + (i,) + cid[axis + 1:]] for i in range(nb[axis])], axis=axis)
        order = np.argsort(a, axis=axis, kind="stable")
        if k > 0:
            order = order[tuple(slice(None, None, -1) if i == axis else slice(None) for i in range(a.ndim))]
        out[cid] = order[tuple(slice(0, abs(k)) if i == axis else slice(None) for i in range(a.ndim))].astype(np.intp)
    keep = min(abs(k), x.shape[axis])
    return Blocked(out, tuple((keep,) if d == axis else c for d, c in enumerate(x.chunks)))


# =============================================================================== sliding-window reductions
def da_sliding_window_reduce(x, window, axis, reducer="sum", workers=0):
    """``reduction(sliding_window_view(x, window, axis), axis=-1)`` the way the reference's overlap plan
    evaluates it (``_overlap.py:1365-1433`` + the reduction chunk functions): every block is extended by the
    ``window - 1`` elements that follow it along ``axis`` and reduced over NumPy's strided window view.
    (The native-chunk plan ``SlidingWindowReduction``, ``reductions/_sliding_window.py:405-560``, produces the
    same values from suffix / total / prefix pieces.)  Returns a Blocked with the trimmed chunks (:431-446)."""
    full = x.to_array()
    n = full.shape[axis]
    remaining = n - window + 1
    emit, starts, pos = [], [], 0
    for c in x.chunks[axis]:
        if remaining <= 0:
            break
        take = min(c, remaining)
        emit.append(take)
        starts.append(pos)
        pos += c
        remaining -= take
    fn = getattr(np, reducer)

    def one(i):
        sl = [slice(None)] * full.ndim
        sl[axis] = slice(starts[i], starts[i] + emit[i] + window - 1)
        blk = full[tuple(sl)]
        return fn(np.lib.stride_tricks.sliding_window_view(blk, window, axis=axis), axis=-1)

    parts = [p for _, p in sorted(_pmap(lambda i: (i, one(i)), list(range(len(emit))), workers))]
    out = np.concatenate(parts, axis=axis)
    chunks = list(x.chunks)
    chunks[axis] = tuple(emit)
    return Blocked.from_array(out, tuple(chunks))

