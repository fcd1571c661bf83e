"""Generate golden fixtures by running the REFERENCE's own hot-path functions.

Run by hand in the build container (``python tests/golden/generate.py``); needs
``/root/reference``.  The reference sources are imported unmodified through
``_refshim`` (a stub of the absent ``dask`` / ``toolz`` packages).  Outputs:
``tests/golden/hotpath.npz`` (arrays) and ``tests/golden/hotpath.json`` (structure).
The GPU box never runs this; the tests only read the two committed files.
"""
from __future__ import annotations

import json
import os
import sys
from functools import partial
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()

from dask_array import _chunk as chunk  # noqa: E402
from dask_array import _core_utils as CU  # noqa: E402
from dask_array import _rechunk as RC  # noqa: E402
from dask_array._dispatch import _numel  # noqa: E402
from dask_array.linalg import _tensordot as TD  # noqa: E402
from dask_array.reductions import _common as C  # noqa: E402
from dask_array.reductions import _reduction as RED  # noqa: E402

A = {}   # arrays
J = {}   # json-able structure


def put(name, value):
    if isinstance(value, dict):
        for k, v in value.items():
            A[f"{name}.{k}"] = np.asarray(v)
    else:
        A[name] = np.asarray(value)


def blocks_of(x, chunks):
    edges = [np.concatenate([[0], np.cumsum(c)]) for c in chunks]
    out = {}
    for bid in np.ndindex(*[len(c) for c in chunks]):
        out[bid] = x[tuple(slice(int(edges[d][i]), int(edges[d][i + 1])) for d, i in enumerate(bid))]
    return out


def lol(blocks, numblocks, axes, groups_fixed):
    """nested list over the reduced axes (axis order), other axes fixed"""
    def rec(d, prefix):
        if d == len(numblocks):
            return blocks[prefix]
        if d in axes:
            return [rec(d + 1, prefix + (i,)) for i in groups_fixed[d]]
        return rec(d + 1, prefix + (groups_fixed[d],))
    return rec(0, ())


rng = np.random.default_rng(20261018)

# ---------------------------------------------------------------- 1. mean: chunk / combine / agg
for tag, dt, acc in [("f4", np.float32, "f4"), ("i4", np.int32, "f8"), ("f8", np.float64, "f8")]:
    x = (rng.random((12, 10)) * 100).astype(dt)
    put(f"mean.{tag}.x", x)
    chunks = ((5, 7), (4, 6))
    bl = blocks_of(x, chunks)
    for axis in [(0,), (1,), (0, 1)]:
        at = "".join(map(str, axis))
        parts = {bid: C.mean_chunk(b, dtype=acc, axis=axis, keepdims=True) for bid, b in bl.items()}
        for bid, p in parts.items():
            put(f"mean.{tag}.ax{at}.chunk{bid[0]}{bid[1]}", p)
        fixed = {0: [0, 1] if 0 in axis else 0, 1: [0, 1] if 1 in axis else 1}
        nested = lol(parts, (2, 2), axis, fixed)
        put(f"mean.{tag}.ax{at}.combine", C.mean_combine(nested, dtype=acc, axis=axis, keepdims=True))
        put(f"mean.{tag}.ax{at}.agg", C.mean_agg(nested, dtype=acc, axis=axis, keepdims=False))

# ---------------------------------------------------------------- 2. moments (var)
for tag, dt, acc in [("f4", np.float32, "f4"), ("f8", np.float64, "f8"), ("i4", np.int32, "f8")]:
    x = (rng.random((12, 10)) * 10 + 1000).astype(dt)      # large mean: cancellation-prone
    put(f"var.{tag}.x", x)
    bl = blocks_of(x, ((5, 7), (4, 6)))
    for axis in [(0,), (1,), (0, 1)]:
        at = "".join(map(str, axis))
        parts = {bid: C.moment_chunk(b, dtype=acc, axis=axis, keepdims=True) for bid, b in bl.items()}
        for bid, p in parts.items():
            put(f"var.{tag}.ax{at}.chunk{bid[0]}{bid[1]}", p)
        fixed = {0: [0, 1] if 0 in axis else 0, 1: [0, 1] if 1 in axis else 1}
        nested = lol(parts, (2, 2), axis, fixed)
        put(f"var.{tag}.ax{at}.combine", C.moment_combine(nested, dtype=acc, axis=axis))
        for ddof in (0, 1):
            put(f"var.{tag}.ax{at}.agg.ddof{ddof}", C.moment_agg(nested, dtype=acc, axis=axis, keepdims=False, ddof=ddof))

# ---------------------------------------------------------------- 3. arg reductions (ties, NaN, inf)
x = np.floor(rng.random((9, 14)) * 6)
x[1, 3] = np.nan; x[1, 9] = np.nan; x[2, 0] = np.inf; x[3, 5] = -np.inf; x[4, :] = 2.0
put("arg.x", x)
chunks = ((4, 5), (6, 8))
bl = blocks_of(x, chunks)
starts = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in chunks]
for nm, func, argfunc in [("max", chunk.max, chunk.argmax), ("min", chunk.min, chunk.argmin)]:
    for axis in [(0,), (1,), (0, 1)]:
        at = "".join(map(str, axis))
        parts = {}
        for bid, b in bl.items():
            off = tuple(int(starts[d][i]) for d, i in enumerate(bid))
            info = (off, x.shape) if len(axis) == 2 else off[axis[0]]
            parts[bid] = C.arg_chunk(func, argfunc, b, axis, info)
            A[f"arg.{nm}.ax{at}.chunk{bid[0]}{bid[1]}.vals"] = parts[bid]["vals"]
            A[f"arg.{nm}.ax{at}.chunk{bid[0]}{bid[1]}.arg"] = parts[bid]["arg"]
        fixed = {0: [0, 1] if 0 in axis else 0, 1: [0, 1] if 1 in axis else 1}
        nested = lol(parts, (2, 2), axis, fixed)
        data = CU._concatenate2(nested, axes=sorted(axis))
        comb = C.arg_combine(argfunc, data, axis=axis)
        A[f"arg.{nm}.ax{at}.combine.vals"] = comb["vals"]
        A[f"arg.{nm}.ax{at}.combine.arg"] = comb["arg"]
        A[f"arg.{nm}.ax{at}.agg"] = np.asarray(C.arg_agg(argfunc, data, axis=axis, keepdims=False))

# ---------------------------------------------------------------- 4. min/max chunk kernels, numel, concatenate
y = rng.random((6, 7)).astype(np.float32); y[2, 2] = np.nan
put("minmax.x", y)
put("minmax.min1", C.chunk_min(y, axis=(1,), keepdims=True))
put("minmax.max0", C.chunk_max(y, axis=(0,), keepdims=True))
put("numel.ax0", np.array(_numel(y, axis=(0,), keepdims=True, dtype="f4")))
put("numel.all", np.array(_numel(y, axis=(0, 1), keepdims=True, dtype="f8")))
a, b = rng.random((2, 3)), rng.random((2, 3))
put("cat2.a", a); put("cat2.b", b)
put("cat2.ax0", CU._concatenate2([a, b], axes=[0]))
put("cat2.ax01", CU._concatenate2([[a, b], [b, a]], axes=[0, 1]))
put("cat3", CU.concatenate3([[a, b], [b, a]]))

# ---------------------------------------------------------------- 5. tree shape (split_every, PartialReduce nesting)
J["split_every"] = {
    "None,(0,)": RED._normalize_split_every(None, (0,)),
    "None,(0,1)": RED._normalize_split_every(None, (0, 1)),
    "None,(0,1,2)": RED._normalize_split_every(None, (0, 1, 2)),
    "2,(0,1)": RED._normalize_split_every(2, (0, 1)),
    "{0:3},(0,1)": RED._normalize_split_every({0: 3}, (0, 1)),
    "4,(1,)": RED._normalize_split_every(4, (1,)),
}
J["split_every"] = {k: {str(a): n for a, n in v.items()} for k, v in J["split_every"].items()}
layers = {}
for tag, numblocks, split, keepdims in [
    ("std8x8", (8, 8), {0: 4, 1: 4}, True), ("mean8x8", (8, 8), {0: 16}, False),
    ("ragged5x3", (5, 3), {0: 2, 1: 2}, True), ("ax1_3x7", (3, 7), {1: 4}, False),
]:
    fake = SimpleNamespace(
        array=SimpleNamespace(numblocks=numblocks, name="x", ndim=len(numblocks)),
        split_every=split, keepdims=keepdims, func="F", _name="out",
    )
    dsk = RED.PartialReduce._layer(fake)
    layers[tag] = {
        "numblocks": list(numblocks), "split_every": {str(k): v for k, v in split.items()}, "keepdims": keepdims,
        "tasks": [[list(k[1:]), json.loads(json.dumps(v[1]))] for k, v in dsk.items()],
    }
J["partial_reduce_layers"] = layers

# ---------------------------------------------------------------- 6. rechunk planner + intersections
J["plan_rechunk"] = {}
for tag, item in [("c4_f8", 8), ("c4_f4", 4)]:
    old = ((16384,), (256,) * 64)
    new = ((256,) * 64, (16384,))
    steps = RC.plan_rechunk(old, new, item)
    J["plan_rechunk"][tag] = [[list(map(int, d)) for d in step] for step in steps]
inter = {}
for tag, old, new in [
    ("ragged", ((4, 4, 3), (2, 2, 2)), ((2, 6, 3), (6,))),
    ("split", ((10,), (10,)), ((3, 3, 4), (5, 5))),
    ("panels", ((8,), (2, 2, 2, 2)), ((2, 2, 2, 2), (8,))),
]:
    res = list(RC.intersect_chunks(old, new))
    inter[tag] = {
        "old": [list(c) for c in old], "new": [list(c) for c in new],
        "pieces": [[[[int(i), int(s.start), int(s.stop)] for (i, s) in piece] for piece in newblock] for newblock in res],
    }
J["intersect_chunks"] = inter

# ---------------------------------------------------------------- 7. matmul block kernel
ma = rng.random((5, 4)).astype(np.float32); mb = rng.random((4, 3)).astype(np.float32)
put("matmul.a", ma); put("matmul.b", mb)
put("matmul.out", TD._matmul(ma, mb))

# ---------------------------------------------------------------- 8. getitem copy semantics, broadcast trick
from dask_array.creation._utils import _broadcast_trick_inner  # noqa: E402

bt = _broadcast_trick_inner(np.ones_like, (3, 4), meta=np.empty((0, 0), np.float64))
J["broadcast_trick"] = {"shape": list(bt.shape), "strides": list(bt.strides), "value": float(bt[0, 0])}
g = chunk.getitem(x, (slice(0, 2), slice(0, 3)))
J["getitem_small_is_copy"] = bool(g.flags.owndata)

np.savez_compressed(os.path.join(HERE, "hotpath.npz"), **A)
with open(os.path.join(HERE, "hotpath.json"), "w") as f:
    json.dump(J, f, indent=1, sort_keys=True)
print(f"wrote {len(A)} arrays, {os.path.getsize(os.path.join(HERE, 'hotpath.npz'))} bytes")
