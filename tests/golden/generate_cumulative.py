"""Golden fixtures for cumulative reductions: builds the reference's OWN task
graph with ``CumReduction._layer`` (``/root/reference/dask_array/reductions/_cumulative.py:174-264``,
unmodified, through ``_refshim``) on seeded blocks, evaluates it with a 20-line interpreter of the
classic dask task tuples and records every output block.

Run by hand in the build container:  python tests/golden/generate_cumulative.py
Writes tests/golden/cumulative.npz.  TEST INFRASTRUCTURE ONLY.
"""
import operator
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()


def _apply(func, args, kwargs=None):        # dask.utils.apply
    return func(*args, **(kwargs or {}))


import dask.utils  # noqa: E402  (the shim's stub module)

dask.utils.apply = _apply
import dask_array.reductions._cumulative as C  # noqa: E402


def evaluate(dsk, key, cache):
    """Classic dask graph semantics: a tuple whose head is callable is a task, a key is looked up,
    lists are evaluated element-wise, everything else is a literal."""
    def ev(x):
        if isinstance(x, tuple) and x and callable(x[0]):
            return x[0](*[ev(a) for a in x[1:]])
        if isinstance(x, list):
            return [ev(a) for a in x]
        try:
            if x in dsk:
                if x not in cache:
                    cache[x] = ev(dsk[x])
                return cache[x]
        except TypeError:
            pass
        return x
    return ev(key)


def run_reference(xh, chunks, axis, func, binop, ident, dtype):
    from itertools import product
    bounds = [np.cumsum((0,) + c) for c in chunks]
    numblocks = tuple(len(c) for c in chunks)
    dsk = {}
    for bid in product(*map(range, numblocks)):
        sl = tuple(slice(bounds[d][i], bounds[d][i + 1]) for d, i in enumerate(bid))
        dsk[("x",) + bid] = xh[sl]
    arr = types.SimpleNamespace(name="x", numblocks=numblocks, chunks=chunks, ndim=xh.ndim, _meta=xh[(slice(0, 0),) * xh.ndim])
    me = types.SimpleNamespace(array=arr, axis=axis, func=func, binop=binop, ident=ident, dtype=dtype, _name="cum")
    dsk.update(C.CumReduction._layer(me))
    cache = {}
    return {bid: np.asarray(evaluate(dsk, ("cum",) + bid, cache)) for bid in product(*map(range, numblocks))}


def nancumsum(x, axis, dtype=None):          # _chunk.py nancumsum: np.nancumsum
    return np.nancumsum(x, axis=axis, dtype=dtype)


def nancumprod(x, axis, dtype=None):
    return np.nancumprod(x, axis=axis, dtype=dtype)


CASES = []
rng = np.random.default_rng(42)
f32 = rng.standard_normal((37, 50)).astype(np.float32)
f64 = rng.standard_normal((24, 18, 10))
i32 = rng.integers(-50, 50, size=(40, 33)).astype(np.int32)
u8 = rng.integers(0, 4, size=(64,)).astype(np.uint8)
vec = rng.standard_normal(10000)
nanv = f64.copy()
nanv[rng.random(f64.shape) < 0.1] = np.nan
small = rng.integers(1, 3, size=(12, 9)).astype(np.int64)
CASES = [
    ("f32_axis0", f32, ((10, 10, 10, 7), (25, 25)), 0, "cumsum", False),
    ("f32_axis1", f32, ((10, 10, 10, 7), (20, 20, 10)), 1, "cumsum", False),
    ("f64_3d_axis1", f64, ((12, 12), (5, 5, 5, 3), (10,)), 1, "cumsum", False),
    ("f64_3d_axis2", f64, ((24,), (9, 9), (4, 6)), 2, "cumsum", False),
    ("i32_axis0", i32, ((16, 16, 8), (33,)), 0, "cumsum", False),
    ("i32_axis1_prod", (i32 % 3 + 1).astype(np.int32), ((40,), (11, 11, 11)), 1, "cumprod", False),
    ("u8_vec", u8, ((16,) * 4,), 0, "cumsum", False),
    ("vec_f64", vec, ((3000, 3000, 3000, 1000),), 0, "cumsum", False),
    ("nan_f64_axis0", nanv, ((12, 12), (18,), (10,)), 0, "cumsum", True),
    ("nan_f64_axis2_prod", nanv, ((24,), (18,), (5, 5)), 2, "cumprod", True),
    ("i64_prod_axis0", small, ((4, 4, 4), (9,)), 0, "cumprod", False),
]


def main():
    out = {}
    for name, xh, chunks, axis, kind, nan in CASES:
        if kind == "cumsum":
            func, binop, ident = (nancumsum, operator.add, 0) if nan else (np.cumsum, C._cumsum_merge, 0)
        else:
            func, binop, ident = (nancumprod, operator.mul, 1) if nan else (np.cumprod, C._cumprod_merge, 1)
        dtype = func(np.ones((0,), dtype=xh.dtype), axis=0).dtype           # CumReduction.dtype (:115-119)
        blocks = run_reference(xh, chunks, axis, func, binop, ident, dtype)
        bounds = [np.cumsum((0,) + c) for c in chunks]
        full = np.empty(xh.shape, dtype=dtype)
        for bid, blk in blocks.items():
            full[tuple(slice(bounds[d][i], bounds[d][i + 1]) for d, i in enumerate(bid))] = blk
        out[name + "/x"] = xh
        out[name + "/result"] = full
        out[name + "/meta"] = np.array([repr((chunks, axis, kind, nan))])
    np.savez_compressed(os.path.join(HERE, "cumulative.npz"), **out)
    print(sorted(k for k in out if k.endswith("result")))


if __name__ == "__main__":
    main()
