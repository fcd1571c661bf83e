"""Golden fixtures for topk / argtopk: the reference's OWN chunk functions (``_chunk.py:200-290``: ``topk``,
``topk_aggregate``, ``argtopk_preprocess``, ``argtopk``, ``argtopk_aggregate``; the module needs NumPy
only and is loaded from /root/reference by path) applied the way ``routines/_topk.py:14-80`` wires them
into ``reduction``: chunk on every block, combine on groups of ``split_every`` partials concatenated
along the axis, aggregate at the end.  Run by hand:  python tests/golden/generate_topk.py
Writes tests/golden/topk.npz.  TEST INFRASTRUCTURE ONLY."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()
spec = importlib.util.spec_from_file_location("ref_chunk", "/root/reference/dask_array/_chunk.py")
chunk = importlib.util.module_from_spec(spec)
spec.loader.exec_module(chunk)


def blocks_along(xh, chunks, axis):
    bounds = [np.cumsum((0,) + tuple(c)) for c in chunks]
    import itertools
    out = {}
    for bid in itertools.product(*[range(len(c)) for c in chunks]):
        out[bid] = xh[tuple(slice(bounds[d][i], bounds[d][i + 1]) for d, i in enumerate(bid))]
    return out


def reference_topk(xh, chunks, k, axis, split_every=4):
    import itertools
    blocks = blocks_along(xh, chunks, axis)
    nb = [len(c) for c in chunks]
    res = {}
    for cid in itertools.product(*[range(n) if d != axis else [0] for d, n in enumerate(nb)]):
        parts = [chunk.topk(blocks[cid[:axis] + (i,) + cid[axis + 1:]], k, (axis,), True) for i in range(nb[axis])]
        while len(parts) > split_every:
            parts = [chunk.topk(np.concatenate(parts[i:i + split_every], axis=axis), k, (axis,), True)
                     for i in range(0, len(parts), split_every)]
        res[cid] = chunk.topk_aggregate(np.concatenate(parts, axis=axis), k, (axis,), True)
    return res


def reference_argtopk(xh, chunks, k, axis, split_every=4):
    import itertools
    blocks = blocks_along(xh, chunks, axis)
    nb = [len(c) for c in chunks]
    starts = np.cumsum((0,) + tuple(chunks[axis]))
    res = {}
    for cid in itertools.product(*[range(n) if d != axis else [0] for d, n in enumerate(nb)]):
        parts = []
        for i in range(nb[axis]):
            a = blocks[cid[:axis] + (i,) + cid[axis + 1:]]
            idx = np.arange(starts[i], starts[i + 1], dtype=np.intp)
            idx = idx[tuple(slice(None) if d == axis else np.newaxis for d in range(a.ndim))]
            parts.append(chunk.argtopk(chunk.argtopk_preprocess(a, idx), k, (axis,), True))
        while len(parts) > split_every:
            parts = [chunk.argtopk(parts[i:i + split_every], k, (axis,), True) for i in range(0, len(parts), split_every)]
        res[cid] = chunk.argtopk_aggregate(parts, k, (axis,), True)
    return res


def assemble(res, chunks, axis, keep, dtype):
    shape = tuple(keep if d == axis else sum(c) for d, c in enumerate(chunks))
    out = np.empty(shape, dtype=dtype)
    bounds = [np.cumsum((0,) + tuple(c)) for c in chunks]
    for cid, blk in res.items():
        sl = tuple(slice(0, keep) if d == axis else slice(bounds[d][i], bounds[d][i + 1]) for d, i in enumerate(cid))
        out[sl] = blk
    return out


def main():
    rng = np.random.default_rng(7)
    perm = lambda n: rng.permutation(n)                       # distinct values: argtopk is then unambiguous
    cases = {
        "vec_f8": (rng.permutation(5000).astype(np.float64) / 7, ((1200, 1300, 2500),), 5, 0),
        "vec_i4_smallest": (perm(3000).astype(np.int32) - 1500, ((1000,) * 3,), -7, 0),
        "mat_axis1": (perm(40 * 300).reshape(40, 300).astype(np.float32), ((20, 20), (100, 100, 100)), 4, 1),
        "mat_axis0": (perm(300 * 24).reshape(300, 24).astype(np.int64), ((64, 64, 64, 64, 44), (12, 12)), 3, 0),
        "k_exceeds_block": (perm(50).astype(np.float64), ((4,) * 12 + (2,),), 6, 0),
        "k_exceeds_axis": (perm(9).astype(np.float32), ((5, 4),), 20, 0),
        "cube_axis1_smallest": (perm(6 * 70 * 5).reshape(6, 70, 5).astype(np.float64), ((3, 3), (30, 40), (5,)), -2, 1),
        "long_rows": (perm(2 * 20000).reshape(2, 20000).astype(np.float32), ((2,), (20000,)), 10, 1),
    }
    nanv = rng.random(400)
    nanv[[3, 77, 250]] = np.nan
    cases["nan_largest"] = (nanv, ((100,) * 4,), 5, 0)
    out = {}
    for name, (xh, chunks, k, axis) in cases.items():
        keep = min(abs(k), xh.shape[axis])
        out[name + "/x"] = xh
        out[name + "/meta"] = np.array([repr((chunks, k, axis))])
        out[name + "/topk"] = assemble(reference_topk(xh, chunks, k, axis), chunks, axis, keep, xh.dtype)
        # argtopk: skipped for NaNs (their order among themselves is unspecified) and for k >= the whole axis
        # over several blocks, where the reference's own aggregate fails (argtopk returns its list input
        # unmerged, _chunk.py:258-259, and argtopk_aggregate unpacks it as (a, idx), :276)
        if name not in ("nan_largest", "k_exceeds_axis"):
            out[name + "/argtopk"] = assemble(reference_argtopk(xh, chunks, k, axis), chunks, axis, keep, np.intp)
    np.savez_compressed(os.path.join(HERE, "topk.npz"), **out)
    print(sorted(k for k in out if k.endswith("/topk")), out["vec_f8/topk"], out["vec_i4_smallest/argtopk"])


if __name__ == "__main__":
    main()
