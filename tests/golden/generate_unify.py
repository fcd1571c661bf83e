"""Golden fixtures for chunk unification: runs the reference's OWN
``unify_chunks_expr`` / ``coarse_blockdim`` / ``common_blockdim`` / ``moved_fraction``
(``/root/reference/dask_array/_expr.py:586-905``, ``_core_utils.py:893-960``), unmodified, through
``_refshim`` on light operand stand-ins (name, shape, chunks, dtype) and records what they return.

Run by hand in the build container (never on the GPU box):  python tests/golden/generate_unify.py
Writes tests/golden/unify.json.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()
import dask_array._expr as E  # noqa: E402
from dask_array._core_utils import common_blockdim  # noqa: E402


class Operand:
    """What ``unify_chunks_expr`` touches of an array expression."""

    def __init__(self, name, shape, chunks, dtype="float64"):
        self._name, self.shape, self.chunks, self.dtype = name, tuple(shape), tuple(map(tuple, chunks)), np.dtype(dtype)
        self.ndim = len(self.shape)
        self.numblocks = tuple(len(c) for c in self.chunks)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    def rechunk(self, chunks):
        return Operand(self._name + "-rechunked", self.shape, chunks, self.dtype)


def uniform(n, c):
    q, r = divmod(n, c)
    return (c,) * q + ((r,) if r else ())


CASES = {
    # name: [(shape, chunks, dtype), ...]   -- element-wise operands, NumPy right-aligned
    "same": [((200, 200), ((100, 100), (100, 100)), "f8"), ((200, 200), ((100, 100), (100, 100)), "f8")],
    "nested_2d": [((200, 200), ((100, 100), (100, 100)), "f8"), ((200, 200), ((50,) * 4, (200,)), "f8")],
    "interleaved": [((10,), ((4, 6),), "f8"), ((10,), ((6, 4),), "f8")],
    "vector_broadcast": [((200, 200), ((100, 100), (100, 100)), "f8"), ((200,), ((50,) * 4,), "f8")],
    "light_coarse_refused": [((1000, 64), ((10,) * 100, (64,)), "f8"), ((1000, 1), ((500, 500), (1,)), "f8")],
    "comparable_merge": [((1000, 64), ((10,) * 100, (64,)), "f4"), ((1000, 64), ((500, 500), (64,)), "f8")],
    "roll_shift": [((400,), ((100,) * 4,), "f8"), ((400,), ((50, 100, 100, 100, 50),), "f8")],
    "roll_sliver": [((2880,), ((720,) * 4,), "f8"), ((2880,), ((1, 720, 720, 720, 719),), "f8")],
    "size_limit": [((65536, 65536), ((65536,), (64,) * 1024), "f8"), ((65536, 65536), ((64,) * 1024, (65536,)), "f8")],
    "three_operands": [((120, 120), (uniform(120, 30), uniform(120, 40)), "f8"),
                       ((120, 120), (uniform(120, 60), uniform(120, 20)), "f4"),
                       ((120,), (uniform(120, 10),), "i8")],
    "row_broadcast": [((200, 200), ((100, 100), (100, 100)), "f8"), ((1, 200), ((1,), (50,) * 4), "f8")],
    "ragged_nested": [((1000, 700), ((300, 300, 300, 100), (256, 256, 188)), "f4"),
                      ((1000, 700), ((600, 400), (700,)), "f4")],
    "zero_d": [((64, 64), ((32, 32), (16,) * 4), "f8"), ((), (), "f8")],
    "transpose_config4": [((16384, 16384), ((16384,), (256,) * 64), "f4"), ((16384, 16384), ((256,) * 64, (16384,)), "f4")],
    "single_vs_chunked": [((10,), ((10,),), "f8"), ((10,), ((5, 5),), "f8")],
    "many_fine_vs_coarse_heavy": [((4096, 4096), (uniform(4096, 128), uniform(4096, 128)), "f4"),
                                  ((4096, 4096), (uniform(4096, 1024), uniform(4096, 1024)), "f4")],
}

BLOCKDIM_SETS = [
    [(12, 12, 12, 12), (6,) * 8], [(10,), (5, 5)], [(4, 6), (6, 4)], [(3,), (2, 1)], [(1, 2), (2, 1)],
    [(100,) * 4, (50, 100, 100, 100, 50)], [(30, 30, 30), (10,) * 9, (90,)], [(7, 3), (5, 5), (2, 8)], [(1,), (1,)],
]
MOVED = [((1, 719, 720), (720, 720)), ((10,) * 6, (30, 30)), ((30, 30), (10,) * 6), ((100,) * 4, (50, 100, 100, 100, 50)),
         ((5, 5), (5, 5)), ((4, 6), (6, 4)), ((256,) * 64, (16384,)), ((300, 300, 300, 100), (600, 400))]


def main():
    out = {"unify": {}, "coarse_blockdim": [], "common_blockdim": [], "moved_fraction": []}
    for name, ops in CASES.items():
        arrays = [Operand(f"op{k}", s, c, d) for k, (s, c, d) in enumerate(ops)]
        args = []
        for a in arrays:
            args += [a, tuple(range(a.ndim))[::-1]]
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            chunkss, new, changed = E.unify_chunks_expr(*args)
        out["unify"][name] = {
            "operands": [[list(s), [list(c) for c in ch], d] for s, ch, d in ops],
            "chunkss": {str(k): list(v) for k, v in sorted(chunkss.items())},
            "result_chunks": [[list(c) for c in a.chunks] for a in new],
            "changed": bool(changed),
            "warned": [type(x.message).__name__ for x in w],
        }
    for bd in BLOCKDIM_SETS:
        out["coarse_blockdim"].append([[list(b) for b in bd], list(E.coarse_blockdim(set(bd)))])
        out["common_blockdim"].append([[list(b) for b in bd], list(common_blockdim(set(bd)))])
    for src, dst in MOVED:
        out["moved_fraction"].append([list(src), list(dst), E.moved_fraction(src, dst)])
    with open(os.path.join(HERE, "unify.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print({k: (v["chunkss"] if len(str(v["chunkss"])) < 120 else "...") for k, v in out["unify"].items()})


if __name__ == "__main__":
    main()
