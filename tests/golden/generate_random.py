"""Golden vectors for the Random leaf, recorded from the reference's OWN functions
``_spawn_bitgens`` and ``_apply_random_func`` (``dask_array/random/_expr.py:29-41``) imported through
``_refshim`` (build container only; ``python tests/golden/generate_random.py``).

What is pinned: for a generator ``PCG64(seed)``, three successive draws of different block counts / dtypes /
distributions -- the per-block arrays the reference's tasks would produce, in block order.  Successive draws
advance the generator's SeedSequence (``bitgen._seed_seq.spawn``), which is what the product's
``da.random.default_rng`` must reproduce bit for bit."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()


def _lenient(modname):
    m = sys.modules[modname]

    def ga(name, _m=modname):
        if name.startswith("__"):
            raise AttributeError(name)
        return _refshim._StubMeta(name, (_refshim._Stub,), {"__module__": _m})
    m.__getattr__ = ga


for sub in ("io", "creation", "core", "slicing", "manipulation", "stacking", "routines", "reductions", "linalg"):
    _lenient("dask_array." + sub)
_lenient("dask_array")
from dask_array.random import _expr as R  # noqa: E402

DRAWS = [  # (distribution, shape, chunks, kwargs, args)
    ("random", (6, 10), (3, 5), {"dtype": np.float32}, ()),
    ("random", (7,), (3,), {}, ()),
    ("standard_normal", (4, 4), (2, 4), {}, ()),
    ("integers", (8,), (4,), {"dtype": np.int64}, (0, 1000)),
]


def main():
    out = {}
    for seed in (0, 42):
        bitgen = np.random.PCG64(seed)
        for d, (dist, shape, chunks, kwargs, args) in enumerate(DRAWS):
            nb = [-(-n // c) for n, c in zip(shape, chunks)]
            sizes = [tuple(min(c, n - i * c) for i, n, c in zip(idx, shape, chunks)) for idx in np.ndindex(*nb)]
            kids = R._spawn_bitgens(bitgen, len(sizes))                       # advances bitgen._seed_seq
            for k, size in enumerate(sizes):
                blk = R._apply_random_func(type(bitgen), dist, kids[k]._seed_seq, size, args, kwargs)
                out[f"seed{seed}_draw{d}_block{k}"] = np.asarray(blk)
    np.savez_compressed(os.path.join(HERE, "random.npz"), **out)
    print(len(out), "blocks;", out["seed42_draw1_block2"])


if __name__ == "__main__":
    main()
