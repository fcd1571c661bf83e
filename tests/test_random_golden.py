"""The Random leaf against vectors recorded from the reference's own `_spawn_bitgens` /
`_apply_random_func` (tests/golden/generate_random.py): per-block host streams, bit for bit, including the
advance of the generator's SeedSequence from draw to draw.  CPU-only (the leaf is a host RNG, staged once)."""
import os

import numpy as np

import dask_array_b200 as da

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "random.npz"))
DRAWS = [("random", (6, 10), (3, 5), np.float32, ()), ("random", (7,), (3,), np.float64, ()),
         ("standard_normal", (4, 4), (2, 4), np.float64, ()), ("integers", (8,), (4,), np.int64, (0, 1000))]


def test_random_blocks_equal_the_reference_streams():
    for seed in (0, 42):
        rng = da.random.default_rng(seed)
        for d, (dist, shape, chunks, dtype, args) in enumerate(DRAWS):
            if dist == "integers":
                arr = rng.integers(*args, size=shape, chunks=chunks, dtype=dtype)
            else:
                arr = getattr(rng, dist)(shape, chunks=chunks, dtype=dtype)
            e = arr.expr
            for k, bid in enumerate(e.block_ids()):
                want = GOLD[f"seed{seed}_draw{d}_block{k}"]
                got = e.host_block(bid)
                assert got.dtype == want.dtype and got.shape == want.shape
                assert np.array_equal(got, want), (seed, d, k)


def test_random_names_follow_the_spawn_sequence():
    a, b = da.random.default_rng(1), da.random.default_rng(1)
    x1, x2 = a.random((8,), chunks=4), a.random((8,), chunks=4)
    y1 = b.random((8,), chunks=4)
    assert x1.name == y1.name and x1.name != x2.name          # same seed, same draw -> same name; next draw differs
    assert da.random.random((8,), chunks=4).name != da.random.random((8,), chunks=4).name      # seed=None: OS entropy
