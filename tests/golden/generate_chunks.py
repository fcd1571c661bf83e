"""Golden fixtures for the chunk conventions (``normalize_chunks`` / ``auto_chunks`` with and without
``previous_chunks``, ``_balance_chunksizes``): runs the reference's OWN functions
(``/root/reference/dask_array/_core_utils.py:524-885``, ``_rechunk.py:519-560``), unmodified, through ``_refshim`` on
seeded random requests and records what they return (an exception is recorded by its type name).

Run by hand in the build container (never on the GPU box):  python tests/golden/generate_chunks.py
Writes tests/golden/chunks.json.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()
from dask_array._core_utils import normalize_chunks  # noqa: E402
from dask_array._rechunk import _balance_chunksizes  # noqa: E402


def blocks(rng, n):
    if n == 0:
        return (0,)
    if rng.random() < 0.5:
        c = min(n, rng.choice([1, 2, 5, 10, 64, 100, 1000, 4096, n]))
        full, rest = divmod(n, c)
        return (c,) * full + ((rest,) if rest else ())
    parts, rem = [], n
    while rem > 0:
        t = rng.randint(1, max(1, min(rem, rng.choice([3, 50, 2000]))))
        parts.append(t)
        rem -= t
    return tuple(parts)


def rle(seq):
    """Run-length form [[value, count], ...] -- regular blockings are long runs of one length."""
    out = []
    for v in seq:
        v = int(v)
        if out and out[-1][0] == v:
            out[-1][1] += 1
        else:
            out.append([v, 1])
    return out


def call(fn, *a, **k):
    try:
        out = fn(*a, **k)
        return [rle(c) for c in out]
    except Exception as e:          # noqa: BLE001 -- the error type is part of the recorded behaviour
        return type(e).__name__


def main():
    rng = random.Random(2024)
    plain, scaled, balanced = [], [], []
    sizes = [0, 1, 7, 100, 5000, 40000]
    for _ in range(400):
        nd = rng.randint(0, 3)
        shape = [rng.choice(sizes) for _ in range(nd)]
        chunks = [rng.choice(["auto", "auto", -1, None, 1, 3, 64, 1000, "2MiB"]) for _ in range(nd)]
        if rng.random() < 0.2:
            chunks = rng.choice(["auto", 5, "64MiB", -1])
        dt = rng.choice(["f4", "f8", "i1", "i8", "c16"])
        limit = rng.choice([None, None, "1MiB", 5000, 10**9])
        req = tuple(chunks) if isinstance(chunks, list) else chunks
        plain.append(dict(chunks=chunks, shape=shape, dtype=dt, limit=limit,
                          out=call(normalize_chunks, req, tuple(shape), limit=limit, dtype=np.dtype(dt))))
    for _ in range(400):
        nd = rng.randint(1, 3)
        shape = [rng.choice(sizes[1:]) for _ in range(nd)]
        prev = [blocks(rng, n) for n in shape]
        chunks = [rng.choice(["auto", "auto", "auto", -1, 3, 64, "2MiB"]) for _ in range(nd)]
        if rng.random() < 0.2:
            chunks = rng.choice(["auto", "64MiB", "10kiB"])
        dt = rng.choice(["f4", "f8", "i1", "c16"])
        limit = rng.choice([None, None, None, "1MiB", 5000, 10**9, 100])
        req = tuple(chunks) if isinstance(chunks, list) else chunks
        scaled.append(dict(chunks=chunks, shape=shape, dtype=dt, limit=limit, previous=[rle(p) for p in prev],
                           out=call(normalize_chunks, req, tuple(shape), limit=limit, dtype=np.dtype(dt),
                                    previous_chunks=tuple(prev))))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(300):
            ch = blocks(rng, rng.choice([5, 17, 100, 1000, 4097])) if rng.random() < 0.6 else \
                tuple(rng.randint(1, rng.choice([3, 20, 500])) for _ in range(rng.randint(1, 12)))
            balanced.append(dict(chunks=rle(ch), out=rle(_balance_chunksizes(ch))))
    path = os.path.join(HERE, "chunks.json")
    with open(path, "w") as f:
        json.dump(dict(normalize=plain, previous=scaled, balance=balanced), f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes;", len(plain), len(scaled), len(balanced), "cases")


if __name__ == "__main__":
    main()
