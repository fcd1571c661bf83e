"""Chunk unification (host logic, CPU): the product's policy (`dask_array_b200/_unify.py`) against
the golden outputs of the reference's own `unify_chunks_expr` (tests/golden/unify.json) and against
the oracle on seeded random layouts; and the expression-level effect (Elemwise.chunks, Rechunk
insertion) -- reference: `_expr.py:586-905`, `_blockwise.py:94-98,1003-1028`."""
import json
import os
import random

import numpy as np
import pytest

from oracle import reference as ref

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "unify.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", sorted(_golden()["unify"]))
def test_product_policy_matches_reference_golden(case):
    from dask_array_b200._unify import unify
    rec = _golden()["unify"][case]
    operands = [(tuple(s), tuple(map(tuple, ch)), np.dtype(d).itemsize) for s, ch, d in rec["operands"]]
    out, targets = unify(operands)
    assert [[list(c) for c in t] for t in targets] == rec["result_chunks"]
    nd = len(out)
    assert {str(nd - 1 - d): list(c) for d, c in enumerate(out)} == rec["chunkss"]


def test_helpers_match_reference_golden():
    from dask_array_b200 import _unify as u
    g = _golden()
    for sets, want in g["coarse_blockdim"]:
        assert list(u.coarsest_nested([tuple(s) for s in sets])) == want, sets
    for sets, want in g["common_blockdim"]:
        assert list(u.refine([tuple(s) for s in sets])) == want, sets
    for src, dst, want in g["moved_fraction"]:
        assert u.moved_share(src, dst) == want


def _random_layout(rng, n):
    kind = rng.choice(["uniform", "uniform", "ragged", "single", "shifted"])
    if kind == "single" or n < 4:
        return (n,)
    if kind == "uniform":
        c = rng.choice([d for d in (2, 3, 4, 5, 8, 10, 16, 25, 32, 50, 64, 100) if d <= n])
        q, r = divmod(n, c)
        return (c,) * q + ((r,) if r else ())
    if kind == "shifted":
        c = max(2, n // rng.randint(2, 6))
        s = rng.randint(1, c - 1)
        rest = n - s
        q, r = divmod(rest, c)
        return (s,) + (c,) * q + ((r,) if r else ())
    cuts = sorted(rng.sample(range(1, n), min(n - 1, rng.randint(1, 6))))
    edges = [0] + cuts + [n]
    return tuple(b - a for a, b in zip(edges, edges[1:]))


@pytest.mark.parametrize("seed", range(40))
def test_product_policy_matches_oracle_on_random_layouts(seed):
    from dask_array_b200._unify import unify
    rng = random.Random(seed)
    nd = rng.randint(1, 3)
    shape = tuple(rng.choice([1, 12, 60, 100, 240]) for _ in range(nd))
    operands = []
    for _ in range(rng.randint(2, 4)):
        r = rng.randint(1, nd)
        shp = tuple((1 if rng.random() < 0.15 else n) for n in shape[nd - r:])
        chunks = tuple(_random_layout(rng, n) for n in shp)
        operands.append((shp, chunks, rng.choice([1, 4, 8])))
    chunkss, want, _ = ref.unify_chunks(operands)
    out, got = unify(operands)
    assert list(got) == list(want)
    assert {len(out) - 1 - d: c for d, c in enumerate(out)} == chunkss


def test_elemwise_chunks_and_rechunk_insertion():
    import dask_array_b200 as da
    from dask_array_b200._rechunk import Rechunk, TasksRechunk
    a = da.from_array(np.zeros((200, 200)), chunks=(100, 100))
    b = da.from_array(np.zeros((200, 200)), chunks=(50, 200))
    c = a + b
    assert c.chunks == ((100, 100), (100, 100))              # golden case "nested_2d"
    v = da.from_array(np.zeros(200), chunks=50)
    assert (a + v).chunks == ((100, 100), (100, 100))        # "vector_broadcast": the vector is merged up
    def walk(e, acc):
        acc.append(e)
        for d in e.dependencies():
            walk(d, acc)
        return acc
    # host leaves absorb the inserted rechunks (reference test_lower_inserted_rechunk_pushes_into_from_array)...
    assert not {type(e).__name__ for e in walk(c.optimize().expr, [])} & {"TasksRechunk", "Rechunk"}
    assert {e.chunks for e in walk(c.optimize().expr, []) if type(e).__name__ == "FromArray"} == {((100, 100), (100, 100))}
    # ... opaque leaves get a real rechunk
    oa = da.from_host_blocks(lambda bid: None, (200, 200), (100, 100), np.float64, token="oa")
    ob = da.from_host_blocks(lambda bid: None, (200, 200), (50, 200), np.float64, token="ob")
    assert {type(e).__name__ for e in walk((oa + ob).optimize().expr, [])} & {"TasksRechunk"}
    # light coarse operand must not inflate the heavy fine one (golden "light_coarse_refused")
    heavy = da.from_array(np.zeros((1000, 64)), chunks=(10, 64))
    light = da.from_array(np.zeros((1000, 1)), chunks=(500, 1))
    assert (heavy * light).chunks == ((10,) * 100, (64,))
