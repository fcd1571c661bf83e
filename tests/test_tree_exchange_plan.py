"""Pure (CPU) planning of the tree-reduction exchange across ranks (reductions/_reduction.py:751-806):
which PartialReduce levels every owner folds locally, and the slab layout of the partial blocks that doubles as
the send buffer of the peer-memory all-gather -- identical on every rank by construction."""
import numpy as np
import pytest

import dask_array_b200 as da
from dask_array_b200._exchange import level_is_local, mean_count, owner_of, partials_layout
from dask_array_b200._reductions import PartialReduce


def _levels(expr):
    out = []
    while isinstance(expr, PartialReduce) or expr.dependencies():
        if isinstance(expr, PartialReduce):
            out.append(expr)
        deps = expr.dependencies()
        if not deps:
            break
        expr = deps[0]
    return out


@pytest.mark.parametrize("W", [2, 4, 8])
def test_c2_tree_levels_local_or_exchanged(W):
    """BASELINE config 2 dealt over W GPUs: mean(axis=0) needs no exchange at all (owner = block column), the
    std() tree needs exactly one (its first level mixes block columns)."""
    x = da.random.default_rng(0).random((32768, 32768), dtype=np.float32, chunks=(4096, 4096))
    y = da.sin(x) * 2 + x**2
    (agg,) = _levels(y.mean(axis=0).expr.optimize())
    assert agg.operand("final") and level_is_local(agg, W)
    lv = _levels(y.std().expr.optimize())
    assert [l.operand("final") for l in lv] == [True, False]
    assert not level_is_local(lv[1], W)                      # chunk partials -> first combine level: all-gather
    # c3: argmax(axis=1) of row panels: the reduced axis lies inside a block, every level is local
    z = da.from_array(np.zeros((64, 16)), chunks=(8, 16))
    assert all(level_is_local(l, W) for l in _levels(z.argmax(axis=1).expr.optimize()))
    # weak-scaling variant: (32768, 32768 * W)
    xw = da.random.default_rng(0).random((32768, 4096 * 8 * W), dtype=np.float32, chunks=(4096, 4096))
    (aggw,) = _levels((xw * 2).mean(axis=0).expr.optimize())
    assert level_is_local(aggw, W)


@pytest.mark.parametrize("W", [1, 2, 3, 8])
def test_partials_layout_is_a_partition_and_counts_add_up(W):
    x = da.from_array(np.zeros((50, 30)), chunks=(10, 7))
    lv = _levels(x.mean().expr.optimize())
    chunk_step = lv[-1].operand("array")
    layout, sizes = partials_layout(chunk_step, "mean", W)
    seen = set()
    for r in range(W):
        spans = []
        for bid, ent in layout[r].items():
            assert owner_of(chunk_step, bid, W) == r and bid not in seen
            seen.add(bid)
            for name, off, shp, dt in ent:
                assert off % 16 == 0 and name == "total" and dt == np.float64
                spans.append((off, off + int(np.prod(shp)) * dt.itemsize))
        spans.sort()
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and (not spans or spans[-1][1] <= sizes[r])
    assert seen == set(chunk_step.block_ids())
    # element counts behind the totals: per chunk block, and summed through the levels up to the whole array
    assert sum(mean_count(chunk_step, b) for b in chunk_step.block_ids()) == 50 * 30
    top = lv[0]
    assert sum(mean_count(top.operand("array"), b) for b in top.operand("array").block_ids()) == 50 * 30
    lay_m, _ = partials_layout(chunk_step, "moment", W)
    assert all(ent[0][2][-1] == 3 and ent[0][3] == np.float64 for r in range(W) for ent in lay_m[r].values())
