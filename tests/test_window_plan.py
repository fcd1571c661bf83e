"""Host logic of the sliding-window reductions on CPU: the parent rewrite, the trimmed output chunks
(reductions/_sliding_window.py:431-446) and ``WindowHalo.pieces`` -- simulated with NumPy blocks and checked
against the oracle's restatement of the reference's overlap-plan result."""

import numpy as np
import pytest

import dask_array_b200 as da
from oracle import reference as ref

swv = np.lib.stride_tricks.sliding_window_view


def _simulate(expr, xh, chunks):
    """Run WindowReduce(WindowHalo(FromArray)) with NumPy: gather the halo blocks through ``pieces``, reduce them."""
    red = expr
    halo = red.operand("array")
    src = halo.operand("array")
    assert type(red).__name__ == "WindowReduce" and type(halo).__name__ == "WindowHalo"
    blocks = ref.Blocked.from_array(xh, chunks).blocks
    w, ax = red.operand("window"), red.operand("axis")
    out = {}
    for nbid in halo.block_ids():
        buf = np.full(halo.block_shape(nbid), np.nan)
        for obid, ssl, dsl in halo.pieces(nbid):
            buf[dsl] = blocks[obid][ssl]
        assert not np.isnan(buf).any()                       # the pieces tile the halo block exactly
        out[nbid] = getattr(np, red.operand("redop"))(swv(buf, w, axis=ax), axis=-1)
    return out, src


@pytest.mark.parametrize("shape,chunks,axis,w", [((20, 5), (3, 5), 0, 7), ((20, 5), (3, 2), 0, 7), ((9, 40), (9, 6), 1, 20),
                                                  ((33, 4, 10), (8, 4, 3), 2, 5), ((100,), (7,), 0, 30), ((12,), (12,), 0, 12)])
def test_window_halo_pieces_and_chunks(shape, chunks, axis, w):
    xh = np.random.default_rng(0).random(shape)
    x = da.from_array(xh, chunks=chunks)
    r = da.sliding_window_view(x, w, axis=axis).sum(axis=-1)
    opt = r.expr.optimize()
    got_blocks, _ = _simulate(opt, xh, x.chunks)
    want = ref.da_sliding_window_reduce(ref.Blocked.from_array(xh, x.chunks), w, axis, "sum")
    assert opt.chunks == want.chunks                                  # trimmed like the reference (:431-446)
    assert set(got_blocks) == set(want.blocks)
    for bid, blk in want.blocks.items():
        np.testing.assert_allclose(got_blocks[bid], blk, rtol=1e-12)
    np.testing.assert_allclose(want.to_array(), swv(xh, w, axis=axis).sum(axis=-1), rtol=1e-12)


def test_rewrite_rules():
    x = da.from_array(np.zeros((30, 8)), chunks=(4, 8))
    v = da.sliding_window_view(x, 6, axis=0)
    assert type(v.expr).__name__ == "SlidingWindowView" and v.shape == (25, 8, 6)
    for kind in ("sum", "prod", "min", "max", "mean", "any", "all"):
        assert type(getattr(v, kind)(axis=-1).expr.optimize()).__name__ in ("WindowReduce", "FusedBlockwise")
        assert "WindowReduce" in getattr(v, kind)(axis=-1).expr.optimize().tree_repr()
    assert "WindowReduce" not in v.var(axis=-1).expr.optimize().tree_repr()          # no native kernel: generic plan
    assert "WindowReduce" not in v.sum(axis=0).expr.optimize().tree_repr()           # not the window axis
    v2 = da.sliding_window_view(x, (3, 2), axis=(0, 1))
    assert "WindowReduce" not in v2.sum(axis=(-1, -2)).expr.optimize().tree_repr()  # two window axes
    k = v.max(axis=-1, keepdims=True)
    assert k.expr.optimize().chunks[-1] == (1,) and k.expr.optimize().shape == (25, 8, 1)
    # int32 sum accumulates in int64 (np.sum dtype rule); bool any -> bool
    xi = da.from_array(np.zeros((30, 8), dtype=np.int32), chunks=(4, 8))
    assert da.sliding_window_view(xi, 6, axis=0).sum(axis=-1).dtype == np.int64
    assert da.sliding_window_view(xi, 6, axis=0).any(axis=-1).dtype == np.bool_
    with pytest.raises(ValueError):
        da.sliding_window_view(x, 31, axis=0)


def test_moving_window_keeps_the_input_chunks():
    """MovingWindowReduction has the input's shape AND chunks (reductions/_sliding_window.py:253): the front pad is
    moved onto the last block before the window kernel, or -- chunks shorter than the window -- the result is
    re-blocked."""
    import dask_array_b200 as da

    for shape, chunks, axis, w in (((40, 64), (20, 32), 1, 4), ((60, 24), (7, 24), 0, 10), ((200,), (33,), 0, 40),
                                   ((16, 90), (16, 11), 1, 25), ((8, 8), (8, 8), 0, 1)):
        x = da.from_array(np.zeros(shape), chunks=chunks)
        for red in ("move_sum", "move_mean", "move_min", "move_max"):
            y = getattr(da, red)(x, w, axis=axis)
            assert y.chunks == x.chunks and y.shape == x.shape and y.dtype == np.float64
    x = da.from_array(np.zeros((40, 64), dtype="f4"), chunks=(20, 32))
    tree = da.move_sum(x, 4, axis=1).expr.optimize().tree_repr()
    assert tree.count("TasksRechunk") == 2 and "(32, 35)" in tree     # values + counts: ONE re-block each, before the halo
