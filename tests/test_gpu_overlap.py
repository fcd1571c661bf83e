"""Halo exchange, map_blocks and map_overlap against the reference's documented
answer (`_overlap.py:935-962`) and the oracle's whole-array restatement (`oracle.reference.overlap`).
Pure data movement: bit-exact."""
import numpy as np
import pytest

from oracle import reference as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def test_overlap_docstring_example_of_the_reference(da):
    x = np.arange(64).reshape((8, 8))
    d = da.from_array(x, chunks=(4, 4))
    g = da.overlap.overlap(d, depth={0: 2, 1: 1}, boundary={0: 100, 1: "reflect"})
    assert g.chunks == ((8, 8), (6, 6))
    rows = [[100] * 12] * 2 + [[r * 8 + c for c in (0, 0, 1, 2, 3, 4, 3, 4, 5, 6, 7, 7)] for r in (0, 1, 2, 3, 4, 5)]
    rows += [[r * 8 + c for c in (0, 0, 1, 2, 3, 4, 3, 4, 5, 6, 7, 7)] for r in (2, 3, 4, 5, 6, 7)] + [[100] * 12] * 2
    assert np.array_equal(g.compute(), np.array(rows))


@pytest.mark.parametrize("boundary", ["none", "periodic", "reflect", "nearest", 7])
@pytest.mark.parametrize("shape,chunks,depth", [((40, 36), (10, 12), {0: 2, 1: 3}), ((64,), (16,), {0: 5}),
                                                ((12, 20, 18), (6, 10, 9), {0: 1, 1: 0, 2: 2}),
                                                ((30, 30), (10, 30), {0: 4, 1: 2})])
def test_overlap_matches_oracle(da, boundary, shape, chunks, depth):
    rng = np.random.default_rng(1)
    xh = rng.integers(0, 1000, size=shape).astype(np.int32)
    x = da.from_array(xh, chunks=chunks).persist()
    bnd = {ax: boundary for ax in range(len(shape))}
    got = da.overlap.overlap(x, depth=depth, boundary=bnd)
    want = ref.overlap(ref.Blocked.from_array(xh, chunks), depth, bnd)
    assert got.chunks == want.chunks
    assert np.array_equal(got.compute(), want.to_array())
    back = da.overlap.trim_internal(got, depth, bnd)
    assert back.chunks == x.chunks and np.array_equal(back.compute(), xh)


def test_small_chunks_are_merged_to_hold_the_depth(da):
    xh = np.arange(47, dtype=np.float64)
    x = da.from_array(xh, chunks=((20, 20, 1, 6),)).persist()
    g = da.overlap.overlap(x, depth=10, boundary="none")
    assert g.chunks == ((30, 31, 26),)                   # ensure_minimum_chunksize(10, (20, 20, 1, 6)) = (20, 11, 16)
    with pytest.raises(ValueError):
        da.overlap.overlap(x, depth=10, boundary="none", allow_rechunk=False)


def test_map_blocks_and_map_overlap_stencils(da):
    rng = np.random.default_rng(2)
    xh = rng.random((96, 80))
    x = da.from_array(xh, chunks=(32, 20)).persist()
    y = x.map_blocks(lambda b: np.sqrt(b) * 2 + 1)
    assert y.dtype == np.float64 and np.array_equal(y.compute(), np.sqrt(xh) * 2 + 1)
    z = da.map_blocks(lambda a, b, k: a * k - b, x, y, 3.0)
    assert np.array_equal(z.compute(), xh * 3.0 - (np.sqrt(xh) * 2 + 1))
    # second difference along axis 0 with symmetric edges: the halo makes every block self-sufficient
    lap = da.overlap.overlap(x, depth={0: 1, 1: 0}, boundary={0: "reflect", 1: "none"}).map_blocks(
        lambda b: b[2:] + b[:-2] - 2 * b[1:-1], chunks=x.chunks)
    p = np.pad(xh, ((1, 1), (0, 0)), mode="symmetric")
    assert np.array_equal(lap.compute(), p[2:] + p[:-2] - 2 * p[1:-1])
    # map_overlap with trim: a shape-preserving function of the haloed block
    m = x.map_overlap(lambda b: b * 2, depth=2, boundary="periodic")
    assert m.chunks == x.chunks and np.array_equal(m.compute(), xh * 2)
    # 5-point stencil, periodic in both directions, computed on the interior of every haloed block
    def five(b):
        return b[1:-1, 1:-1] * -4 + b[2:, 1:-1] + b[:-2, 1:-1] + b[1:-1, 2:] + b[1:-1, :-2]
    s = da.overlap.overlap(x, depth=1, boundary="periodic").map_blocks(five, chunks=x.chunks)
    want = -4 * xh + np.roll(xh, 1, 0) + np.roll(xh, -1, 0) + np.roll(xh, 1, 1) + np.roll(xh, -1, 1)
    np.testing.assert_allclose(s.compute(), want, rtol=1e-13, atol=1e-13)
    with pytest.raises(TypeError):
        x.map_blocks(lambda b: b.to_numpy(), dtype=np.float64).compute()      # host results are refused, loudly


def test_compiled_replay_with_map_blocks(da):
    xh = np.random.default_rng(3).random((64, 64))
    x = da.from_array(xh, chunks=(32, 32)).persist()
    step = da.compile((x.map_overlap(lambda b: b + 1, depth=1, boundary="reflect") * 2).sum())
    step.run(); step.run()
    np.testing.assert_allclose(step.result(), ((xh + 1) * 2).sum(), rtol=1e-12)


@pytest.mark.parametrize("shape,chunks,window,axis", [((200,), (64,), 7, 0), ((60, 50), (20, 25), 5, 0),
                                                      ((60, 50), (20, 25), (3, 4), (0, 1)), ((30, 40, 8), (10, 40, 8), 6, 1),
                                                      ((40,), (3,), 5, 0)])
def test_sliding_window_view_and_rolling_reductions(da, shape, chunks, window, axis):
    """`sliding_window_view` (`_overlap.py:1365-1433`) against NumPy's; rolling sum / mean / max read the
    overlapping windows in place (no window is materialised)."""
    rng = np.random.default_rng(5)
    xh = rng.integers(-50, 50, size=shape).astype(np.float64)
    x = da.from_array(xh, chunks=chunks).persist()
    v = da.sliding_window_view(x, window, axis=axis)
    want = np.lib.stride_tricks.sliding_window_view(xh, window, axis=axis)
    assert v.shape == want.shape
    assert np.array_equal(v.compute(), want)
    nwin = len(window) if isinstance(window, tuple) else 1
    red = tuple(range(-nwin, 0))
    assert np.array_equal(v.sum(axis=red).compute(), want.sum(axis=red))               # integer-valued: exact
    assert np.array_equal(v.max(axis=red).compute(), want.max(axis=red))
    np.testing.assert_allclose(v.mean(axis=red).compute(), want.mean(axis=red), rtol=1e-12)
    step = da.compile(v.sum(axis=red))
    step.run(); step.run()
    assert np.array_equal(step.result(), want.sum(axis=red))
