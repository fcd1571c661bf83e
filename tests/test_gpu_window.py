"""Sliding-window reductions (reductions/_sliding_window.py:405-560, _overlap.py:500-566):
``sliding_window_view(x, w, axis).<reducer>(axis=-1)`` runs as WindowHalo + b2_window_reduce on the input's own
chunks.  Against NumPy's sliding_window_view reductions: integers / bool / min / max bit-exact, float sums
within rtol 1e-5 / 1e-12 of |x| summed over the window.  Cases mirror the reference's
tests/test_sliding_window_reduction.py shapes: chunks smaller than the window, windows spanning several
blocks, trailing blocks that emit nothing, N-d arrays, both axes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
swv = np.lib.stride_tricks.sliding_window_view


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def _close(got, want, xabs, w, axis, dtype):
    if np.dtype(dtype).kind != "f":
        assert np.array_equal(got, want)
        return
    rtol = 1e-5 if np.dtype(dtype) == np.float32 else 1e-12
    bound = swv(xabs, w, axis=axis).sum(axis=-1) * rtol + 1e-30
    assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= bound)


@pytest.mark.parametrize("dtype", ["float32", "float64", "int32", "int64"])
@pytest.mark.parametrize("shape,chunks,axis,w", [
    ((200, 96), (30, 40), 0, 7), ((200, 96), (30, 40), 1, 7), ((200, 96), (3, 96), 0, 50),       # chunks << window
    ((64, 300), (64, 17), 1, 120), ((33, 5, 40), (8, 5, 13), 0, 4), ((33, 5, 40), (8, 5, 13), 2, 21),
    ((1000,), (128,), 0, 300), ((1000,), (1000,), 0, 1), ((50, 70), (50, 70), 1, 70)])
def test_window_sum_min_max(da, dtype, shape, chunks, axis, w):
    rng = np.random.default_rng(hash((shape, axis, w)) % 2**32)
    xh = (rng.random(shape) * 200 - 100).astype(dtype)
    x = da.from_array(xh, chunks=chunks)
    v = da.sliding_window_view(x, w, axis=axis)
    xabs = np.abs(xh.astype(np.float64))
    r = v.sum(axis=-1)
    assert type(r.expr.optimize()).__name__ == "WindowReduce"
    want = swv(xh, w, axis=axis).sum(axis=-1)
    got = r.compute()
    assert got.dtype == want.dtype and got.shape == want.shape
    _close(got, want, xabs, w, axis, dtype)
    assert np.array_equal(v.max(axis=-1).compute(), swv(xh, w, axis=axis).max(axis=-1))
    assert np.array_equal(v.min(axis=-1, keepdims=True).compute(), swv(xh, w, axis=axis).min(axis=-1, keepdims=True))
    if np.dtype(dtype).kind == "f":
        gm = v.mean(axis=-1).compute()
        wm = swv(xh, w, axis=axis).mean(axis=-1)
        assert gm.dtype == wm.dtype
        bound = swv(xabs, w, axis=axis).mean(axis=-1) * (1e-5 if dtype == "float32" else 1e-12) + 1e-30
        assert np.all(np.abs(gm.astype(np.float64) - wm.astype(np.float64)) <= bound)


def test_window_bool_prod_nan_and_fallbacks(da):
    rng = np.random.default_rng(3)
    bh = rng.random((90, 40)) < 0.1
    b = da.from_array(bh, chunks=(16, 40))
    v = da.sliding_window_view(b, 9, axis=0)
    assert np.array_equal(v.any(axis=-1).compute(), swv(bh, 9, axis=0).any(axis=-1))
    assert np.array_equal(v.all(axis=-1).compute(), swv(bh, 9, axis=0).all(axis=-1))
    assert np.array_equal(v.sum(axis=-1).compute(), swv(bh, 9, axis=0).sum(axis=-1))           # bool -> int64 counts
    ph = 1.0 + rng.random((40, 64)) * 0.01
    p = da.sliding_window_view(da.from_array(ph, chunks=(40, 10)), 25, axis=1)
    np.testing.assert_allclose(p.prod(axis=-1).compute(), swv(ph, 25, axis=1).prod(axis=-1), rtol=1e-12)
    nh = rng.random((60, 30))
    nh[rng.random(nh.shape) < 0.2] = np.nan
    n = da.sliding_window_view(da.from_array(nh, chunks=(7, 30)), 11, axis=0)
    np.testing.assert_allclose(da.nansum(n, axis=-1).compute(), np.nansum(swv(nh, 11, axis=0), axis=-1), rtol=1e-12)
    # max / min propagate NaN like np.maximum / np.minimum
    assert np.array_equal(n.max(axis=-1).compute(), swv(nh, 11, axis=0).max(axis=-1), equal_nan=True)
    # not a window-axis reduction, or a reducer without a native kernel: the generic plan (window views) runs
    g = n.sum(axis=0)
    assert type(g.expr.optimize()).__name__ != "WindowReduce"
    np.testing.assert_allclose(da.nansum(n, axis=0).compute(), np.nansum(swv(nh, 11, axis=0), axis=0), rtol=1e-12)
    np.testing.assert_allclose(n.std(axis=-1).compute()[~np.isnan(swv(nh, 11, axis=0).std(axis=-1))],
                               swv(nh, 11, axis=0).std(axis=-1)[~np.isnan(swv(nh, 11, axis=0).std(axis=-1))], rtol=1e-10)
    # the view on its own, and two window axes
    xh = rng.random((20, 18))
    x = da.from_array(xh, chunks=(6, 7))
    assert np.array_equal(da.sliding_window_view(x, 4, axis=1).compute(), swv(xh, 4, axis=1))
    assert np.array_equal(da.sliding_window_view(x, (3, 2), axis=(0, 1)).max(axis=(-1, -2)).compute(),
                          swv(xh, (3, 2), axis=(0, 1)).max(axis=(-1, -2)))


def test_window_kernel_through_the_c_abi(da):
    """b2_window_reduce_batched directly (ctypes): both orientations, a window that spans many segments, two jobs."""
    import ctypes as C
    from dask_array_b200 import DeviceChunk, _lib
    from dask_array_b200._device import alloc_bytes, current_stream_ptr
    rng = np.random.default_rng(5)
    for shapes, along, w in (([(2, 300, 70), (1, 41, 70)], 0, 37), ([(1, 50, 1000)], 1, 256), ([(1, 9, 4100), (1, 3, 1700)], 1, 1500)):
        hosts = [rng.integers(-50, 50, sh).astype(np.int64) for sh in shapes]
        srcs = [DeviceChunk.from_numpy(h) for h in hosts]
        outs, jobs = [], []
        for (B, R, Cc), src in zip(shapes, srcs):
            out = DeviceChunk.empty((B, R - w + 1, Cc) if not along else (B, R, Cc - w + 1), np.int64)
            j = _lib.WindowJob()
            j.src, j.dst, j.B, j.R, j.C = src.ptr, out.ptr, B, R, Cc
            outs.append(out)
            jobs.append(j)
        arr = (_lib.WindowJob * len(jobs))(*jobs)
        d_jobs = alloc_bytes(C.sizeof(arr))
        _lib.check(_lib.lib.b2_window_reduce_batched(_lib.RED_SUM, _lib.dtype_code("int64"), arr, len(jobs), d_jobs.data_ptr(),
                                                     w, along, 0, current_stream_ptr()))
        for h, out in zip(hosts, outs):
            assert np.array_equal(out.to_numpy(), swv(h, w, axis=2 if along else 1).sum(axis=-1))


def _move_ref(x, w, reducer, min_count, axis):
    """bottleneck move_* semantics in plain NumPy: trailing window, NaN-skipping, NaN below min_count."""
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, -1)
    n = x.shape[-1]
    out = np.full(x.shape, np.nan)
    limit = w if min_count is None else min_count
    fn = {"move_sum": np.nansum, "move_mean": np.nanmean, "move_min": np.nanmin, "move_max": np.nanmax}[reducer]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for t in range(n):
            win = x[..., max(0, t - w + 1):t + 1]
            cnt = (~np.isnan(win)).sum(axis=-1)
            val = np.where(cnt > 0, fn(np.where(cnt[..., None] > 0, win, 0.0) if reducer in ("move_min", "move_max") else win, axis=-1), np.nan)
            out[..., t] = np.where(cnt >= limit, val, np.nan)
    return np.moveaxis(out, -1, axis)


@pytest.mark.parametrize("reducer", ["move_sum", "move_mean", "move_min", "move_max"])
@pytest.mark.parametrize("shape,chunks,axis,w,min_count", [((60, 24), (7, 24), 0, 10, None), ((60, 24), (7, 24), 0, 10, 3),
                                                          ((16, 90), (16, 11), 1, 25, 1), ((200,), (33,), 0, 40, 20),
                                                          ((40, 64), (20, 32), 1, 4, 2), ((96, 8), (32, 8), 0, 33, None)])
def test_moving_window_trailing_nan_skipping(da, reducer, shape, chunks, axis, w, min_count):
    """MovingWindowReduction (reductions/_sliding_window.py:183-246, 249-400): chunks smaller than the window,
    windows clipped at the array start, NaNs skipped, min_count."""
    rng = np.random.default_rng(7)
    xh = rng.random(shape) * 10 - 5
    xh[rng.random(shape) < 0.15] = np.nan
    x = da.from_array(xh, chunks=chunks)
    y = getattr(da, reducer)(x, w, min_count=min_count, axis=axis)
    assert y.chunks == x.chunks                       # "same shape and chunks as the input" (:253)
    got = y.compute()
    want = _move_ref(xh, w, reducer, min_count, axis)
    assert got.shape == want.shape and got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, equal_nan=True)
