"""SURVEY 8(f) widening: structural views (expand_dims / squeeze / broadcast_to / concatenate /
stack) and the NaN-aware reducers, against NumPy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def test_structural_views(da):
    rng = np.random.default_rng(0)
    ah = rng.integers(0, 100, (12, 10)).astype(np.int64)
    bh = rng.integers(0, 100, (7, 10)).astype(np.int64)
    a, b = da.from_array(ah, chunks=(5, 4)), da.from_array(bh, chunks=(3, 4))
    assert np.array_equal(da.concatenate([a, b], axis=0).compute(), np.concatenate([ah, bh], axis=0))
    assert np.array_equal((da.concatenate([a, a * 2], axis=1) + 1).sum(axis=0).compute(),
                          (np.concatenate([ah, ah * 2], axis=1) + 1).sum(axis=0))
    assert np.array_equal(da.stack([a, a + 1], axis=0).compute(), np.stack([ah, ah + 1], axis=0))
    assert np.array_equal(da.stack([a, a + 1], axis=2).max(axis=2).compute(), np.stack([ah, ah + 1], axis=2).max(axis=2))
    e = da.expand_dims(a, 1)
    assert e.shape == (12, 1, 10) and np.array_equal(e.compute(), ah[:, None, :])
    assert np.array_equal((e * da.expand_dims(b, 0)[:, :7, :]).compute(), ah[:, None, :] * bh[None, :, :])
    assert np.array_equal(da.squeeze(e, 1).compute(), ah) and np.array_equal(e.squeeze().compute(), ah)
    assert np.array_equal(a.sum(axis=0, keepdims=True).squeeze().compute(), ah.sum(axis=0))
    bt = da.broadcast_to(da.from_array(ah[:1], chunks=(1, 4)), (6, 10))
    assert np.array_equal((bt + 0).compute(), np.broadcast_to(ah[:1], (6, 10)))
    assert np.array_equal(da.broadcast_to(a, (3, 12, 10)).sum(axis=0).compute(), 3 * ah)
    th, uh = rng.random((12, 40)).astype(np.float32), rng.random((7, 40)).astype(np.float32)
    t, u = da.from_array(th, chunks=(5, 16)), da.from_array(uh, chunks=(3, 16))      # ragged k blocks: 16, 16, 8
    mm = da.tensordot(t, u, axes=((1,), (1,)))
    np.testing.assert_allclose(mm.compute(), th.astype(np.float64) @ uh.T.astype(np.float64), rtol=1e-5)
    # contraction chunks that are not multiples of 8 elements cannot be TMA operands: fp32 then runs in IEEE
    # arithmetic on the exact kernel (b2_gemm_tn_simt); bf16 has no such path and says so
    got = (da.from_array(th[:, :10], chunks=(5, 4)) @ da.from_array(uh[:, :10], chunks=(3, 4)).T).compute()
    np.testing.assert_allclose(got, th[:, :10].astype(np.float64) @ uh[:, :10].T.astype(np.float64), rtol=1e-5)
    import ml_dtypes
    tb, ub = th[:, :10].astype(ml_dtypes.bfloat16), uh[:, :10].astype(ml_dtypes.bfloat16)
    with pytest.raises(NotImplementedError, match="multiples of 8"):
        (da.from_array(tb, chunks=(5, 4)) @ da.from_array(ub, chunks=(3, 4)).T).compute()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_nan_reducers(da, dtype):
    rng = np.random.default_rng(1)
    xh = rng.random((90, 70)).astype(dtype) * 10 + 1      # positive: north_star's rtol applies to the result
    xh[rng.random(xh.shape) < 0.1] = np.nan
    xh[5, :] = np.nan                              # an all-NaN row
    x = da.from_array(xh, chunks=(32, 25))
    rtol = 1e-5 if dtype == "float32" else 1e-12
    with np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        for axis in (None, 0, 1):
            np.testing.assert_allclose(da.nansum(x, axis=axis).compute(), np.nansum(xh, axis=axis), rtol=rtol)
            np.testing.assert_allclose(da.nanmean(x, axis=axis).compute(), np.nanmean(xh, axis=axis), rtol=rtol, equal_nan=True)
            np.testing.assert_allclose(da.nanvar(x, axis=axis, ddof=1).compute(), np.nanvar(xh, axis=axis, ddof=1), rtol=rtol, equal_nan=True)
            np.testing.assert_allclose(da.nanstd(x, axis=axis).compute(), np.nanstd(xh, axis=axis), rtol=rtol, equal_nan=True)
            assert np.array_equal(da.nanmin(x, axis=axis).compute(), np.nanmin(xh, axis=axis), equal_nan=True)
            assert np.array_equal(da.nanmax(x, axis=axis).compute(), np.nanmax(xh, axis=axis), equal_nan=True)
        ok = ~np.all(np.isnan(xh), axis=1)
        assert np.array_equal(da.nanargmax(x, axis=1).compute()[ok], np.nanargmax(xh[ok], axis=1))
        assert np.array_equal(da.nanargmin(x, axis=0).compute(), np.nanargmin(xh, axis=0))
    ih = rng.integers(-9, 9, (40, 30)).astype(np.int32)
    i = da.from_array(ih, chunks=(16, 16))
    assert da.nansum(i).compute() == ih.sum() and np.array_equal(da.nanmax(i, axis=0).compute(), ih.max(axis=0))
    assert np.array_equal(i.prod(axis=1).compute(), ih.prod(axis=1)) and i.any().compute() == ih.any()
    assert np.array_equal((i > 0).all(axis=0).compute(), (ih > 0).all(axis=0))


def test_elementwise_over_mismatched_chunks_unifies_like_the_reference():
    """operands on different block grids are rechunked to the reference's unified
    layout (tests/golden/unify.json) and fused; values bit-exact."""
    import dask_array_b200 as da
    rng = np.random.default_rng(11)
    ah, bh, vh = rng.random((1000, 700)), rng.random((1000, 700)), rng.random(700)
    a = da.from_array(ah, chunks=((300, 300, 300, 100), (256, 256, 188)))
    b = da.from_array(bh, chunks=((600, 400), (700,)))
    v = da.from_array(vh, chunks=100)
    y = a * 2 + b - v
    assert y.chunks[0] == (600, 400)                       # golden case "ragged_nested"
    assert np.array_equal(y.compute(), ah * 2 + bh - vh)
    xh = rng.integers(0, 100, size=400).astype(np.int64)
    x = da.from_array(xh, chunks=100)
    shifted = da.concatenate([x[350:], x[:350]])            # the roll pattern: (50, 100, 100, 100, 50)
    z = x + shifted
    assert z.chunks == ((100,) * 4,)                        # golden case "roll_shift"
    assert np.array_equal(z.compute(), xh + np.roll(xh, 50))


def test_roll_diff_flip_compositions():
    """`roll` (manipulation/_roll.py), `diff` (routines/_diff.py), `flip*` (manipulation/_flip.py): the
    reference composes them from slices / concatenate / subtract, and so does the B200 package."""
    import dask_array_b200 as da
    rng = np.random.default_rng(12)
    xh = rng.integers(-100, 100, size=(37, 50)).astype(np.int64)
    x = da.from_array(xh, chunks=(10, 16)).persist()
    for shift, axis in [(3, 1), (-5, 0), (50, 1), ((2, -7), (0, 1)), (0, 0)]:
        assert np.array_equal(da.roll(x, shift, axis=axis).compute(), np.roll(xh, shift, axis=axis)), (shift, axis)
    v = da.from_array(xh[0], chunks=7).persist()
    assert np.array_equal(da.roll(v, 11).compute(), np.roll(xh[0], 11))
    for n, axis in [(1, -1), (2, 0), (3, 1)]:
        assert np.array_equal(da.diff(x, n, axis=axis).compute(), np.diff(xh, n, axis=axis)), (n, axis)
    assert np.array_equal(da.diff(x, axis=1, prepend=0, append=7).compute(), np.diff(xh, axis=1, prepend=0, append=7))
    assert np.array_equal(da.flip(x).compute(), np.flip(xh))
    assert np.array_equal(da.flipud(x).compute(), np.flipud(xh)) and np.array_equal(da.fliplr(x).compute(), np.fliplr(xh))
    assert (da.diff(x, axis=0) ** 2).sum().compute() == (np.diff(xh, axis=0) ** 2).sum()
