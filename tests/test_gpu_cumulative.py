"""cumsum / cumprod / nancumsum / nancumprod on the GPU against the oracle
(`oracle.reference.da_cumulative`, pinned to the reference's own task graph by
tests/golden/cumulative.npz) -- the reference's cases: tests/test_reductions.py cumulative tests
(axis sweep, dtypes, nan variants, 1-D and N-d, ragged chunks).

Integers (wrap-around arithmetic is associative): bit-exact.  Floats: the scan order inside a block
differs (warp scans / block totals first), so |got - want| <= rtol * cumsum(|x|) with rtol 1e-5 (fp32)
and 1e-12 (fp64)."""
import ast

import numpy as np
import pytest

from oracle import reference as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def _check(got, want, xh, axis, kind):
    assert got.dtype == want.dtype and got.shape == want.shape
    if want.dtype.kind in "iu":
        assert np.array_equal(got, want)
        return
    rtol = 1e-5 if want.dtype == np.float32 else 1e-12
    if kind == "cumsum":
        bound = np.cumsum(np.abs(np.nan_to_num(xh.astype(np.float64))), axis=axis) * rtol + 1e-30
        assert np.all(np.abs(got.astype(np.float64) - want.astype(np.float64)) <= bound)
    else:
        np.testing.assert_allclose(got, want, rtol=rtol * 50)


def test_golden_cases(da):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cumulative.npz"))
    for case in sorted({k.split("/")[0] for k in g.files}):
        chunks, axis, kind, nan = ast.literal_eval(str(g[case + "/meta"][0]))
        xh, want = g[case + "/x"], g[case + "/result"]
        x = da.from_array(xh, chunks=chunks)
        fn = getattr(da, ("nan" if nan else "") + kind)
        got = fn(x, axis=axis).compute()
        _check(got, want, xh, axis, kind)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64, np.uint8, np.bool_])
@pytest.mark.parametrize("shape,chunks,axis", [
    ((1000, 700), (300, 256), 0), ((1000, 700), (300, 256), 1), ((513, 129), (128, 64), -1),
    ((64, 48, 40), (32, 16, 40), 1), ((64, 48, 40), (64, 48, 8), 2), ((20, 8, 6, 12), (5, 8, 3, 12), 0),
    ((100000,), (30000,), 0), ((4097,), (4097,), 0), ((7,), (3,), 0),
])
def test_cumsum_matrix(da, dtype, shape, chunks, axis):
    rng = np.random.default_rng(hash((shape, axis)) % 2**32)
    if np.dtype(dtype).kind == "f":
        xh = rng.standard_normal(shape).astype(dtype)
    elif dtype == np.bool_:
        xh = rng.random(shape) < 0.3
    else:
        xh = rng.integers(0, 100, size=shape).astype(dtype)
    want = ref.da_cumulative(ref.Blocked.from_array(xh, chunks), axis % len(shape), "cumsum").to_array()
    got = da.from_array(xh, chunks=chunks).cumsum(axis=axis).compute()
    _check(got, want, xh, axis, "cumsum")
    assert np.array_equal(want, np.cumsum(xh, axis=axis)) or want.dtype.kind == "f"


@pytest.mark.parametrize("dtype", [np.float64, np.int64, np.float32])
def test_cumprod_and_nan_variants(da, dtype):
    rng = np.random.default_rng(3)
    shape, chunks = (96, 80), (32, 20)
    xh = (rng.random(shape) * 0.5 + 0.75).astype(dtype) if np.dtype(dtype).kind == "f" else \
        rng.integers(1, 3, size=shape).astype(dtype)
    for axis in (0, 1):
        want = ref.da_cumulative(ref.Blocked.from_array(xh, chunks), axis, "cumprod").to_array()
        got = da.cumprod(da.from_array(xh, chunks=chunks), axis=axis).compute()
        _check(got, want, xh, axis, "cumprod")
    if np.dtype(dtype).kind == "f":
        xn = xh.copy()
        xn[rng.random(shape) < 0.2] = np.nan
        for axis in (0, 1):
            for kind in ("cumsum", "cumprod"):
                want = ref.da_cumulative(ref.Blocked.from_array(xn, chunks), axis, kind, nan=True).to_array()
                got = getattr(da, "nan" + kind)(da.from_array(xn, chunks=chunks), axis=axis).compute()
                _check(got, want, xn, axis, kind)
        # plain cumsum propagates NaN from its first occurrence on, like NumPy
        got = da.from_array(xn, chunks=chunks).cumsum(axis=1).compute()
        assert np.array_equal(np.isnan(got), np.isnan(np.cumsum(xn, axis=1)))


def test_long_vector_two_level_carries(da):
    """> 16 * 4096 segments: the scan of the segment totals recurses."""
    n = 4096 * 70000 // 1000 * 1000 // 100         # ~2.8 M elements, ragged blocks
    xh = np.random.default_rng(0).integers(-3, 4, size=n).astype(np.int32)
    got = da.from_array(xh, chunks=(n // 3 + 1,)).cumsum().compute()
    assert got.dtype == np.int64 and np.array_equal(got, np.cumsum(xh))
    from dask_array_b200 import _executor
    old = _executor._Cum.SEG
    _executor._Cum.SEG = 8                         # force the recursive path on a small input
    try:
        xs = np.random.default_rng(1).integers(-3, 4, size=8 * 16 * 8 * 5 + 3).astype(np.int64)
        assert np.array_equal(da.from_array(xs, chunks=(1000,)).cumsum().compute(), np.cumsum(xs))
    finally:
        _executor._Cum.SEG = old


@pytest.mark.parametrize("shape,chunks", [((5, 4, 3), (2, 2, 3)), ((37, 22), (8, 5)), ((6, 1, 9), (4, 1, 2))])
def test_cumsum_axis_none_nd(da, shape, chunks):
    """``cumsum(axis=None)`` / ``cumprod`` of an N-d array (``_prepare_cumulative``, _cumulative.py:77-97): the trailing
    axes are rechunked to one chunk, every block is viewed flat, the vector is scanned; also ``ravel`` on its own."""
    rng = np.random.default_rng(4)
    xh = rng.integers(-5, 6, size=shape).astype(np.int64)
    x = da.from_array(xh, chunks=chunks)
    assert np.array_equal(x.ravel().compute(), xh.ravel())
    assert np.array_equal(x.cumsum().compute(), np.cumsum(xh))
    fh = 1.0 + rng.random(shape) * 1e-3
    np.testing.assert_allclose(da.cumprod(da.from_array(fh, chunks=chunks), axis=None).compute(), np.cumprod(fh), rtol=1e-12)
    assert np.array_equal((x.T * 2).ravel().compute(), (xh.T * 2).ravel())


def test_errors(da):
    x = da.from_array(np.zeros((4, 4)), chunks=2)
    assert x.cumsum().shape == (16,)               # axis=None on N-d: flattened first (test_cumsum_axis_none_nd)
    with pytest.raises(ValueError):
        x.cumsum(axis=0, method="nope")
    with pytest.raises(np.exceptions.AxisError):
        x.cumsum(axis=3)
