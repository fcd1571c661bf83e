"""The reference's own reduction test matrix, mirrored: ``reduction_1d_test`` /
``reduction_2d_test`` / ``test_reductions_1D`` / ``test_reductions_2D`` / ``test_reductions_1D_nans``
/ ``test_arg_reductions`` / ``test_reduction_errors`` of
/root/reference/dask_array/tests/test_reductions.py:185-449, with the reference's ``assert_eq``
contract (dtype, shape, ``allclose(rtol=1e-5, atol=1e-8, equal_nan=True)``,
``_test_utils.py:26-37,122-250``).  Complex dtypes are outside the B200 kernels (not on the path)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def assert_eq(a, b):
    a = a.compute() if hasattr(a, "compute") else a
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-8, equal_nan=True), (a, b)


def reduction_1d_test(da_func, darr, np_func, narr, use_dtype=True, split_every=True):
    assert_eq(da_func(darr), np_func(narr))
    assert_eq(da_func(narr), np_func(narr))          # NumPy input
    assert_eq(da_func(darr, keepdims=True), np_func(narr, keepdims=True))
    assert_eq(da_func(darr, axis=()), np_func(narr, axis=()))
    assert da_func(darr).name == da_func(darr).name   # same_keys: deterministic names
    if use_dtype:
        assert_eq(da_func(darr, dtype="f8"), np_func(narr, dtype="f8"))
        assert_eq(da_func(darr, dtype="i8"), np_func(narr, dtype="i8"))
    if split_every:
        a1, a2 = da_func(darr, split_every=2), da_func(darr, split_every={0: 2})
        assert a1.name == a2.name
        assert_eq(a1, np_func(narr))
        assert_eq(a2, np_func(narr))
        assert_eq(da_func(darr, keepdims=True, split_every=2), np_func(narr, keepdims=True))


def reduction_2d_test(da_func, darr, np_func, narr, use_dtype=True, split_every=True):
    assert_eq(da_func(darr, keepdims=True), np_func(narr, keepdims=True))
    assert_eq(da_func(darr, axis=()), np_func(narr, axis=()))
    for ax in (0, 1, -1, -2, (1, 0)):
        assert_eq(da_func(darr, axis=ax), np_func(narr, axis=ax))
    assert_eq(da_func(darr, axis=1, keepdims=True), np_func(narr, axis=1, keepdims=True))
    assert_eq(da_func(darr, axis=(), keepdims=True), np_func(narr, axis=(), keepdims=True))
    assert da_func(darr, axis=1).name == da_func(darr, axis=1).name
    if use_dtype:
        assert_eq(da_func(darr, dtype="f8"), np_func(narr, dtype="f8"))
        assert_eq(da_func(darr, dtype="i8"), np_func(narr, dtype="i8"))
    if split_every:
        a1, a2 = da_func(darr, split_every=4), da_func(darr, split_every={0: 2, 1: 2})
        assert a1.name == a2.name
        assert_eq(a1, np_func(narr))
        assert_eq(a2, np_func(narr))
        assert_eq(da_func(darr, keepdims=True, split_every=4), np_func(narr, keepdims=True))
        assert_eq(da_func(darr, axis=0, split_every=2), np_func(narr, axis=0))
        assert_eq(da_func(darr, axis=0, keepdims=True, split_every=2), np_func(narr, axis=0, keepdims=True))
        assert_eq(da_func(darr, axis=1, split_every=2), np_func(narr, axis=1))
        assert_eq(da_func(darr, axis=1, keepdims=True, split_every=2), np_func(narr, axis=1, keepdims=True))


@pytest.mark.parametrize("dtype", ["f4", "i4"])
def test_reductions_1D(dtype):
    import dask_array_b200 as da
    x = np.arange(5).astype(dtype)
    a = da.from_array(x, chunks=(2,))
    for f, nf in ((da.sum, np.sum), (da.prod, np.prod), (da.mean, np.mean), (da.var, np.var), (da.std, np.std)):
        reduction_1d_test(f, a, nf, x)
    for f, nf in ((da.min, np.min), (da.max, np.max), (da.any, np.any), (da.all, np.all)):
        reduction_1d_test(f, a, nf, x, False)
    reduction_1d_test(da.nansum, a, np.nansum, x)
    reduction_1d_test(da.nanprod, a, np.nanprod, x)
    reduction_1d_test(da.nanmean, a, np.mean, x)
    reduction_1d_test(da.nanvar, a, np.var, x)
    reduction_1d_test(da.nanstd, a, np.std, x)
    reduction_1d_test(da.nanmin, a, np.nanmin, x, False)
    reduction_1d_test(da.nanmax, a, np.nanmax, x, False)


@pytest.mark.parametrize("x", [np.array([np.inf, np.nan, -np.inf, 2]), np.array([np.nan, np.nan, 3, 2])])
def test_reductions_1D_nans(x):
    import dask_array_b200 as da
    x = x.astype("f4")
    a = da.from_array(x, chunks=(1,))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        reduction_1d_test(da.nansum, a, np.nansum, x)
        reduction_1d_test(da.nanprod, a, np.nanprod, x)
        reduction_1d_test(da.nanmean, a, np.nanmean, x, False)
        reduction_1d_test(da.nanvar, a, np.nanvar, x, False)
        reduction_1d_test(da.nanstd, a, np.nanstd, x, False)
        reduction_1d_test(da.nanmin, a, np.nanmin, x, False)
        reduction_1d_test(da.nanmax, a, np.nanmax, x, False)


@pytest.mark.parametrize("dtype", ["f4", "i4"])
def test_reductions_2D(dtype):
    import dask_array_b200 as da
    x = np.arange(1, 122).reshape((11, 11)).astype(dtype)
    a = da.from_array(x, chunks=(4, 4))
    reduction_2d_test(da.sum, a, np.sum, x)
    reduction_2d_test(da.mean, a, np.mean, x)
    reduction_2d_test(da.var, a, np.var, x, False)
    reduction_2d_test(da.std, a, np.std, x, False)
    for f, nf in ((da.min, np.min), (da.max, np.max), (da.any, np.any), (da.all, np.all)):
        reduction_2d_test(f, a, nf, x, False)
    reduction_2d_test(da.nansum, a, np.nansum, x)
    reduction_2d_test(da.nanmean, a, np.mean, x)
    reduction_2d_test(da.nanvar, a, np.nanvar, x, False)
    reduction_2d_test(da.nanstd, a, np.nanstd, x, False)
    reduction_2d_test(da.nanmin, a, np.nanmin, x, False)
    reduction_2d_test(da.nanmax, a, np.nanmax, x, False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        reduction_2d_test(da.prod, a, np.prod, x)
        reduction_2d_test(da.nanprod, a, np.nanprod, x)


@pytest.mark.parametrize("dfunc,func", [("argmin", np.argmin), ("argmax", np.argmax), ("nanargmin", np.nanargmin), ("nanargmax", np.nanargmax)])
def test_arg_reductions(dfunc, func):
    """tests/test_reductions.py:362-397."""
    import dask_array_b200 as da
    x = np.random.default_rng(0).random((10, 10, 10))
    a = da.from_array(x, chunks=(3, 4, 5))
    f = getattr(da, dfunc)
    assert_eq(f(a), func(x))
    for ax in (0, 1, 2, -1):
        assert_eq(f(a, axis=ax), func(x, axis=ax))
        assert_eq(f(a, axis=ax, keepdims=True), func(x, axis=ax, keepdims=True))
    assert_eq(f(a, axis=1, split_every=2), func(x, axis=1))
    assert_eq(f(a, keepdims=True), func(x, keepdims=True))
    with pytest.raises(TypeError):
        f(a, axis=(1, 2))
    x2 = np.arange(10)
    a2 = da.from_array(x2, chunks=3)
    assert_eq(f(a2), func(x2))
    assert_eq(f(a2, axis=0), func(x2, axis=0))
    assert_eq(f(a2, axis=0, split_every=2), func(x2, axis=0))


def test_reduction_errors():
    import dask_array_b200 as da
    x = da.ones((5, 5), chunks=(3, 3))
    with pytest.raises(ValueError):
        x.sum(axis=2)
    with pytest.raises(ValueError):
        x.sum(axis=-3)


def test_chunk_structure_independence():
    """tests/test_reductions.py:1060-1079: results do not depend on the chunking."""
    import dask_array_b200 as da
    rng = np.random.default_rng(5)
    x = rng.random((17, 23))
    want = {k: getattr(np, k)(x, axis=0) for k in ("sum", "mean", "var", "max", "argmin")}
    for chunks in [(17, 23), (1, 23), (17, 1), (4, 5), (16, 22), (6, 3)]:
        a = da.from_array(x, chunks=chunks)
        for k, w in want.items():
            got = getattr(a, k)(axis=0).compute()
            if k in ("max", "argmin"):
                assert np.array_equal(got, w), (k, chunks)
            else:
                np.testing.assert_allclose(got, w, rtol=1e-12, err_msg=f"{k} {chunks}")
