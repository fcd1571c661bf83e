"""``Array.__array_function__`` (dask_array/_collection.py:866-923): NumPy functions applied to an Array dispatch to
the same-named function of the package and stay lazy; unknown functions warn and run on computed results (which
needs a GPU: here they must fail loudly, never compute on the host silently)."""
import warnings

import numpy as np
import pytest

import dask_array_b200 as da


def test_numpy_functions_stay_lazy():
    x = da.from_array(np.arange(24.0).reshape(4, 6), chunks=(2, 3))
    e = da.expand_dims(x, 0)
    checks = [
        (np.sum(x, axis=0), x.sum(axis=0)), (np.mean(x), x.mean()), (np.max(x, axis=1), x.max(axis=1)),
        (np.amax(x, axis=1), x.max(axis=1)), (np.clip(x, 1, 5), x.clip(1, 5)), (np.round(x, 2), da.round(x, 2)),
        (np.reshape(x, (2, 2, 6)), x.reshape((2, 2, 6))), (np.transpose(x), x.T), (np.swapaxes(x, 0, 1), x.swapaxes(0, 1)),
        (np.concatenate([x, x], axis=1), da.concatenate([x, x], axis=1)), (np.where(x > 3, x, 0), da.where(x > 3, x, 0)),
        (np.tensordot(x, x.T, axes=1), da.tensordot(x, x.T, axes=1)), (np.cumsum(x, axis=1), x.cumsum(axis=1)),
        (np.argmax(x, axis=0), x.argmax(axis=0)), (np.expand_dims(x, 0), e), (np.squeeze(e), da.squeeze(e)),
        (np.nansum(x), da.nansum(x)), (np.stack([x, x]), da.stack([x, x])), (np.broadcast_to(x, (2, 4, 6)), da.broadcast_to(x, (2, 4, 6))),
        (np.ravel(x), x.ravel()), (np.moveaxis(x, 0, 1), da.moveaxis(x, 0, 1)), (np.var(x, ddof=1), x.var(ddof=1)),
        (np.std(x, axis=0), x.std(axis=0)), (np.prod(x), x.prod()), (np.all(x), x.all()), (np.diff(x, axis=1), da.diff(x, axis=1)),
        (np.matmul(x, x.T), x @ x.T), (np.roll(x, 2, axis=1), da.roll(x, 2, axis=1)), (np.flip(x, 0), da.flip(x, 0)),
    ]
    for k, (got, want) in enumerate(checks):
        assert isinstance(got, da.Array), (k, type(got))
        assert got.name == want.name and got.shape == want.shape and got.dtype == want.dtype, k
    # ufuncs go through __array_ufunc__ as before
    assert isinstance(np.sin(x), da.Array) and np.add(x, 1).name == (x + 1).name


def test_unknown_numpy_functions_warn_and_need_the_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("box has a GPU: the computed-arguments path is exercised in tests/test_zz_late_gpu.py")
    x = da.ones((4, 4), chunks=2)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            np.linalg.norm(x)
    assert any(issubclass(i.category, FutureWarning) and "numpy.linalg.norm" in str(i.message) for i in w)
