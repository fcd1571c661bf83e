"""GPU tests written AFTER the round's GPU budget was spent: they have not been run on a B200 yet (see
profiles/r2_pytest_gpu_late.md).  The file name sorts last on purpose, so that ``pytest -x`` reaches every
validated test first.  Everything here goes through paths the validated tests already exercise (HostBlocks
staging, fused element-wise kernels, transposes); the CPU suite checks the same operators' dtypes and that their
kernels compile (tests/test_ufunc_vocabulary.py), and the block contents of arange / linspace
(tests/test_normalize_chunks.py).
"""
import numpy as np
import pytest

# Non-strict xfail: these tests have never met a GPU, so an unexpected mismatch here must not mask the validated
# suite (the driver runs ``pytest -x``).  They are reported as XPASS when they pass and xfailed when they do not --
# ``pytest -rxX tests/test_zz_late_gpu.py -m gpu`` lists both; remove the marker once they have been run.
pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="written after the GPU budget was spent: not yet run on a B200")]

F32 = dict(rtol=1e-5, atol=1e-6)          # north-star tolerance, fp32
F64 = dict(rtol=1e-12, atol=1e-13)        # north-star tolerance, fp64


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def _close(got, want, dtype):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype, (got.dtype, want.dtype)
    if want.dtype.kind in "biu":
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, equal_nan=True, **(F32 if np.dtype(dtype) == np.float32 else F64))


# operator -> (lo, hi) of the inputs; everything else draws from (-3, 3)
_DOMAIN = {"log": (0.05, 9), "log2": (0.05, 9), "log10": (0.05, 9), "log1p": (-0.9, 9), "sqrt": (0, 9), "arcsin": (-1, 1),
           "arccos": (-1, 1), "arccosh": (1, 9), "arctanh": (-0.99, 0.99), "tan": (-1.3, 1.3), "reciprocal": (0.2, 3),
           "power": (0.1, 3), "float_power": (0.1, 3), "exp": (-5, 5), "exp2": (-5, 5), "expm1": (-5, 5),
           "sinh": (-5, 5), "cosh": (-5, 5)}
_UNARY = ["negative", "positive", "exp", "exp2", "log", "log2", "log10", "log1p", "expm1", "sqrt", "square", "cbrt",
          "reciprocal", "sin", "cos", "tan", "arcsin", "arccos", "arctan", "sinh", "cosh", "tanh", "arcsinh", "arccosh",
          "arctanh", "deg2rad", "rad2deg", "degrees", "radians", "isfinite", "isinf", "isnan", "signbit", "floor", "ceil",
          "trunc", "rint", "fabs", "sign", "absolute", "logical_not"]
_BINARY = ["add", "subtract", "multiply", "divide", "true_divide", "floor_divide", "power", "float_power", "remainder",
           "mod", "fmod", "logaddexp", "arctan2", "hypot", "greater", "greater_equal", "less", "less_equal", "not_equal",
           "equal", "logical_and", "logical_or", "logical_xor", "maximum", "minimum", "fmax", "fmin", "copysign", "nextafter"]


def test_arange_and_linspace(da):
    """creation/_arange.py, creation/_linspace.py: host-generated per block, staged once, then on the device path."""
    a = da.arange(2, 2000, 3, chunks=100)
    _close(a.compute(), np.arange(2, 2000, 3), "i8")
    _close((a * 2 + 1).sum().compute(), (np.arange(2, 2000, 3) * 2 + 1).sum(), "i8")
    _close(a[::-7].compute(), np.arange(2, 2000, 3)[::-7], "i8")
    f = da.linspace(1.4, 4.9, 1300, chunks=500)
    _close(f.compute(), np.linspace(1.4, 4.9, 1300), "f8")
    _close(f.mean().compute(), np.linspace(1.4, 4.9, 1300).mean(), "f8")
    assert da.arange(0).compute().shape == (0,)


def test_reshape_values(da):
    """manipulation/_reshape.py (tests/test_reshape.py value cases): merges, splits, kept axes, -1, one block,
    merge_chunks=False, and a reshape between device ops."""
    rng = np.random.default_rng(14)
    xh = rng.random((6, 5, 4))
    x = da.from_array(xh, chunks=(3, 2, 2))
    for shape in ((30, 4), (3, 2, 5, 4), (6, 20), (120,), (6, 5, 2, 2), (2, 3, 20), (-1, 4), (1, 30, 4, 1)):
        _close(x.reshape(shape).compute(), xh.reshape(shape), "f8")
    _close(da.from_array(xh, chunks=(6, 5, 4)).reshape((4, 5, 6)).compute(), xh.reshape((4, 5, 6)), "f8")
    _close(x.reshape((30, 4), merge_chunks=False).compute(), xh.reshape((30, 4)), "f8")
    y = (x.T * 2).reshape((20, 6))                       # non-contiguous blocks in front of the views
    _close(y.compute(), (xh.T * 2).reshape((20, 6)), "f8")
    _close(y.reshape((4, 5, 6)).sum(axis=1).compute(), (xh.T * 2).reshape((4, 5, 6)).sum(axis=1), "f8")
    ih = np.arange(64 * 48, dtype=np.int32).reshape(64, 48)
    i = da.from_array(ih, chunks=(16, 12))
    assert np.array_equal(i.reshape((8, 8, 48)).compute(), ih.reshape((8, 8, 48)))
    assert np.array_equal(i.reshape((64, 6, 8)).max(axis=2).compute(), ih.reshape((64, 6, 8)).max(axis=2))


def test_round_clip_and_axis_moves(da):
    rng = np.random.default_rng(13)
    xh = (rng.random((6, 10, 14)) * 200 - 100)
    x = da.from_array(xh, chunks=(4, 5, 6))
    for d in (0, 2, -1):
        _close(da.round(x, d).compute(), np.round(xh, d), "f8")
    _close(x.round(1).compute(), xh.round(1), "f8")
    _close(x.clip(-10, 25.5).compute(), xh.clip(-10, 25.5), "f8")
    _close(da.clip(x, None, 3).compute(), np.clip(xh, None, 3), "f8")
    _close(x.clip(min=-1).compute(), xh.clip(min=-1), "f8")
    _close(x.swapaxes(0, 2).compute(), xh.swapaxes(0, 2), "f8")
    _close(da.moveaxis(x, 0, -1).compute(), np.moveaxis(xh, 0, -1), "f8")
    _close(da.moveaxis(x, (0, 1), (2, 0)).compute(), np.moveaxis(xh, (0, 1), (2, 0)), "f8")
    _close(da.rollaxis(x, 2, 0).compute(), np.rollaxis(xh, 2, 0), "f8")
    _close(x.imag.compute(), xh.imag, "f8")
    assert x.real.name == x.name and x.conj().name == x.name


def test_from_array_of_a_device_chunk(da):
    """io/_from_array.py:148-152: an array that already lives on the GPU is blocked on the device."""
    from dask_array_b200 import DeviceChunk

    xh = np.random.default_rng(16).random((96, 80), dtype=np.float32)
    dev = DeviceChunk.from_numpy(xh)
    x = da.from_array(dev, chunks=(32, 40))
    assert type(x.expr).__name__ == "Resident" and x.chunks == ((32,) * 3, (40, 40))
    _close(x.compute(), xh, "f4")
    _close((x * 2).sum(axis=0).compute(), (xh * 2).sum(axis=0), "f4")
    _close((x.T + 1).max(axis=1).compute(), (xh.T + 1).max(axis=1), "f4")
    one = da.from_array(dev)
    assert one.numblocks == (1, 1) and one.expr.operand("store").blocks[(0, 0)].ptr == dev.ptr      # zero copy
    _close(da.asarray(dev).mean().compute(), xh.mean(dtype=np.float32), "f4")


def test_numpy_functions_on_arrays(da):
    """Array.__array_function__ (_collection.py:866-923): same-named functions stay on the device; an unknown one
    warns, computes its arguments and runs NumPy on the results, like the reference."""
    import warnings

    xh = np.random.default_rng(15).random((40, 60))
    x = da.from_array(xh, chunks=(16, 25))
    _close(np.sum(x, axis=0).compute(), xh.sum(axis=0), "f8")
    _close(np.clip(x, 0.2, 0.7).compute(), np.clip(xh, 0.2, 0.7), "f8")
    _close(np.concatenate([x, x], axis=1).max(axis=0).compute(), np.concatenate([xh, xh], axis=1).max(axis=0), "f8")
    _close(np.matmul(x, x.T).compute(), xh @ xh.T, "f8")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = np.linalg.norm(x)
    assert any(issubclass(i.category, FutureWarning) for i in w)
    _close(got, np.linalg.norm(xh), "f8")


def test_integer_vocabulary_values(da):
    rng = np.random.default_rng(12)
    xh = rng.integers(-50, 50, (37, 53)).astype(np.int32)
    yh = rng.integers(-7, 8, (37, 53)).astype(np.int32)
    x, y = da.from_array(xh, chunks=(20, 30)), da.from_array(yh, chunks=(20, 30))
    with np.errstate(all="ignore"):
        for op in ("bitwise_not", "invert", "negative", "positive", "square", "absolute", "sign", "logical_not"):
            _close(getattr(da, op)(x).compute(), getattr(np, op)(xh), "i4")
        for op in ("bitwise_and", "bitwise_or", "bitwise_xor", "add", "subtract", "multiply", "maximum", "minimum",
                   "greater", "equal", "true_divide"):
            _close(getattr(da, op)(x, y).compute(), getattr(np, op)(xh, yh), "f8")
        nz = np.where(yh == 0, 3, yh).astype(np.int32)                       # division by zero is tested apart
        ynz = da.from_array(nz, chunks=(20, 30))
        for op in ("floor_divide", "remainder", "mod", "fmod"):
            _close(getattr(da, op)(x, ynz).compute(), getattr(np, op)(xh, nz), "i4")
            _close(getattr(da, op)(x, y).compute(), getattr(np, op)(xh, yh), "i4")      # NumPy: x // 0 == x % 0 == 0
        sh = np.abs(yh)
        s = da.from_array(sh, chunks=(20, 30))
        for op in ("left_shift", "right_shift"):
            _close(getattr(da, op)(x, s).compute(), getattr(np, op)(xh, sh), "i4")
        _close(da.power(x, s).compute(), np.power(xh, sh), "i4")


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_float_vocabulary_values(da, dtype):
    """Appendix A: every float operator against NumPy on a ragged 2-D array (two blocks per axis), inputs inside
    each operator's domain (special values: the last test of this file)."""
    rng = np.random.default_rng(11)
    with np.errstate(all="ignore"):
        for op in _UNARY:
            lo, hi = _DOMAIN.get(op, (-3, 3))
            xh = (rng.random((37, 53)) * (hi - lo) + lo).astype(dtype)
            got = getattr(da, op)(da.from_array(xh, chunks=(20, 30))).compute()
            _close(got, getattr(np, op)(xh), dtype)
        for op in _BINARY:
            lo, hi = _DOMAIN.get(op, (-3, 3))
            xh = (rng.random((37, 53)) * (hi - lo) + lo).astype(dtype)
            yh = (rng.random((37, 53)) * (hi - lo) + lo).astype(dtype)
            if op in ("floor_divide", "remainder", "mod", "fmod", "divide", "true_divide"):
                yh[np.abs(yh) < 0.05] = 0.5               # keep quotients away from the rounding cliff of floor()
            x, y = da.from_array(xh, chunks=(20, 30)), da.from_array(yh, chunks=(20, 30))
            _close(getattr(da, op)(x, y).compute(), getattr(np, op)(xh, yh), dtype)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_float_vocabulary_special_values(da, dtype):
    """NaN, +-inf and +-0 through the operators whose domain is the whole line (kept last: the riskiest cases)."""
    sp = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1.5, -2.5], dtype=dtype)
    xh = np.tile(sp, (5, 3))
    x = da.from_array(xh, chunks=(3, 8))
    with np.errstate(all="ignore"):
        for op in _UNARY:
            if op in _DOMAIN:
                continue
            _close(getattr(da, op)(x).compute(), getattr(np, op)(xh), dtype)
        yh = np.ascontiguousarray(np.tile(sp[::-1], (5, 3)))
        y = da.from_array(yh, chunks=(3, 8))
        for op in ("maximum", "minimum", "fmax", "fmin", "add", "subtract", "multiply", "less", "greater_equal", "equal",
                   "not_equal", "copysign", "logical_and", "logical_or", "hypot", "arctan2"):
            _close(getattr(da, op)(x, y).compute(), getattr(np, op)(xh, yh), dtype)
