"""Random pipelines mirrored in NumPy -- the approach of the reference's
tests/test_fuzz_optimize.py:62-201: arange-valued arrays (all values distinct, so any block
index-mapping bug changes the result) pushed through random chains of
neg / add / mul / transpose / getitem / rechunk / binary ops and closed by a reduction.
Integer pipelines must be bit-exact; everything runs through the optimiser (pushdowns, fusion,
conflict rule) and the GPU kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pipeline(rng, da, dtype):
    shape = (int(rng.integers(20, 90)), int(rng.integers(20, 90)))
    chunks = (int(rng.integers(5, 40)), int(rng.integers(5, 40)))
    xh = np.arange(shape[0] * shape[1], dtype=dtype).reshape(shape) - 100
    x, ref = da.from_array(xh, chunks=chunks), xh
    y_other = None
    steps = []
    for _ in range(int(rng.integers(2, 7))):
        op = rng.choice(["neg", "add_s", "mul_s", "T", "slice", "rechunk", "add_T", "add_self", "where", "abs"])
        steps.append(op)
        if op == "neg":
            x, ref = -x, -ref
        elif op == "add_s":
            s = int(rng.integers(-5, 6))
            x, ref = x + s, ref + s
        elif op == "mul_s":
            s = int(rng.integers(-3, 4))
            x, ref = x * s, ref * s
        elif op == "T":
            x, ref = x.T, ref.T
        elif op == "slice":
            r0 = int(rng.integers(0, ref.shape[0] // 2 + 1)); r1 = int(rng.integers(r0 + 1, ref.shape[0] + 1))
            c0 = int(rng.integers(0, ref.shape[1] // 2 + 1)); c1 = int(rng.integers(c0 + 1, ref.shape[1] + 1))
            x, ref = x[r0:r1, c0:c1], ref[r0:r1, c0:c1]
        elif op == "rechunk":
            nc = (int(rng.integers(3, max(4, ref.shape[0]))), int(rng.integers(3, max(4, ref.shape[1]))))
            x = x.rechunk(nc)
        elif op == "add_T" and ref.shape[0] == ref.shape[1]:
            x, ref = x + x.T, ref + ref.T
        elif op == "add_self":
            x, ref = x + x * 2, ref + ref * 2
        elif op == "where":
            x, ref = da.where(x > 0, x, -x), np.where(ref > 0, ref, -ref)
        elif op == "abs":
            x, ref = abs(x), abs(ref)
    return x, ref, steps


@pytest.mark.parametrize("seed", range(40))
def test_random_integer_pipelines_bit_exact(seed):
    import dask_array_b200 as da
    rng = np.random.default_rng(seed)
    x, ref, steps = _pipeline(rng, da, np.int64)
    assert x.shape == ref.shape, steps
    got = x.compute()
    assert got.dtype == ref.dtype and np.array_equal(got, ref), steps
    fin = rng.choice(["sum", "sum0", "max1", "argmax", "argmin0", "mean", "min"])
    if fin == "sum":
        assert x.sum().compute() == ref.sum(), steps
    elif fin == "sum0":
        assert np.array_equal(x.sum(axis=0).compute(), ref.sum(axis=0)), steps
    elif fin == "max1":
        assert np.array_equal(x.max(axis=1).compute(), ref.max(axis=1)), steps
    elif fin == "argmax":
        assert np.array_equal(x.argmax(axis=1).compute(), ref.argmax(axis=1)), steps
    elif fin == "argmin0":
        assert np.array_equal(x.argmin(axis=0).compute(), ref.argmin(axis=0)), steps
    elif fin == "mean":
        np.testing.assert_allclose(x.mean(axis=1).compute(), ref.mean(axis=1), rtol=1e-12, err_msg=str(steps))
    else:
        assert x.min().compute() == ref.min(), steps


@pytest.mark.parametrize("seed", range(12))
def test_random_float_pipelines(seed):
    import dask_array_b200 as da
    rng = np.random.default_rng(1000 + seed)
    x, ref, steps = _pipeline(rng, da, np.float64)
    y, yref = da.sqrt(abs(x) + 1) * da.cos(x), np.sqrt(abs(ref) + 1) * np.cos(ref)
    np.testing.assert_allclose(y.compute(), yref, rtol=1e-13, atol=1e-13, err_msg=str(steps))
    np.testing.assert_allclose(y.std(axis=0).compute(), yref.std(axis=0), rtol=1e-10, err_msg=str(steps))
    np.testing.assert_allclose(y.var(ddof=1).compute(), yref.var(ddof=1), rtol=1e-10, err_msg=str(steps))
