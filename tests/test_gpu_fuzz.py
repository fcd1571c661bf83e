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


def _pipeline2(rng, da):
    """The round's new operators mixed with the old ones: stepped / reversed slices, np.newaxis,
    operands on different chunk grids (unification), cumulative scans, halo exchange, concatenate."""
    shape = (int(rng.integers(24, 80)), int(rng.integers(24, 80)))
    chunks = (int(rng.integers(6, 30)), int(rng.integers(6, 30)))
    xh = (np.arange(shape[0] * shape[1], dtype=np.int64).reshape(shape) * 7919) % 1009 - 500
    x, ref = da.from_array(xh, chunks=chunks), xh
    if rng.random() < 0.5:
        x = x.persist()
    steps = []
    for _ in range(int(rng.integers(2, 6))):
        op = rng.choice(["step_slice", "reverse", "regrid_add", "cumsum", "overlap_trim", "roll", "T", "mul_s",
                         "rechunk", "newaxis_mul", "stencil"])
        steps.append(op)
        if op == "step_slice":
            s0, s1 = int(rng.choice([-3, -2, -1, 1, 2, 3])), int(rng.choice([-2, -1, 1, 2, 4]))
            if ref.shape[0] > 8 and ref.shape[1] > 8:
                x, ref = x[::s0, 1::s1], ref[::s0, 1::s1]
        elif op == "reverse":
            x, ref = x[::-1], ref[::-1]
        elif op == "regrid_add":
            other = da.from_array(np.ascontiguousarray(ref) * 2, chunks=(int(rng.integers(3, 20)), int(rng.integers(3, 20))))
            x, ref = x + other, ref + ref * 2
        elif op == "cumsum":
            ax = int(rng.integers(0, 2))
            x, ref = x.cumsum(axis=ax), np.cumsum(ref, axis=ax)
        elif op == "overlap_trim":
            d = int(rng.integers(1, 3))
            bnd = str(rng.choice(["none", "periodic", "reflect", "nearest"]))
            if min(ref.shape) > 2 * d + 2:
                x = da.overlap.trim_internal(da.overlap.overlap(x, depth=d, boundary=bnd), {0: d, 1: d}, bnd)
        elif op == "roll":
            k = int(rng.integers(1, max(2, ref.shape[0] - 1)))
            x, ref = da.concatenate([x[k:], x[:k]]), np.concatenate([ref[k:], ref[:k]])
        elif op == "T":
            x, ref = x.T, ref.T
        elif op == "mul_s":
            s = int(rng.integers(-3, 4))
            x, ref = x * s, ref * s
        elif op == "rechunk":
            x = x.rechunk((int(rng.integers(3, max(4, ref.shape[0]))), int(rng.integers(3, max(4, ref.shape[1])))))
        elif op == "newaxis_mul":
            x, ref = x * x[0][None, :] - x[:, 0][:, None], ref * ref[0][None, :] - ref[:, 0][:, None]
        elif op == "stencil" and min(ref.shape) > 6:
            x = da.overlap.overlap(x, depth={0: 1, 1: 0}, boundary={0: "periodic", 1: "none"}).map_blocks(
                lambda b: b[2:] + b[:-2] - 2 * b[1:-1], chunks=x.chunks)
            ref = np.roll(ref, -1, 0) + np.roll(ref, 1, 0) - 2 * ref
    return x, ref, steps


@pytest.mark.parametrize("seed", range(40))
def test_random_pipelines_with_new_operators_bit_exact(seed):
    import dask_array_b200 as da
    rng = np.random.default_rng(7000 + seed)
    x, ref, steps = _pipeline2(rng, da)
    assert x.shape == ref.shape, steps
    got = x.compute()
    assert got.dtype == ref.dtype and np.array_equal(got, ref), steps
    assert np.array_equal(x.sum(axis=0).compute(), ref.sum(axis=0)), steps
    k = min(3, ref.shape[1])
    assert np.array_equal(x.topk(k, axis=1).compute(), np.sort(ref, axis=1)[:, ::-1][:, :k]), steps
