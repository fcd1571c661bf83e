"""``f(x, x.T)`` chains (the reference's ``a + a.T`` fusion case, tests/test_collection.py:996-1135,
README example): bit-exact against NumPy, and the launch really is the mirror-pair kernel
(`b2_run_ewt_sym`: every input tile read once) where the block grid allows it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def _variants(step):
    return [k.spec.variant for k in step.fused_launches()]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("n,chunk", [(512, 128), (1000, 300), (768, 768), (224, 64), (96, 32)])
def test_sub_transpose_bit_exact(da, dtype, n, chunk):
    rng = np.random.default_rng(n + chunk)
    xh = (rng.standard_normal((n, n)) * 1000).astype(dtype)
    x = da.from_array(xh, chunks=(chunk, chunk)).persist()
    for expr, want in ((x - x.T, xh - xh.T), (x.T - x, xh.T - xh), (x.T * 3 + x, xh.T * 3 + xh)):
        step = da.compile(expr)
        assert _variants(step) == ["sym"], _variants(step)
        step.run()
        got = step.results()[0]
        assert got.dtype == want.dtype
        assert np.array_equal(got, want)


def test_comparison_output_and_ragged_edges(da):
    rng = np.random.default_rng(5)
    xh = rng.integers(-5, 5, size=(1100, 1100)).astype(np.float32)
    x = da.from_array(xh, chunks=(256, 256)).persist()         # edge blocks 76 wide (76 % 4 == 0)
    step = da.compile(x > x.T)
    assert _variants(step) == ["sym"]
    step.run()
    assert np.array_equal(step.results()[0], xh > xh.T)
    # extents that are not a multiple of the vector width: the generic staged kernel takes over
    yh = xh[:1001, :1001].copy()
    y = da.from_array(yh, chunks=(250, 250)).persist()
    step = da.compile(y + y.T)
    assert "sym" not in _variants(step)
    step.run()
    assert np.array_equal(step.results()[0], yh + yh.T)


def test_not_a_mirror_falls_back(da):
    rng = np.random.default_rng(6)
    ah, bh = rng.random((512, 512)), rng.random((512, 512))
    a, b = da.from_array(ah, chunks=(128, 128)).persist(), da.from_array(bh, chunks=(128, 128)).persist()
    step = da.compile(a + b.T)                                   # two different arrays: no pairing
    assert "sym" not in _variants(step)
    step.run()
    assert np.array_equal(step.results()[0], ah + bh.T)
    # non-square: out block (i, j) has no transposed partner of the same launch shape pattern
    ch = rng.random((512, 256))
    c = da.from_array(ch, chunks=(128, 128)).persist()
    got = (c.T[:, :256] + c[:256, :]).compute()
    assert np.array_equal(got, ch.T[:, :256] + ch[:256, :])


def test_symmetric_result_and_checksum_large(da):
    n = 4096
    xh = np.arange(n * n, dtype=np.int32).reshape(n, n).view(np.float32).copy()
    xh[~np.isfinite(xh)] = 1.0
    x = da.from_array(xh, chunks=(1024, 1024)).persist()
    step = da.compile(x.T + x)
    assert _variants(step) == ["sym"]
    step.run()
    got = step.results()[0]
    assert np.array_equal(got, xh.T + xh)
