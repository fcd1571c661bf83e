"""The chunk-type boundary: NumPy-shaped chunk functions -- here the oracle's
restatements of the reference's mean_chunk / moment_chunk / moment_agg / arg_chunk structure and
plain np.* calls -- run UNMODIFIED on DeviceChunk through NEP-13 / NEP-18."""
import numpy as np
import pytest

from oracle import reference as ref

pytestmark = pytest.mark.gpu


def test_ufuncs_operators_and_mixed_host_operands():
    from dask_array_b200 import DeviceChunk
    rng = np.random.default_rng(0)
    xh = rng.random((64, 48), dtype=np.float32)
    x = DeviceChunk.from_numpy(xh)
    y = np.sin(x) * 2 + x**2                      # NEP-13 via np.sin and the operator dunders
    assert isinstance(y, DeviceChunk)
    np.testing.assert_allclose(y.to_numpy(), np.sin(xh) * 2 + xh**2, rtol=3e-7)
    n = np.full((1, 48), 64, dtype=np.int64)      # host operand, as numel() produces
    got = np.sum(x, axis=0, keepdims=True) / n
    assert got.dtype == (np.sum(xh, axis=0, keepdims=True) / n).dtype
    np.testing.assert_allclose(got.to_numpy(), np.sum(xh, axis=0, keepdims=True) / n, rtol=1e-5)
    assert np.array_equal((x > 0.5).to_numpy(), xh > 0.5)
    assert np.array_equal(np.transpose(x).to_numpy(), xh.T)
    assert np.array_equal(np.where(x > 0.5, x, 0).to_numpy(), np.where(xh > 0.5, xh, 0))
    assert np.array_equal(np.concatenate([x, x.T.T], axis=1).to_numpy(), np.concatenate([xh, xh], axis=1))
    assert np.array_equal(x.astype("float64").to_numpy(), xh.astype("float64"))
    with pytest.raises(TypeError):
        np.empty_like(x, dtype=[("vals", "f4"), ("arg", "i8")])       # selects arg_chunk's dict path
    with pytest.raises(TypeError):
        np.linalg.svd(x)                                               # no silent host fallback


def test_reference_shaped_chunk_functions_run_on_device_chunks():
    from dask_array_b200 import DeviceChunk
    rng = np.random.default_rng(1)
    xh = (rng.random((96, 80)) * 10 + 100).astype(np.float32)
    x = DeviceChunk.from_numpy(xh)
    for axis in [(0,), (1,), (0, 1)]:
        got = ref.mean_chunk(x, np.dtype("f4"), axis, True)
        want = ref.mean_chunk(xh, np.dtype("f4"), axis, True)
        np.testing.assert_allclose(got["total"].to_numpy(), want["total"], rtol=1e-6)
        assert np.array_equal(got["n"], want["n"])
        got = ref.moment_chunk(x, np.dtype("f4"), axis, True)         # A - u, d**2, np.stack ... on device
        want = ref.moment_chunk(xh, np.dtype("f4"), axis, True)
        np.testing.assert_allclose(got["total"].to_numpy(), want["total"], rtol=1e-6)
        np.testing.assert_allclose(got["M"].to_numpy(), want["M"], rtol=2e-5)
    assert np.array_equal(np.argmax(x, axis=1).to_numpy(), np.argmax(xh, axis=1))
    assert np.array_equal(ref.chunk_max(x, axis=(0,), keepdims=True).to_numpy(), xh.max(axis=0, keepdims=True))
