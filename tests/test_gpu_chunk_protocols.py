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


def test_reference_cumulative_layer_runs_on_device_chunks():
    """The reference's sequential CumReduction plan (`_cumulative.py:174-264`): np.cumsum per block,
    `extra = extra + tail`, `result = extra + block` -- executed here with DeviceChunk blocks."""
    from dask_array_b200 import DeviceChunk
    rng = np.random.default_rng(2)
    xh = rng.integers(-9, 9, size=(90, 40)).astype(np.int32)
    blocks = [DeviceChunk.from_numpy(np.ascontiguousarray(xh[i:i + 30])) for i in range(0, 90, 30)]
    scanned = [np.cumsum(b, axis=0) for b in blocks]                     # NEP-18 -> one scan launch each
    assert all(isinstance(s, DeviceChunk) and s.dtype == np.int64 for s in scanned)
    out, extra = [scanned[0]], None
    for prev, cur in zip(scanned, scanned[1:]):
        tail = prev[-1:]                                                  # _cum_tail: last hyperplane
        extra = tail if extra is None else extra + tail
        out.append(extra + cur)
    got = np.concatenate([o.to_numpy() for o in out], axis=0)
    assert np.array_equal(got, np.cumsum(xh, axis=0))
    f = rng.random((7, 33))
    f[2, 5] = np.nan
    fd = DeviceChunk.from_numpy(f)
    np.testing.assert_allclose(np.nancumsum(fd, axis=1).to_numpy(), np.nancumsum(f, axis=1), rtol=1e-12)
    np.testing.assert_allclose(np.cumprod(fd[:2], axis=1).to_numpy(), np.cumprod(f[:2], axis=1), rtol=1e-12)
