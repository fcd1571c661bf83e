"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle only
finishes in seconds at small sizes):

* C2 (32768 x 32768 fp32, 4096^2 chunks): the array is one seeded 4096^2 block tiled 8 x 8, so
  the exact answer follows from fp64 NumPy on ONE block: column means repeat per block column,
  std() equals the block's std; plus chunk-structure independence (2048^2 chunks, same data).
* C3 (65536 x 16384 fp64, (8192,16384) chunks): one adversarial 8192-row block (ties, NaN, +-inf,
  -0.0) tiled 8 x: argmax/argmin/max/min(axis=1) must equal NumPy's on the block, bit for bit.
* C4 (16384^2, distinct int32 values): rechunk there-and-back is the identity (checked on the
  device), x.T + x is symmetric and its checksum is 2 * sum(x) exactly.
* C5 (32768^2 bf16, 4096^2 chunks): linearity, sum(x @ y.T) == colsum(x) . colsum(y).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def test_c2_fused_chain_full_size(da):
    base = np.random.default_rng(0).random((4096, 4096), dtype=np.float32)
    x = da.from_host_blocks(lambda bid: base, (32768, 32768), (4096, 4096), np.float32, token="full-c2").persist()
    y = da.sin(x) * 2 + x**2
    mean0, std = da.compute(y.mean(axis=0), y.std())
    b64 = base.astype(np.float64)
    y64 = np.sin(b64) * 2 + b64**2
    assert mean0.shape == (32768,) and mean0.dtype == np.float32
    np.testing.assert_allclose(mean0, np.tile(y64.mean(axis=0), 8), rtol=1e-5)
    np.testing.assert_allclose(std, y64.std(), rtol=1e-5)
    # linearity / consistency: mean of column means == global mean; std^2 == E[y^2] - E[y]^2
    gm = y.mean().compute()
    np.testing.assert_allclose(mean0.astype(np.float64).mean(), gm, rtol=1e-5)
    np.testing.assert_allclose((y * y).mean().compute() - float(gm) ** 2, float(std) ** 2, rtol=1e-4)
    # chunk-structure independence (tests/test_reductions.py:1060-1079) at full size
    x2 = x.rechunk((2048, 2048)).persist()
    y2 = da.sin(x2) * 2 + x2**2
    np.testing.assert_allclose(y2.mean(axis=0).compute(), mean0, rtol=1e-5)
    np.testing.assert_allclose(y2.std().compute(), std, rtol=1e-5)


def test_c3_arg_minmax_full_size_bit_exact(da):
    rng = np.random.default_rng(0)
    base = rng.random((8192, 16384))
    base[::7, 100] = base[::7, 9000] = 2.0            # duplicated row maxima: first occurrence must win
    base[::11, 50] = base[::11, 12000] = -1.0         # duplicated row minima
    base[5, 77] = np.nan; base[5, 3] = np.nan         # NaN wins, first NaN
    base[9, 1] = np.inf; base[10, 2] = -np.inf
    base[12, :] = 0.0; base[12, 5] = -0.0             # all-equal row with a signed zero
    x = da.from_host_blocks(lambda bid: base, (65536, 16384), (8192, 16384), np.float64, token="full-c3").persist()
    amax, amin, vmax, vmin = da.compute(x.argmax(axis=1), x.argmin(axis=1), x.max(axis=1), x.min(axis=1))
    assert amax.dtype == np.int64 and amax.shape == (65536,)
    assert np.array_equal(amax, np.tile(np.argmax(base, axis=1), 8))
    assert np.array_equal(amin, np.tile(np.argmin(base, axis=1), 8))
    wmax, wmin = np.max(base, axis=1), np.min(base, axis=1)
    ok = np.arange(8192) != 12                         # +-0 ties: sign of zero is declared out of contract
    assert np.array_equal(vmax.reshape(8, -1)[:, ok], np.tile(wmax[ok], (8, 1)), equal_nan=True)
    assert np.array_equal(vmin.reshape(8, -1)[:, ok], np.tile(wmin[ok], (8, 1)), equal_nan=True)
    assert np.all(vmax.reshape(8, -1)[:, 12] == 0.0)


def test_c4_rechunk_round_trip_and_symmetry_full_size(da):
    n = 16384
    xh = np.arange(n * n, dtype=np.int32).reshape(n, n)        # distinct values
    x = da.from_array(xh, chunks=(n, 256)).persist()
    r = x.rechunk((256, n)).persist()
    assert r.chunks == ((256,) * 64, (n,))
    back = r.rechunk((n, 256))
    assert not (back != x).any().compute()                    # round trip == identity, checked on the device
    assert r.sum(dtype="int64").compute() == int(xh.sum(dtype=np.int64))   # checksum
    blk = r.expr.operand("store").blocks[(3, 0)].to_numpy()    # one re-blocked panel against the source
    assert np.array_equal(blk, xh[768:1024, :])
    sq = da.from_array(xh, chunks=(2048, 2048)).persist()
    y = (sq.T + sq).persist()
    assert not (y != y.T).any().compute()                     # symmetric
    assert y.sum(dtype="int64").compute() == 2 * int(xh.sum(dtype=np.int64))
    assert np.array_equal(y.expr.operand("store").blocks[(1, 5)].to_numpy(),
                          xh[2048:4096, 10240:12288] + xh[10240:12288, 2048:4096].T)


def test_c5_matmul_linearity_full_size(da):
    import ml_dtypes
    n, cb = 32768, 4096
    rng = np.random.default_rng(0)
    bx = (rng.random((cb, cb), dtype=np.float32) - 0.5).astype(ml_dtypes.bfloat16)
    by = (rng.random((cb, cb), dtype=np.float32) - 0.5).astype(ml_dtypes.bfloat16)
    x = da.from_host_blocks(lambda bid: bx, (n, n), (cb, cb), ml_dtypes.bfloat16, token="full-c5x").persist()
    y = da.from_host_blocks(lambda bid: by, (n, n), (cb, cb), ml_dtypes.bfloat16, token="full-c5y").persist()
    z = (x @ y.T).persist()
    total = z.sum().compute()
    cx = np.tile(bx.astype(np.float64).sum(axis=0) * 8, 8)     # column sums of the tiled operands
    cy = np.tile(by.astype(np.float64).sum(axis=0) * 8, 8)
    want = float(cx @ cy)
    scale = float(np.abs(bx.astype(np.float64)).sum() * 64) * float(np.abs(by.astype(np.float64)).mean())
    assert abs(float(total) - want) <= 1e-5 * scale
    # one output block against fp64 NumPy on the bf16 values: every block equals 8 * bx @ by.T
    ref = 8.0 * (bx[:256].astype(np.float64) @ by[:256].astype(np.float64).T)
    got = z.expr.operand("store").blocks[(2, 5)][:256, :256]
    bound = 8.0 * (np.abs(bx[:256].astype(np.float64)) @ np.abs(by[:256].astype(np.float64)).T)
    from dask_array_b200 import _eager
    assert np.all(np.abs(_eager.copy(got).to_numpy() - ref) <= 2e-5 * bound)
