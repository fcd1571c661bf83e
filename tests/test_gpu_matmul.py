"""Blocked matmul on tcgen05 (b2_gemm_tn_pairs) against the oracle's per-block np.matmul +
sequential k-sum (linalg/_tensordot.py:194-249).  Tolerances: bf16 operands -> products are
exact, only fp32 accumulation order differs (rtol 2e-5 of |A||B|); fp32 operands via bf16x3
split -> rtol 1e-5 (north_star) relative to the row/column magnitude bound."""
import numpy as np
import pytest

from oracle import reference as ref

pytestmark = pytest.mark.gpu


def _bf16():
    import ml_dtypes
    return ml_dtypes.bfloat16


def _check(got, want, a64, b64, rtol):
    bound = np.abs(a64) @ np.abs(b64)           # magnitude bound of every dot product
    err = np.abs(got.astype(np.float64) - want)
    assert np.all(err <= rtol * bound + 1e-30), float((err / bound).max())


@pytest.mark.parametrize("M,N,K,cm,cn,ck", [(256, 256, 256, 128, 128, 128), (384, 256, 512, 128, 128, 256),
                                            (200, 136, 264, 100, 72, 88), (1024, 1024, 1024, 512, 512, 256)])
def test_matmul_bf16(M, N, K, cm, cn, ck):
    import dask_array_b200 as da
    rng = np.random.default_rng(3732)
    ah = (rng.random((M, K)) - 0.5).astype(_bf16())
    yh = (rng.random((N, K)) - 0.5).astype(_bf16())
    a = da.from_array(ah, chunks=(cm, ck))
    y = da.from_array(yh, chunks=(cn, ck))
    got = (a @ y.T).compute()
    assert got.dtype == np.float32 and got.shape == (M, N)
    a64, b64 = ah.astype(np.float64), yh.astype(np.float64).T
    _check(got, a64 @ b64, a64, b64, 2e-5)
    # the reference's own block/k order on the bf16-rounded inputs in fp32
    want = ref.matmul(ref.Blocked.from_array(ah.astype(np.float32), (cm, ck)),
                      ref.Blocked.from_array(yh.astype(np.float32).T, (ck, cn))).to_array()
    _check(got, want.astype(np.float64), a64, b64, 4e-5)


@pytest.mark.parametrize("M,N,K,c", [(256, 256, 256, 128), (320, 192, 448, 64)])
def test_matmul_fp32_split(M, N, K, c):
    import dask_array_b200 as da
    rng = np.random.default_rng(3732)
    ah = (rng.random((M, K)) - 0.5).astype(np.float32)
    bh = (rng.random((K, N)) - 0.5).astype(np.float32)
    a = da.from_array(ah, chunks=(c, c))
    b = da.from_array(bh, chunks=(c, c))
    got = (a @ b).compute()                      # plain a @ b: b.T is materialised by the transpose kernel
    a64, b64 = ah.astype(np.float64), bh.astype(np.float64)
    _check(got, a64 @ b64, a64, b64, 1e-5)
    want = ref.matmul(ref.Blocked.from_array(ah, (c, c)), ref.Blocked.from_array(bh, (c, c))).to_array()
    _check(got, want.astype(np.float64), a64, b64, 1e-5)
    s = (a @ b).sum().compute()
    np.testing.assert_allclose(s, (a64 @ b64).sum(), rtol=1e-4, atol=1e-2)
