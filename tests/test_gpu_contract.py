"""N-d tensordot / einsum / matmul for every number type (linalg/_tensordot.py:45-136,
253-334, _einsum.py:181-271).  Cases mirror the reference's own tests (tests/test_routines.py:321-399: the
1-D / 2-D rows of test_matmul with seed 3732, test_tensordot, test_tensordot_2,
test_tensordot_double_contraction_*).  Contract: fp64 rtol 1e-12; integers bit-exact; fp32 on the tensor
cores |err| <= 1e-5 * (|A| . |B|) (the bf16 x 3 split; stated in DESIGN.md)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


@pytest.mark.parametrize("x_shape,y_shape,x_chunks,y_chunks", [
    ((7,), (7,), (), ()), ((7,), (7,), (2,), (3,)), ((7, 11), (11,), (), ()), ((7, 11), (11,), (3, 5), (2,)),
    ((11,), (11, 7), (), ()), ((11,), (11, 7), (4,), (3, 2)), ((7, 11), (11, 7), (), ()),
    ((7, 11), (11, 7), (3, 5), (4, 2)), ((7, 11), (11, 7), (7, 11), (11, 7))])
def test_matmul_fp64_reference_matrix(da, x_shape, y_shape, x_chunks, y_chunks):
    rng = np.random.default_rng(3732)
    x, y = rng.random(x_shape), rng.random(y_shape)
    a = da.from_array(x, chunks=x_chunks or tuple(max(i // 2, 1) for i in x.shape))
    b = da.from_array(y, chunks=y_chunks or tuple(max(i // 2, 1) for i in y.shape))
    want = np.matmul(x, y)
    got = da.matmul(a, b).compute()
    assert got.dtype == want.dtype and np.shape(got) == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-12)
    with pytest.raises(ValueError):
        da.matmul(da.from_array(np.float64(2.0)), b)


def test_tensordot_integers_bit_exact(da):
    x = np.arange(400).reshape((20, 20))
    y = np.arange(200).reshape((20, 10))
    a, b = da.from_array(x, chunks=(5, 4)), da.from_array(y, chunks=(4, 5))
    for axes in [1, (1, 0), (-1, 0)]:
        want = np.tensordot(x, y, axes=axes)
        for p, q in ((a, b), (x, b), (a, y)):
            got = da.tensordot(p, q, axes=axes).compute()
            assert got.dtype == want.dtype and np.array_equal(got, want)
    assert da.tensordot(a, b, axes=(1, 0)).name == da.tensordot(a, b, axes=(1, 0)).name


@pytest.mark.parametrize("axes", [1, (0, 1), (1, 0), ((1, 0), (2, 1)), ((1, 2), (2, 0)), ((2, 0), (1, 2))])
def test_tensordot_3d(da, axes):
    x = np.arange(4 * 4 * 4).reshape((4, 4, 4))
    y = da.from_array(x, chunks=2)
    assert np.array_equal(da.tensordot(y, y, axes=axes).compute(), np.tensordot(x, x, axes=axes))


@pytest.mark.parametrize("chunks", [(4, 6), (2, 3), (4, 3), (2, 6)])
def test_tensordot_double_contraction(da, chunks):
    x = np.arange(24).reshape(4, 6)
    assert np.array_equal(da.tensordot(da.from_array(x, chunks=chunks), da.from_array(x, chunks=chunks), axes=2).compute(),
                          np.tensordot(x, x, axes=2))
    u, v = np.arange(60.0).reshape(3, 4, 5), np.arange(60.0).reshape(4, 5, 3)
    np.testing.assert_allclose(da.tensordot(da.from_array(u, chunks=3), da.from_array(v), axes=2).compute(),
                               np.tensordot(u, v, axes=2), rtol=1e-12)


def test_tensordot_fp32_nd_tensor_cores_and_mixed_dtypes(da):
    rng = np.random.default_rng(7)
    a = (rng.random((24, 16, 32)) - 0.5).astype(np.float32)
    b = (rng.random((32, 16, 40)) - 0.5).astype(np.float32)
    got = da.tensordot(da.from_array(a, chunks=(12, 8, 16)), da.from_array(b, chunks=(16, 8, 20)), axes=((2, 1), (0, 1))).compute()
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    want = np.tensordot(a64, b64, axes=((2, 1), (0, 1)))
    bound = np.tensordot(np.abs(a64), np.abs(b64), axes=((2, 1), (0, 1)))
    assert got.dtype == np.float32 and np.all(np.abs(got - want) <= 1e-5 * bound)
    # fp32 with a contraction length that is not a multiple of 8: IEEE fp32 on the exact kernel
    c, d = a[:, :, :7], b[:7]
    got = da.tensordot(da.from_array(c, chunks=(12, 16, 7)), da.from_array(d, chunks=(7, 16, 40)), axes=((2,), (0,))).compute()
    np.testing.assert_allclose(got, np.tensordot(c, d, axes=((2,), (0,))), rtol=2e-5, atol=1e-5)
    # mixed int32 x float64 promotes like NumPy
    i = rng.integers(-5, 5, (6, 9)).astype(np.int32)
    f = rng.random((9, 4))
    got = da.tensordot(da.from_array(i, chunks=(3, 3)), da.from_array(f, chunks=(3, 2)), axes=1).compute()
    assert got.dtype == np.float64
    np.testing.assert_allclose(got, i @ f, rtol=1e-12)


@pytest.mark.parametrize("subs,shapes", [
    ("ij,jk->ik", [(6, 8), (8, 5)]), ("ij,jk", [(6, 8), (8, 5)]), ("ijk,kjl->il", [(4, 6, 8), (8, 6, 3)]),
    ("ij->ji", [(5, 7)]), ("ij->", [(5, 7)]), ("ij->j", [(5, 7)]), ("ij,ij->", [(5, 7), (5, 7)]),
    ("i,j->ij", [(5,), (7,)]), ("ij,jk,kl->li", [(6, 5), (5, 4), (4, 3)]), ("abc,cd->dab", [(3, 4, 5), (5, 6)])])
def test_einsum(da, subs, shapes):
    rng = np.random.default_rng(11)
    ops = [rng.random(s) for s in shapes]
    want = np.einsum(subs, *ops)
    got = da.einsum(subs, *[da.from_array(o, chunks=tuple(max(n // 2, 1) for n in o.shape)) for o in ops]).compute()
    np.testing.assert_allclose(got, want, rtol=1e-12)
    iops = [rng.integers(-9, 9, s) for s in shapes]
    goti = da.einsum(subs, *[da.from_array(o, chunks=tuple(max(n // 2, 1) for n in o.shape)) for o in iops]).compute()
    assert np.array_equal(goti, np.einsum(subs, *iops))


def test_einsum_refusals(da):
    x = da.ones((4, 4, 4), chunks=2)
    with pytest.raises(NotImplementedError, match="batch"):
        da.einsum("bij,bjk->bik", x, x)
    with pytest.raises(NotImplementedError, match="diagonal"):
        da.einsum("ii->i", da.ones((4, 4), chunks=2))


def test_matmul_fp64_large_blocks_against_the_oracle(da):
    """fp64 blocked matmul with real block sizes against the oracle's per-block np.matmul + k-sum order."""
    from oracle import reference as ref
    rng = np.random.default_rng(3732)
    ah, bh = rng.random((300, 520)) - 0.5, rng.random((520, 260)) - 0.5
    got = (da.from_array(ah, chunks=(128, 200)) @ da.from_array(bh, chunks=(200, 128))).compute()
    want = ref.matmul(ref.Blocked.from_array(ah, (128, 200)), ref.Blocked.from_array(bh, (200, 128))).to_array()
    bound = np.abs(ah) @ np.abs(bh)
    assert np.all(np.abs(got - want) <= 1e-12 * bound)
