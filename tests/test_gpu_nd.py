"""N-d blocks through the canonical (B, R, C) view: 1-D, 3-D and 4-D arrays, broadcasting across
ranks, reductions over leading / middle / trailing / interleaved axis sets, arg reductions along
every axis, transposes with arbitrary permutations -- against NumPy (ints bit-exact)."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def test_1d(da):
    xh = np.arange(1000, dtype=np.int64) * 3 - 500
    x = da.from_array(xh, chunks=137)
    assert np.array_equal((x * 2 + 1).compute(), xh * 2 + 1)
    assert x.sum().compute() == xh.sum() and x.argmax().compute() == xh.argmax()
    assert np.array_equal(x[100:900].rechunk(250).compute(), xh[100:900])
    np.testing.assert_allclose(x.std().compute(), xh.std(), rtol=1e-12)


@pytest.mark.parametrize("shape,chunks", [((12, 10, 14), (5, 4, 6)), ((6, 5, 4, 7), (4, 3, 2, 5))])
def test_nd_reductions_all_axis_sets(da, shape, chunks):
    rng = np.random.default_rng(0)
    xh = rng.integers(-1000, 1000, shape).astype(np.int64)
    fh = xh.astype(np.float64) / 7
    x, f = da.from_array(xh, chunks=chunks), da.from_array(fh, chunks=chunks)
    nd = len(shape)
    for r in range(1, nd + 1):
        for axes in itertools.combinations(range(nd), r):
            assert np.array_equal(x.sum(axis=axes).compute(), xh.sum(axis=axes)), axes
            assert np.array_equal(x.max(axis=axes, keepdims=True).compute(), xh.max(axis=axes, keepdims=True)), axes
            np.testing.assert_allclose(f.mean(axis=axes).compute(), fh.mean(axis=axes), rtol=1e-12, err_msg=str(axes))
            np.testing.assert_allclose(f.var(axis=axes).compute(), fh.var(axis=axes), rtol=1e-12, err_msg=str(axes))
    for ax in range(nd):
        assert np.array_equal(x.argmax(axis=ax).compute(), xh.argmax(axis=ax)), ax
        assert np.array_equal(x.argmin(axis=ax).compute(), xh.argmin(axis=ax)), ax


def test_nd_elementwise_broadcast_and_transpose(da):
    rng = np.random.default_rng(1)
    ah = rng.integers(-50, 50, (6, 8, 10)).astype(np.int32)
    bh = rng.integers(-50, 50, (8, 10)).astype(np.int32)
    ch = rng.integers(-50, 50, (6, 1, 10)).astype(np.int32)
    a, b, c = da.from_array(ah, chunks=(3, 4, 5)), da.from_array(bh, chunks=(4, 5)), da.from_array(ch, chunks=(3, 1, 5))
    assert np.array_equal((a * b + c).compute(), ah * bh + ch)
    for perm in itertools.permutations(range(3)):
        assert np.array_equal(a.transpose(perm).compute(), ah.transpose(perm)), perm
        assert np.array_equal((a.transpose(perm) * 2).sum(axis=0).compute(), (ah.transpose(perm) * 2).sum(axis=0)), perm
    four = rng.integers(0, 9, (3, 4, 5, 6)).astype(np.int64)
    d = da.from_array(four, chunks=(2, 2, 3, 4))
    assert np.array_equal((d.transpose(3, 1, 0, 2) + 1).compute(), four.transpose(3, 1, 0, 2) + 1)
    assert np.array_equal(d[1:3, :, 2:5, 1:].compute(), four[1:3, :, 2:5, 1:])
    assert np.array_equal(d.rechunk((3, 1, 5, 2)).compute(), four)


def test_stepped_and_reversed_slices_bit_exact():
    """SliceSlicesIntegers with steps of either sign (slicing/_basic.py:357-493, `_slice_1d`
    slicing/_utils.py:279-440): zero-copy strided views per block, blocks reversed for negative steps."""
    import dask_array_b200 as da
    rng = np.random.default_rng(9)
    xh = rng.integers(-100, 100, size=(97, 61)).astype(np.int32)
    x = da.from_array(xh, chunks=(20, 16)).persist()
    cases = [np.s_[::2], np.s_[::-1], np.s_[5:90:7, ::3], np.s_[::-3, 2], np.s_[90:3:-4, 60:1:-5],
             np.s_[3, ::-1], np.s_[::40], np.s_[-1:-98:-1, 1::2], np.s_[10:11:5], np.s_[50:10:3]]
    for idx in cases:
        y = x[idx]
        want = xh[idx]
        assert y.shape == want.shape, idx
        assert np.array_equal(y.compute(), want), idx
        if want.size:
            assert np.array_equal((y * 2 + 1).compute(), want * 2 + 1), idx          # fused chain on strided views
            assert (y.sum().compute() == want.sum()) and np.array_equal(y.max(axis=0).compute(), want.max(axis=0)), idx
    f = rng.random((64, 48))
    fd = da.from_array(f, chunks=(16, 12)).persist()
    assert np.array_equal((fd[::-1] - fd).compute(), f[::-1] - f)                      # unification rechunks the view
    assert np.array_equal(fd[::2, ::-2].T.compute(), f[::2, ::-2].T)
    assert np.array_equal(fd[::-1].cumsum(axis=0).compute().round(9), np.cumsum(f[::-1], axis=0).round(9))


def test_newaxis_in_index():
    import dask_array_b200 as da
    xh = np.arange(60.0).reshape(3, 4, 5)
    x = da.from_array(xh, chunks=(2, 2, 5)).persist()
    for idx in [np.s_[None], np.s_[:, None], np.s_[..., None], np.s_[1, None, :, None], np.s_[None, ..., None, 2],
                np.s_[::-1, None, 1:3]]:
        assert np.array_equal(x[idx].compute(), xh[idx]), idx
    v = da.from_array(xh[0, 0], chunks=2).persist()
    assert np.array_equal((v[:, None] * v[None, :]).compute(), xh[0, 0][:, None] * xh[0, 0][None, :])
