"""topk / argtopk on the GPU against the golden outputs of the reference's own chunk
functions (tests/golden/topk.npz) and the oracle.  Values bit-exact; indices compared on inputs with
distinct values (the order of equal elements is unspecified in the reference too)."""
import ast
import os

import numpy as np
import pytest

from oracle import reference as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def test_golden_cases(da):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "topk.npz"))
    for case in sorted({k.split("/")[0] for k in g.files}):
        chunks, k, axis = ast.literal_eval(str(g[case + "/meta"][0]))
        xh = g[case + "/x"]
        x = da.from_array(xh, chunks=chunks).persist()
        got = x.topk(k, axis=axis)
        assert got.chunks[axis] == (min(abs(k), xh.shape[axis]),)
        assert np.array_equal(got.compute(), g[case + "/topk"], equal_nan=True), case
        if case + "/argtopk" in g.files:
            assert np.array_equal(x.argtopk(k, axis=axis).compute(), g[case + "/argtopk"]), case


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64, np.uint8, np.int16])
@pytest.mark.parametrize("shape,chunks,axis,k", [((5000,), (1700,), 0, 8), ((64, 9000), (32, 4500), 1, -5),
                                                 ((300, 50), (100, 25), 0, 33), ((12, 40, 30), (6, 13, 30), 1, 2),
                                                 ((3, 100000), (3, 100000), -1, 1000), ((70,), (70,), 0, 70)])
def test_topk_matrix(da, dtype, shape, chunks, axis, k):
    rng = np.random.default_rng(abs(hash((shape, k))) % 2**32)
    n = int(np.prod(shape))
    if np.dtype(dtype).kind == "f":
        xh = (rng.permutation(n).astype(np.float64) / 3).astype(dtype).reshape(shape)
    else:
        info = np.iinfo(dtype)
        xh = rng.integers(info.min, info.max, size=shape, endpoint=True).astype(dtype)
    x = da.from_array(xh, chunks=chunks).persist()
    want = ref.da_topk(ref.Blocked.from_array(xh, chunks), k, axis).to_array()
    got = x.topk(k, axis=axis).compute()
    assert got.dtype == want.dtype and np.array_equal(got, want)
    idx = x.argtopk(k, axis=axis).compute()
    assert idx.dtype == np.intp
    assert np.array_equal(np.take_along_axis(xh, idx, axis % xh.ndim), want)      # valid for ties too


def test_topk_nan_counts_as_largest_and_replay(da):
    xh = np.random.default_rng(1).random((8, 1000))
    xh[2, 5] = np.nan
    xh[2, 700] = np.nan
    x = da.from_array(xh, chunks=(4, 250)).persist()
    step = da.compile(x.topk(3, axis=1), x.topk(-3, axis=1))
    step.run(); step.run()
    big, small = step.results()
    assert np.array_equal(big, np.sort(xh, axis=1)[:, ::-1][:, :3], equal_nan=True)     # np.sort puts NaN last
    assert np.array_equal(small, np.sort(xh, axis=1)[:, :3])
    with pytest.raises(ValueError):
        x.topk(0)
