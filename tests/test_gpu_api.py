"""API-level parity on the GPU: ``import dask_array_b200 as da`` against the oracle (the NumPy
restatement of the reference's threaded compute) on the same seeded inputs.

Tolerances (BASELINE.json north_star): slicing, rechunk, integer ops, min/max and arg
reductions bit-exact; fp32 reductions rtol 1e-5; fp64 reductions rtol 1e-12."""
import numpy as np
import pytest

from oracle import reference as ref

pytestmark = pytest.mark.gpu

RTOL32, RTOL64 = 1e-5, 1e-12


@pytest.fixture(scope="module")
def da():
    import dask_array_b200 as da
    return da


def test_readme_example(da):
    x = da.ones((1000, 1000), chunks=(100, 100))
    y = (x + x.T)[:100, :100]
    out = y.compute()
    assert out.shape == (100, 100) and out.dtype == np.float64
    assert np.array_equal(out, np.full((100, 100), 2.0))
    s = (x + x.T).sum().compute()
    assert s == 2_000_000.0 and s.dtype == np.float64


@pytest.mark.parametrize("shape,chunks", [((1024, 768), (256, 256)), ((1000, 700), (300, 256)), ((513, 129), (128, 64))])
def test_fused_chain_mean_std(da, shape, chunks):
    xh = np.random.default_rng(0).random(shape, dtype=np.float32)
    x = da.from_array(xh, chunks=chunks)
    y = da.sin(x) * 2 + x**2
    want_mean, want_std = ref.fused_chain_mean_std(xh, chunks)
    got_mean, got_std = y.mean(axis=0).compute(), y.std().compute()
    assert got_mean.dtype == want_mean.dtype == np.float32 and got_mean.shape == want_mean.shape
    np.testing.assert_allclose(got_mean, want_mean, rtol=RTOL32)
    np.testing.assert_allclose(got_std, want_std, rtol=RTOL32)
    # against an fp64 ground truth as well
    y64 = np.sin(xh.astype(np.float64)) * 2 + xh.astype(np.float64) ** 2
    np.testing.assert_allclose(got_mean, y64.mean(axis=0), rtol=RTOL32)
    np.testing.assert_allclose(got_std, y64.std(), rtol=RTOL32)
    # the element-wise chain itself: +,-,* exact, sin within 2 ulp of NumPy's
    got = y.compute()
    np.testing.assert_allclose(got, np.sin(xh) * 2 + xh**2, rtol=3e-7)


@pytest.mark.parametrize("dtype", ["float64", "float32", "int32", "int64"])
def test_reductions_2d(da, dtype):
    rng = np.random.default_rng(1)
    # floats: positive data, so that sum / mean are well conditioned and north_star's rtol (1e-5 / 1e-12)
    # applies to the RESULT (the reference's own assert_eq is allclose(rtol=1e-5, atol=1e-8)); sums with
    # cancellation are covered by test_float_sums_with_cancellation below with the bound stated there
    xh = (rng.random((300, 420)) * 200 - (100 if np.dtype(dtype).kind == "i" else 0)).astype(dtype)
    chunks = (128, 100)
    x = da.from_array(xh, chunks=chunks)
    b = ref.Blocked.from_array(xh, chunks)
    rtol = RTOL64 if dtype != "float32" else RTOL32
    for axis in (None, 0, 1, (0, 1)):
        for kd in (False, True):
            got = x.sum(axis=axis, keepdims=kd).compute()
            want = ref.da_sum(b, axis=axis, keepdims=kd)
            assert got.dtype == want.dtype and got.shape == np.shape(want)
            if np.dtype(dtype).kind == "i":
                assert np.array_equal(got, want)
            else:
                np.testing.assert_allclose(got, want, rtol=rtol, atol=1e-8)
            got, want = x.mean(axis=axis, keepdims=kd).compute(), ref.da_mean(b, axis=axis, keepdims=kd)
            assert got.dtype == want.dtype
            np.testing.assert_allclose(got, want, rtol=rtol, atol=1e-8)
            for ddof in (0, 1):
                got = x.var(axis=axis, keepdims=kd, ddof=ddof).compute()
                want = ref.da_var(b, axis=axis, keepdims=kd, ddof=ddof)
                assert got.dtype == want.dtype
                np.testing.assert_allclose(got, want, rtol=rtol)
            np.testing.assert_allclose(x.std(axis=axis, keepdims=kd).compute(), ref.da_std(b, axis=axis, keepdims=kd),
                                       rtol=rtol)
            assert np.array_equal(x.min(axis=axis, keepdims=kd).compute(), ref.da_min(b, axis=axis, keepdims=kd))
            assert np.array_equal(x.max(axis=axis, keepdims=kd).compute(), ref.da_max(b, axis=axis, keepdims=kd))


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_float_sums_with_cancellation(da, dtype):
    """Signed data: the sum is small against sum|x|, so the rounding error of ANY summation order (the
    reference's included) is bounded relative to sum|x|, not to the result: |got - truth| <= rtol * sum|x|,
    truth in fp64 (fp32) / extended precision via math.fsum (fp64)."""
    import math
    rng = np.random.default_rng(1)
    xh = (rng.random((300, 420)) * 200 - 100).astype(dtype)
    x = da.from_array(xh, chunks=(128, 100))
    rtol = RTOL64 if dtype == "float64" else RTOL32
    got = x.sum().compute()
    truth = math.fsum(xh.astype(np.float64).ravel())
    assert abs(float(got) - truth) <= rtol * float(np.abs(xh.astype(np.float64)).sum())
    for axis in (0, 1):
        got = x.sum(axis=axis).compute().astype(np.float64)
        truth = xh.astype(np.longdouble).sum(axis=axis).astype(np.float64)
        assert np.all(np.abs(got - truth) <= rtol * np.abs(xh.astype(np.float64)).sum(axis=axis))
        gm = x.mean(axis=axis).compute().astype(np.float64)
        assert np.all(np.abs(gm - truth / xh.shape[axis]) <= rtol * np.abs(xh.astype(np.float64)).mean(axis=axis))


@pytest.mark.parametrize("offset,spread", [(1000.0, 10.0), (1.0e6, 1.0)])
@pytest.mark.parametrize("chunks", [(64, 64), (100, 37)])
def test_fp32_variance_with_mean_far_above_spread(da, offset, spread, chunks):
    """var / std of fp32 data whose mean dwarfs its spread (the oracle golden case `1000 + 10 u`,
    tests/golden/generate.py, and a harsher 1e6 + u): a naive single pass sum(x^2) - n mean^2 loses every
    digit here; the kernel's shifted sums + fp64 Chan merge must stay within rtol 1e-5 of the fp64 truth, and
    at least as close to it as the reference's two-pass-per-chunk result (moment_chunk, _common.py:393-403)."""
    rng = np.random.default_rng(5)
    xh = (offset + spread * rng.random((256, 300))).astype(np.float32)
    x = da.from_array(xh, chunks=chunks)
    b = ref.Blocked.from_array(xh, chunks)
    x64 = xh.astype(np.float64)
    for axis in (None, 0, 1):
        got = x.var(axis=axis).compute()
        truth = x64.var(axis=axis)
        np.testing.assert_allclose(got, truth, rtol=RTOL32)
        # The reference's fp32 two-pass-per-chunk result is itself off by 2e-5 (1000 + 10 u) to 1.6e-2
        # (1e6 + u) here -- its block mean is rounded to fp32 before the deviations are squared (measured on
        # B200 box, round 2) -- so it cannot be matched to 1e-5; what must hold is that this backend is at
        # least as close to the fp64 truth as the reference is, and agrees with it within the reference's
        # own error.
        want = ref.da_var(b, axis=axis).astype(np.float64)
        ref_err = np.abs(want - truth)
        assert np.all(np.abs(got - truth) <= ref_err + RTOL32 * np.abs(truth))
        assert np.all(np.abs(got - want) <= 2 * ref_err + RTOL32 * np.abs(truth))
        np.testing.assert_allclose(x.std(axis=axis, ddof=1).compute(), x64.std(axis=axis, ddof=1), rtol=RTOL32)
    # through a fused chain as well (the chunk step of the std() headline kernel)
    y = x * 2 + 1
    np.testing.assert_allclose(y.std().compute(), (x64 * 2 + 1).std(), rtol=RTOL32)


def test_random_matches_seed_sequence_spawn_bit_exact(da):
    """da.random.default_rng(seed).random(...) == the reference's per-block streams (random/_expr.py:29-32,
    97-126): block k draws from PCG64(SeedSequence(seed).spawn(nblocks)[k]); a second draw from the same
    generator continues the spawn sequence (different values, reproducible)."""
    rng = da.random.default_rng(42)
    a = rng.random((96, 80), chunks=(32, 40), dtype=np.float32)
    b = rng.random((96, 80), chunks=(32, 40))
    ss = np.random.SeedSequence(42)
    for arr, dt in ((a, np.float32), (b, np.float64)):
        kids = ss.spawn(6)
        want = np.empty((96, 80), dtype=dt)
        for k, (i, j) in enumerate((i, j) for i in range(3) for j in range(2)):
            want[32 * i:32 * i + 32, 40 * j:40 * j + 40] = np.random.Generator(np.random.PCG64(kids[k])).random((32, 40), dtype=dt)
        got = arr.compute()
        assert got.dtype == dt and np.array_equal(got, want)
    assert not np.array_equal(a.compute().astype(np.float64), b.compute())
    again = da.random.default_rng(42).random((96, 80), chunks=(32, 40), dtype=np.float32)
    assert again.name == a.name and np.array_equal(again.compute(), a.compute())
    assert (da.random.random((8,), chunks=4) - da.random.random((8,), chunks=4)).compute().any()
    n = da.random.default_rng(7).standard_normal((64,), chunks=16)
    kids = np.random.SeedSequence(7).spawn(4)
    assert np.array_equal(n.compute(), np.concatenate([np.random.Generator(np.random.PCG64(k)).standard_normal(16) for k in kids]))


@pytest.mark.parametrize("dtype", ["float64", "float32", "int16"])
@pytest.mark.parametrize("split_every", [None, 2])
def test_arg_reductions_bit_exact(da, dtype, split_every):
    rng = np.random.default_rng(2)
    xh = np.floor(rng.random((260, 330)) * 40).astype(dtype)     # many ties
    if np.dtype(dtype).kind == "f":
        xh[5, 7] = np.nan; xh[5, 200] = np.nan; xh[9, 3] = np.inf; xh[11, 300] = -np.inf; xh[200, :] = 3.0
    chunks = (100, 64)
    x = da.from_array(xh, chunks=chunks)
    b = ref.Blocked.from_array(xh, chunks)
    for axis in (None, 0, 1):
        for fn, rfn in (("argmax", ref.da_argmax), ("argmin", ref.da_argmin)):
            got = getattr(x, fn)(axis=axis, split_every=split_every).compute()
            want = rfn(b, axis=axis, split_every=split_every)
            assert got.dtype == np.int64
            assert np.array_equal(got, want), (fn, axis)
            if axis is not None:
                # (for axis=None the reference breaks cross-block ties by BLOCK order --
                # _arg_combine re-runs argmax over the concatenated block maxima -- so it can
                # differ from np.argmax on tied data; the oracle reproduces the reference)
                assert np.array_equal(got, getattr(np, fn)(xh, axis=axis))
    assert np.array_equal(x.max(axis=1).compute(), xh.max(axis=1), equal_nan=True)
    assert np.array_equal(x.min(axis=0).compute(), xh.min(axis=0), equal_nan=True)


@pytest.mark.parametrize("dtype", ["float32", "int32", "float64", "uint8"])
def test_rechunk_and_transpose_bit_exact(da, dtype):
    n = 512
    xh = np.arange(n * n).reshape(n, n).astype(dtype)
    # persisted (device-resident) sources: the rechunk is the gather kernel; un-persisted host
    # sources absorb the rechunk into their upload (reference rechunk pushdown) -- both are checked
    for x in (da.from_array(xh, chunks=(n, 16)).persist(), da.from_array(xh, chunks=(n, 16))):
        r = x.rechunk((16, n))
        assert r.chunks == ((16,) * 32, (n,))
        assert np.array_equal(r.compute(), xh)
        want = ref.rechunk(ref.Blocked.from_array(xh, (n, 16)), (16, n))
        assert np.array_equal(r.compute(), want.to_array())
    x = da.from_array(xh, chunks=(n, 16)).persist()
    # ragged both ways
    for xr in (da.from_array(xh[:500, :300], chunks=(130, 70)).persist(), da.from_array(xh[:500, :300], chunks=(130, 70))):
        assert np.array_equal(xr.rechunk((64, 300)).compute(), xh[:500, :300])
        assert np.array_equal(xr.rechunk({0: 499}).compute(), xh[:500, :300])
    # x.T + x (square chunks: fused transpose) and non-square chunks (rechunk inserted)
    sq = da.from_array(xh, chunks=(128, 128))
    assert np.array_equal((sq.T + sq).compute(), xh.T + xh)
    assert np.array_equal((x.T + x).compute(), xh.T + xh)
    assert np.array_equal(sq.T.compute(), xh.T)


def test_integer_ops_bit_exact(da):
    rng = np.random.default_rng(3)
    ah = rng.integers(-1000, 1000, (200, 300)).astype(np.int32)
    bh = rng.integers(1, 50, (200, 300)).astype(np.int64)
    a, b = da.from_array(ah, chunks=(64, 128)), da.from_array(bh, chunks=(64, 128))
    cases = {
        "floordiv": (a // b, ah // bh), "mod": (a % b, ah % bh), "pow": (a ** 2, ah ** 2),
        "mix": ((a * 3 - b) // 7, (ah * 3 - bh) // 7), "truediv": (a / b, ah / bh),
        "cmp": ((a > 5) & (b < 30), (ah > 5) & (bh < 30)), "neg": (-a, -ah), "abs": (abs(a), abs(ah)),
        "shift": ((a << 3) >> 1, (ah << 3) >> 1), "where": (da.where(a > 0, a, b), np.where(ah > 0, ah, bh)),
        "astype": (a.astype("float32"), ah.astype("float32")), "xor": (a ^ 12345, ah ^ 12345),
        "bcast": (a + b[0:1, :], ah + bh[0:1, :]),
    }
    for name, (got, want) in cases.items():
        g = got.compute()
        assert g.dtype == want.dtype, (name, g.dtype, want.dtype)
        assert np.array_equal(g, want), name


def test_slicing_bit_exact(da):
    xh = np.arange(300 * 200, dtype=np.float64).reshape(300, 200)
    x = da.from_array(xh, chunks=(64, 64))
    for idx in [(slice(10, 250), slice(5, 199)), (slice(0, 64), slice(64, 128)), (7, slice(None)), (slice(100, 101), 3)]:
        assert np.array_equal(x[idx].compute(), xh[idx])
    y = (x + 1)[10:100, 20:30]
    assert np.array_equal(y.compute(), (xh + 1)[10:100, 20:30])
    assert np.array_equal((x + x)[:5].sum(axis=0).compute(), (xh + xh)[:5].sum(axis=0))


def test_persist_keeps_blocks_resident(da):
    from dask_array_b200 import _lib
    xh = np.random.default_rng(4).random((256, 256), dtype=np.float32)
    x = da.from_array(xh, chunks=(64, 64)).persist()
    before = _lib.launch_count()
    m1 = (x * 2).sum().compute()
    assert _lib.launch_count() > before
    np.testing.assert_allclose(m1, (xh * 2).sum(dtype=np.float32), rtol=RTOL32)
    np.testing.assert_allclose(x.mean().compute(), xh.mean(), rtol=RTOL32)


def test_unsupported_fails_loudly(da):
    x = da.from_array(np.zeros((8, 8)), chunks=(4, 4))
    with pytest.raises(NotImplementedError):
        x[[0, 2]]                                   # fancy indexing is outside the hot path
    with pytest.raises(NotImplementedError):
        da.elemwise("frexp", x).compute()


def test_cuda_graph_replay_matches_tape(da):
    """`Compiled.capture()`: the launch tape as one CUDA graph gives the same results, repeatedly."""
    rng = np.random.default_rng(4)
    xh = rng.random((512, 384), dtype=np.float32)
    x = da.from_array(xh, chunks=(128, 128)).persist()
    y = da.sin(x) * 2 + x**2
    step = da.compile(y.mean(axis=0), y.std(), (x.T[:, :384] + x[:384, :384].rechunk((96, 384))).sum(axis=1), x.argmax(axis=1))
    step.run()
    eager = [np.array(r) for r in step.results()]
    step.capture()
    for _ in range(3):
        step.run()
    for a, b in zip(eager, step.results()):
        assert np.array_equal(a, b, equal_nan=True)
    ones = da.ones((1000, 1000), chunks=(100, 100))
    s = da.compile((ones + ones.T).sum()).capture()
    s.run(); s.run()
    assert s.result() == 2_000_000.0


def test_graph_capture_with_parallel_lanes_replays_correctly(da):
    """Compiled.capture(): independent top-level expressions become parallel branches of one CUDA graph;
    expressions that share a computed intermediate stay on one stream.  Replays must give the same results."""
    rng = np.random.default_rng(11)
    xh = rng.random((512, 384), dtype=np.float32)
    x = da.from_array(xh, chunks=(128, 128)).persist()
    y = da.sin(x) * 2 + x**2
    step = da.compile(y.mean(axis=0), y.std(), x.max(axis=1))
    want = [r.copy() for r in step.results()]
    step.capture()
    assert step.lanes == 3
    for _ in range(4):
        step.run()
    for got, w in zip(step.results(), want):
        assert np.array_equal(got, w)
    shared = x.rechunk((256, 96))                    # a computed intermediate read by both expressions
    step2 = da.compile(shared.sum(axis=0), shared.min())
    want2 = [r.copy() for r in step2.results()]
    step2.capture()
    assert step2.lanes == 1
    step2.run(); step2.run()
    for got, w in zip(step2.results(), want2):
        assert np.array_equal(got, w)


def test_two_output_ufuncs_and_where_out(da):
    """frexp / modf / divmod (_ufunc.py:429-451) and elemwise(where=, out=) (_core_utils.py:1031-1036)."""
    xh = np.array([1.5, -2.0, 0.0, -0.0, np.inf, -np.inf, np.nan, 1e-310, 123456.789, -0.3], dtype=np.float64)
    x = da.from_array(xh, chunks=4)
    for dt in (np.float64, np.float32):
        xx, hh = x.astype(dt), xh.astype(dt)
        m, e = da.frexp(xx)
        wm, we = np.frexp(hh)
        assert m.dtype == wm.dtype and e.dtype == we.dtype
        assert np.array_equal(m.compute(), wm, equal_nan=True) and np.array_equal(e.compute(), we)
        f, i = da.modf(xx)
        wf, wi = np.modf(hh)
        gf, gi = f.compute(), i.compute()
        assert np.array_equal(gf, wf, equal_nan=True) and np.array_equal(gi, wi, equal_nan=True)
        assert np.array_equal(np.signbit(gf), np.signbit(wf)) and np.array_equal(np.signbit(gi), np.signbit(wi))
    ih = np.array([7, -7, 9, -9, 0, 5], dtype=np.int32)
    q, r = da.divmod(da.from_array(ih, chunks=3), 4)
    wq, wr = np.divmod(ih, 4)
    assert np.array_equal(q.compute(), wq) and np.array_equal(r.compute(), wr)
    q2, r2 = np.divmod(x, 0.7)                                   # through __array_ufunc__
    with np.errstate(all="ignore"):
        wq2, wr2 = np.divmod(xh, 0.7)
    assert np.array_equal(q2.compute(), wq2, equal_nan=True)
    # where= with out=: unselected positions keep out's values; out is overwritten in place and returned
    ah = np.arange(12.0).reshape(3, 4)
    mask = (ah % 3 == 0)
    a = da.from_array(ah, chunks=(2, 2))
    out = da.full((3, 4), -1.0, chunks=(2, 2))
    ret = np.multiply(a, 10, out=out, where=da.from_array(mask, chunks=(2, 2)))
    assert ret is out
    want = np.full((3, 4), -1.0)
    np.multiply(ah, 10, out=want, where=mask)
    assert np.array_equal(out.compute(), want)
    out2 = da.zeros((3, 4), dtype=np.float32, chunks=(2, 2))
    np.sin(a, out=out2)                                          # out= alone: cast to out's dtype
    np.testing.assert_allclose(out2.compute(), np.sin(ah).astype(np.float32), rtol=3e-7)
    with pytest.raises(NotImplementedError, match="uninitialised"):
        np.add(a, 1, where=da.from_array(mask, chunks=(2, 2)))
    assert np.add(a, 1, dtype=np.float32).dtype == np.float32
