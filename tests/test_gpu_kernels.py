"""Kernel-level GPU tests: FusedLaunch / combine / gather against NumPy on the same inputs.

These exercise the C ABI directly (one level below the expression front-end)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(a):
    from dask_array_b200._device import DeviceChunk
    return DeviceChunk.from_numpy(a)


def _chain_program(dtype="float32"):
    from dask_array_b200 import _codegen as cg
    p = cg.Program()
    x = p.add_input(dtype)
    p.set_output(p.op("add", p.op("multiply", p.op("sin", x), p.const(2)), p.op("power", x, p.const(2))))
    return p


def _identity(dtype):
    from dask_array_b200 import _codegen as cg
    p = cg.Program()
    x = p.add_input(dtype)
    p.set_output(p.op("positive", x))
    return p


def _run_reduce(program, redop, axes, arrays, out_shape, out_dtype, acc_dtype=None, out1=False, **bk):
    import torch
    from dask_array_b200 import _runtime as rt
    from dask_array_b200._device import DeviceChunk
    chunks = [_mk(a) for a in arrays]
    o0 = DeviceChunk.empty(out_shape, out_dtype)
    o1 = DeviceChunk.empty(out_shape, np.int64) if out1 else None
    blk = rt.BlockArgs(shape=chunks[0].shape if chunks else out_shape,
                       inputs=[(c.ptr, c.strides) for c in chunks], out0=o0.ptr,
                       out1=o1.ptr if o1 else 0, **bk)
    L = rt.FusedLaunch(program, redop, axes, [blk], acc_dtype=acc_dtype)
    L.run(); L.run()   # twice: the self-resetting counters must allow re-launch
    torch.cuda.synchronize()
    return (o0.to_numpy(), o1.to_numpy()) if out1 else o0.to_numpy()


@pytest.mark.parametrize("shape", [(512, 1024), (1000, 1000), (37, 5), (4096, 4096)])
def test_fused_chain_sum_axis0(shape):
    from dask_array_b200 import _lib
    x = np.random.default_rng(0).random(shape, dtype=np.float32)
    got = _run_reduce(_chain_program(), _lib.RED_SUM, (0,), [x], (1, shape[1]), np.float32)
    want = np.sum(np.sin(x) * 2 + x**2, axis=0, keepdims=True, dtype=np.float32)
    truth = np.sum(np.sin(x.astype(np.float64)) * 2 + x.astype(np.float64) ** 2, axis=0, keepdims=True)
    np.testing.assert_allclose(got, want, rtol=1e-5)
    np.testing.assert_allclose(got, truth, rtol=2e-6)


@pytest.mark.parametrize("shape", [(512, 1024), (1000, 1000), (37, 5), (2048, 4096)])
def test_fused_chain_moment_all(shape):
    from dask_array_b200 import _lib
    x = np.random.default_rng(1).random(shape, dtype=np.float32)
    got = _run_reduce(_chain_program(), _lib.RED_MOMENT, (0, 1), [x], (3,), np.float64, acc_dtype=np.float32)
    y = np.sin(x.astype(np.float64)) * 2 + x.astype(np.float64) ** 2
    assert got[0] == y.size
    np.testing.assert_allclose(got[1], y.mean(), rtol=1e-6)
    np.testing.assert_allclose(got[2], ((y - y.mean()) ** 2).sum(), rtol=1e-5)


@pytest.mark.parametrize("dtype", ["float64", "float32", "int32"])
@pytest.mark.parametrize("shape", [(64, 16384), (100, 1000), (33, 7)])
def test_arg_minmax_axis1_bit_exact(dtype, shape):
    from dask_array_b200 import _lib
    rng = np.random.default_rng(2)
    x = (rng.random(shape) * 50).astype(dtype)
    x[:, : shape[1] // 2] = np.floor(x[:, : shape[1] // 2])          # plenty of ties
    if np.dtype(dtype).kind == "f":
        x[1, 3] = np.nan
        x[1, 5] = np.nan
        x[2, 0] = np.inf
        x[3, 1] = -np.inf
    for redop, fn, vfn in [(_lib.RED_ARGMAX, np.argmax, np.max), (_lib.RED_ARGMIN, np.argmin, np.min)]:
        v, i = _run_reduce(_identity(dtype), redop, (1,), [x], (shape[0], 1), dtype, out1=True, arg_offset=7)
        np.testing.assert_array_equal(i[:, 0], fn(x, axis=1) + 7)
        np.testing.assert_array_equal(v[:, 0], vfn(x, axis=1))
    for redop, vfn in [(_lib.RED_MAX, np.max), (_lib.RED_MIN, np.min)]:
        v = _run_reduce(_identity(dtype), redop, (1,), [x], (shape[0], 1), dtype)
        np.testing.assert_array_equal(v[:, 0], vfn(x, axis=1))


def test_arg_ravel_and_axis0():
    from dask_array_b200 import _lib
    rng = np.random.default_rng(3)
    x = np.floor(rng.random((300, 700)) * 1000)
    v, i = _run_reduce(_identity("float64"), _lib.RED_ARGMAX, (0, 1), [x], (1, 1), "float64", out1=True,
                       arg_ravel=((300, 700), (600, 1400), (2000, 3000)))
    r, c = np.unravel_index(np.argmax(x), x.shape)
    assert i[0, 0] == (r + 600) * 3000 + (c + 1400)
    assert v[0, 0] == x.max()
    v, i = _run_reduce(_identity("float64"), _lib.RED_ARGMIN, (0,), [x], (1, 700), "float64", out1=True, arg_offset=11)
    np.testing.assert_array_equal(i[0], np.argmin(x, axis=0) + 11)


def test_elementwise_transposed_and_broadcast():
    import torch
    from dask_array_b200 import _codegen as cg, _lib, _runtime as rt
    from dask_array_b200._device import DeviceChunk
    rng = np.random.default_rng(4)
    a = rng.integers(-1000, 1000, (384, 256)).astype(np.int32)
    row = rng.integers(-5, 5, (1, 384)).astype(np.int64)
    p = cg.Program()
    xa, xt, xr = p.add_input("int32"), p.add_input("int32"), p.add_input("int64")
    p.set_output(p.op("add", p.op("multiply", xa, xr), p.op("floor_divide", xt, p.const(7))))
    da_, dr = _mk(a), _mk(row)
    sq = _mk(np.ascontiguousarray(a[:256, :256]))
    out = DeviceChunk.empty((256, 256), np.int64)
    blk = rt.BlockArgs(shape=(256, 256),
                       inputs=[(sq.ptr, sq.strides), (sq.T.ptr, sq.T.strides),
                               (dr.ptr, (0, 1))], out0=out.ptr)
    rt.FusedLaunch(p, _lib.RED_NONE, (), [blk]).run()
    torch.cuda.synchronize()
    s = a[:256, :256]
    np.testing.assert_array_equal(out.to_numpy(), s * row[:, :256] + s.T // 7)


def test_combine_and_gather():
    import torch
    from dask_array_b200 import _lib, _runtime as rt
    from dask_array_b200._device import DeviceChunk
    rng = np.random.default_rng(5)
    parts = [rng.random(1000).astype(np.float32) for _ in range(8)]
    d = [_mk(p) for p in parts]
    out = DeviceChunk.empty((1000,), np.float32)
    keep = rt.combine(_lib.RED_SUM, np.float32, [c.ptr for c in d], None, 1000, out.ptr,
                      post=_lib.POST_MEAN, out_dtype=np.float32, count=8 * 4096)
    torch.cuda.synchronize()
    acc = parts[0].copy()
    for p in parts[1:]:
        acc = acc + p
    np.testing.assert_array_equal(out.to_numpy(), acc / np.float32(8 * 4096))
    # gather: re-block (64, 48) int16 column panels into row panels
    src = rng.integers(0, 30000, (64, 48)).astype(np.int16)
    s = _mk(src)
    dst = DeviceChunk.empty((64, 48), np.int16)
    copies = []
    for j in range(0, 48, 16):
        for i in range(0, 64, 16):
            copies.append((s.ptr + (i * 48 + j) * 2, dst.ptr + (i * 48 + j) * 2, 16, 32, 96, 96))
    rt.GatherLaunch(copies).run()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dst.to_numpy(), src)
    del keep


def test_fast_sin_cos_accuracy_and_slow_path():
    """b2_sinf_fast / b2_cosf_fast (csrc/b2_device.cuh): <= 2 ulp of the correctly rounded value
    on the Cody-Waite range, libdevice beyond it (per-batch slow path), NaN/inf semantics."""
    import torch
    from dask_array_b200 import _codegen as cg, _lib, _runtime as rt
    from dask_array_b200._device import DeviceChunk
    rng = np.random.default_rng(11)
    parts = [
        rng.random(1 << 20, dtype=np.float32) * 2 - 1,
        (rng.random(1 << 20, dtype=np.float32) - 0.5) * 200,
        (rng.random(1 << 20, dtype=np.float32) - 0.5) * 2e5,               # straddles 105615
        (rng.random(1 << 18, dtype=np.float32) - 0.5) * 1e9,               # far beyond: libdevice path
        np.float32(np.pi / 2) * rng.integers(-40000, 40000, 1 << 18).astype(np.float32),   # near zeros/extrema
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, -1e-30, 105615.0, 105616.0, 3.4e38], np.float32),
    ]
    x = np.concatenate(parts)
    x = np.concatenate([x, np.zeros((-len(x)) % 4096, np.float32)]).reshape(-1, 4096)
    for fn, npf in (("sin", np.sin), ("cos", np.cos)):
        p = cg.Program()
        p.set_output(p.op(fn, p.add_input("float32")))
        c = DeviceChunk.from_numpy(x)
        out = DeviceChunk.empty(x.shape, np.float32)
        rt.FusedLaunch(p, _lib.RED_NONE, (), [rt.BlockArgs(shape=x.shape, inputs=[(c.ptr, c.strides)], out0=out.ptr)]).run()
        torch.cuda.synchronize()
        got = out.to_numpy()
        with np.errstate(invalid="ignore"):
            want64 = npf(x.astype(np.float64))
            want32 = want64.astype(np.float32)
            ulp = np.spacing(np.abs(want32))
        fin = np.isfinite(want64)
        assert np.array_equal(np.isnan(got), ~fin)
        err = np.abs(got[fin].astype(np.float64) - want64[fin]) / ulp[fin]
        assert err.max() <= 2.0, (fn, err.max())
        assert got.ravel()[np.flatnonzero(x.ravel() == 0)[1]] == npf(np.float32(-0.0))
