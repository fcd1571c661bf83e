"""The reference-facing boundary (SURVEY.md 8b B1'/B1''/B2/B3, 8f-1): ``dask_array_b200.plugin``.

CPU part (runs everywhere): the expression adapter on stand-in objects shaped like the reference's classes
(same class names, same public attributes), ``FusedPlan.from_reference`` down to a compiled sm_100a cubin,
the scheduler's graph walking on host-only graphs, loud failure without a GPU.
Build-container part (needs /root/reference): ``register()`` against the reference's OWN modules, imported
through ``tests/golden/_refshim.py``, then the reference's own lookups resolve to this backend.
GPU part (``-m gpu``): ``get`` running a graph of the (oracle-restated) reference chunk functions on
``DeviceChunk`` blocks, the arg-reduction chunk/combine path with its ``np.ogrid`` gather, ``plugin.compute``
of a stand-in reference tree against NumPy.
"""
import functools
import operator
import os
import sys
import types

import numpy as np
import pytest

import dask_array_b200 as da
from dask_array_b200 import plugin
from dask_array_b200._blockwise import FusedPlan

HAVE_REF = os.path.exists("/root/reference/dask_array/_chunk_types.py")


# ----------------------------------------------------------------------------- stand-ins
def node(cls_name, **attrs):
    """An object whose class is NAMED like a reference expression class and exposes the same attributes."""
    obj = type(cls_name, (), {})()
    for k, v in attrs.items():
        setattr(obj, k, v)
    if "_name" not in attrs:
        obj._name = f"{cls_name.lower()}-{id(obj):x}"
    return obj


def ref_from_array(a, chunks):
    return node("FromArray", array=a, chunks=chunks, dtype=a.dtype, shape=a.shape, ndim=a.ndim)


def ref_elemwise(op, *args, dtype=None, **user_kwargs):
    arrs = [a for a in args if hasattr(a, "chunks")]
    shape = np.broadcast_shapes(*[a.shape for a in arrs])
    return node("Elemwise", op=op, elemwise_args=tuple(args), user_kwargs=user_kwargs, where=True, out=None,
                dtype=dtype, shape=shape, ndim=len(shape), chunks=arrs[0].chunks)


def ref_chain(xh, chunks):
    x = ref_from_array(xh, chunks)
    s = ref_elemwise(np.sin, x)
    m = ref_elemwise(operator.mul, s, 2)
    p = ref_elemwise(operator.pow, x, 2)
    return x, ref_elemwise(operator.add, m, p), (s, m, p)


# ----------------------------------------------------------------------------- CPU: adapter
def test_lower_reference_elemwise_chain_and_typed_reductions():
    xh = np.random.default_rng(0).random((64, 48), dtype=np.float32)
    x, y, _ = ref_chain(xh, (32, 16))
    ours = plugin.lower_reference(y)
    want = da.from_array(xh, chunks=(32, 16))
    want = da.sin(want) * 2 + want**2
    assert ours.dtype == np.float32 and ours.chunks == want.chunks
    # same structure as the expression our own API builds: same fused program
    a = FusedPlan(ours.optimize()) if type(ours.optimize()).__name__ == "FusedBlockwise" else None
    b = FusedPlan(want.expr.optimize())
    assert a is not None and a.program.key() == b.program.key()
    mean = node("Mean", array=y, axis=(0,), keepdims=False, dtype=np.dtype("f4"), split_every=None, aggregate=None,
                shape=(48,), ndim=1, chunks=((16,) * 3,))
    lowered = plugin.lower_reference(mean)
    assert type(lowered).__name__ == "Reduction" and lowered.operand("kind") == "mean" and lowered.shape == (48,)
    var = node("Var", array=y, axis=None, keepdims=False, dtype=np.dtype("f4"), split_every={0: 2, 1: 2},
               aggregate=functools.partial(lambda *a, **k: None, ddof=1), shape=(), ndim=0, chunks=())
    lv = plugin.lower_reference(var)
    assert lv.operand("ddof") == 1 and lv.operand("split_every") == {0: 2, 1: 2}
    def safe_sqrt(a): ...
    std = ref_elemwise(safe_sqrt, var)
    assert plugin.lower_reference(std).operand("op") == "sqrt"


def test_lower_reference_refuses_what_has_no_kernel():
    xh = np.zeros((4, 4))
    x = ref_from_array(xh, ((4,), (4,)))
    with pytest.raises(NotImplementedError, match="where"):
        plugin.lower_reference(node("Elemwise", op=np.add, elemwise_args=(x, x), user_kwargs={}, where=x, out=None,
                                    dtype=None, shape=(4, 4), ndim=2, chunks=x.chunks))
    with pytest.raises(NotImplementedError, match="Shuffle"):
        plugin.lower_reference(node("Shuffle", array=x))


def test_fused_plan_from_reference_compiles_one_kernel():
    """B2: a reference FusedBlockwise (exprs root first, external dependency = the leaf) -> ONE kernel program,
    compiled for sm_100a here (NVRTC needs no GPU)."""
    from dask_array_b200 import _codegen as cg, _lib, _runtime as rt

    xh = np.random.default_rng(1).random((64, 64), dtype=np.float32)
    x, y, (s, m, p) = ref_chain(xh, (32, 32))
    fused = node("FusedBlockwise", exprs=(y, m, s, p))
    plan = FusedPlan.from_reference(fused)
    assert [n for n, *_ in plan.program.ops] == ["sin", "multiply", "power", "add"] or len(plan.program.ops) == 4
    assert len(plan.leaves) == 1 and type(plan.leaves[0][0]).__name__ == "FromArray"
    geo = cg.choose_geometry(plan.program, _lib.MODE_EW, [(1, 32, 32)], 4)
    spec = cg.KernelSpec(plan.program.key(), ("V",), _lib.MODE_EW, _lib.RED_NONE, acc_dtype="float32", **geo)
    assert len(rt.compile_kernel(plan.program, spec)) > 1000
    # with a transposed member: a + a.T (tests/test_collection.py:996-1135 fusion cases)
    t = node("Transpose", array=x, axes=(1, 0), shape=(64, 64), ndim=2, chunks=x.chunks, dtype=x.dtype)
    z = ref_elemwise(operator.add, x, t)
    plan = FusedPlan.from_reference(node("FusedBlockwise", exprs=(z, t)))
    assert len(plan.leaves) == 2                     # the same array read through two dimension maps


def test_lower_reference_lowered_tree_by_function_names():
    """The reference's LOWERED reduction (Blockwise chunk step + PartialReduce levels, _reduction.py:154-226,
    751-806) is recognised by the names of its chunk / combine / aggregate functions."""
    def mean_chunk(x, **k): ...
    def mean_combine(x, **k): ...
    def mean_agg(x, **k): ...
    def _concatenate2(x, **k): ...

    class Compose:                                    # toolz.functoolz.Compose: .first then .funcs
        def __init__(self, *fs):
            self.first, self.funcs = fs[-1], tuple(reversed(fs[:-1]))

    xh = np.random.default_rng(2).random((64, 64))
    x = ref_from_array(xh, ((16,) * 4, (16,) * 4))
    chunk = node("Blockwise", func=functools.partial(mean_chunk, dtype="f8"), kwargs={"axis": (0,), "keepdims": True},
                 args=(x, (0, 1)), chunks=((1,) * 4, (16,) * 4))
    comb = node("PartialReduce", array=chunk, split_every={0: 2},
                func=Compose(functools.partial(mean_combine, axis=(0,), keepdims=True), functools.partial(_concatenate2, axes=[0])),
                keepdims=True, dtype="f8", name="mean_combine-partial")
    agg = node("PartialReduce", array=comb, split_every={0: 2},
               func=Compose(functools.partial(mean_agg, axis=(0,), keepdims=False), functools.partial(_concatenate2, axes=[0])),
               keepdims=False, dtype="f8", name="mean_agg-aggregate")
    ours = plugin.lower_reference(agg)
    assert type(ours).__name__ == "PartialReduce" and ours.operand("final") and ours.operand("kind") == "mean"
    inner = ours.operand("array")
    assert not inner.operand("final") and type(inner.operand("array")).__name__ == "ChunkReduce"
    assert ours.shape == (64,) and ours.dtype == np.float64
    want = da.from_array(xh, chunks=(16, 16)).mean(axis=0, split_every=2).expr.lower_completely() \
        if hasattr(da.from_array(xh, chunks=(16, 16)).expr, "lower_completely") else None
    if want is not None:
        assert want.chunks == ours.chunks


def _ref_contraction(kind, ah, bh, achunks, bchunks, axes=None):
    """Stand-ins for the reference's contraction lowering (linalg/_tensordot.py:100-136, 253-334)."""
    def _tensordot(a, b, axes=None, is_sparse=False): ...
    def _matmul(a, b): ...
    def _chunk_sum(a, axis=None, dtype=None, keepdims=None): ...

    a, b = ref_from_array(ah, achunks), ref_from_array(bh, bchunks)
    if kind == "tensordot":
        la, lb = axes
        ndim = ah.ndim + bh.ndim - len(lb)
        inner = node("Blockwise", func=_tensordot, args=(a, tuple(range(ah.ndim)), b, None), ndim=ndim,
                     kwargs={"axes": (la, lb), "is_sparse": False}, concatenate=False, dtype=np.result_type(ah, bh))
        return node("Sum", array=inner, axis=tuple(la), keepdims=False, dtype=np.result_type(ah, bh), split_every=None,
                    aggregate=None)
    inner = node("Blockwise", func=_matmul, args=(a, (0, 1), b, (1, 2)), ndim=3, kwargs=None, concatenate=False,
                 dtype=np.result_type(ah, bh))
    if len(a.chunks[1]) == 1:
        return node("Squeeze", array=inner, axis=(1,), dtype=inner.dtype)
    return node("Reduction", array=inner, chunk=_chunk_sum, aggregate=_chunk_sum, axis=(1,), keepdims=False,
                dtype=inner.dtype, concatenate=False)


def test_lower_reference_contractions_become_one_accumulating_node():
    rng = np.random.default_rng(7)
    ah, bh = rng.random((64, 48)), rng.random((48, 32))
    mm = plugin.lower_reference(_ref_contraction("matmul", ah, bh, ((32, 32), (16,) * 3), ((16,) * 3, (32,))))
    assert type(mm).__name__ == "BlockContract" and mm.shape == (64, 32) and mm.dtype == np.float64
    one_k = plugin.lower_reference(_ref_contraction("matmul", ah, bh, ((32, 32), (48,)), ((48,), (32,))))
    assert type(one_k).__name__ == "BlockContract" and one_k.shape == (64, 32)
    f32 = plugin.lower_reference(_ref_contraction("matmul", ah.astype("f4"), bh.astype("f4"), ((32, 32), (16,) * 3),
                                                  ((16,) * 3, (32,))))
    assert type(f32).__name__ == "BlockGEMM" and f32.dtype == np.float32          # tensor-core path
    th, uh = rng.random((8, 6, 10)), rng.random((10, 6, 4))
    td = plugin.lower_reference(_ref_contraction("tensordot", th, uh, ((4, 4), (6,), (5, 5)), ((5, 5), (6,), (4,)),
                                                 axes=((2, 1), (0, 1))))
    assert type(td).__name__ == "BlockContract" and td.shape == (8, 4)
    assert td.operand("la") == (2, 1) and td.operand("lb") == (0, 1)
    # a Sum over the partial that is NOT the contraction fold is refused, not silently mis-lowered
    bad = _ref_contraction("tensordot", th, uh, ((4, 4), (6,), (5, 5)), ((5, 5), (6,), (4,)), axes=((2, 1), (0, 1)))
    bad.axis = (0,)
    with pytest.raises(NotImplementedError, match="contraction fold"):
        plugin.lower_reference(bad)
    # stacked matmul has no kernel
    s3 = _ref_contraction("matmul", rng.random((2, 4, 4)), rng.random((2, 4, 4)), ((2,), (4,), (2, 2)), ((2,), (2, 2), (4,)))
    with pytest.raises(NotImplementedError, match="batch matmul"):
        plugin.lower_reference(s3)


def _ref_window(xh, chunks, window, axis, reducer, keepdims=False, dtype=None):
    x = ref_from_array(xh, chunks)
    return node("SlidingWindowReduction", array=x, window=window, sliding_axis=axis, window_axis=xh.ndim,
                keepdims=keepdims, reducer=reducer, dtype=np.dtype(dtype or xh.dtype))


def test_lower_reference_window_reductions():
    xh = np.random.default_rng(8).random((40, 64), dtype=np.float32)
    sw = plugin.lower_reference(_ref_window(xh, ((20, 20), (32, 32)), 5, 1, "sum"))
    assert type(sw).__name__ == "SlidingWindowReduction" and sw.shape == (40, 60) and sw.chunks == ((20, 20), (32, 28))
    low = sw.lower_completely() if hasattr(sw, "lower_completely") else sw._lower()
    assert type(low).__name__ == "WindowReduce"
    with pytest.raises(NotImplementedError, match="no B200 kernel"):
        plugin.lower_reference(_ref_window(xh, ((20, 20), (32, 32)), 5, 1, "var"))
    mv = node("MovingWindowReduction", array=ref_from_array(xh, ((20, 20), (32, 32))), window=4, min_count=2,
              sliding_axis=1, reducer="move_mean", dtype=np.dtype("f4"))
    out = plugin.lower_reference(mv)
    assert out.shape == xh.shape and out.dtype == np.float32 and out.chunks == ((20, 20), (32, 32))


def _ref_views(xh, yh, chunks):
    x, y = ref_from_array(xh, chunks), ref_from_array(yh, chunks)
    cat = node("Concatenate", array=x, args=[x, y], axis=1, meta=None)
    stk = node("Stack", array=x, args=[x, y], axis=0, meta=None)
    exp = node("ExpandDims", array=x, axes=(0, 2))
    sqz = node("Squeeze", array=exp, axis=(0,))
    row = node("Slice", array=x, index=(slice(0, 1), slice(None)), allow_getitem_optimization=False)
    bto = node("BroadcastTo", array=row, _shape=(3,) + (1, xh.shape[1]), _chunks=((1,) * 3, (1,), chunks[1]),
               _meta_override=None)
    sl = node("Slice", array=x, index=(slice(3, None, 2), 5), allow_getitem_optimization=False)
    return dict(cat=cat, stk=stk, exp=exp, sqz=sqz, bto=bto, sl=sl)


def _ref_cum(xh, chunks, func, axis, method="sequential", dtype=None):
    x = ref_from_array(xh, chunks)
    if method == "blelloch":
        return node("CumReductionBlelloch", array=x, func=func, preop=np.sum, binop=operator.add, axis=axis, _dtype=dtype)
    return node("CumReduction", array=x, func=func, binop=operator.add, ident=0, axis=axis, _dtype=dtype)


def test_lower_reference_views_and_cumulative():
    xh, yh = np.arange(48.0).reshape(6, 8), -np.arange(48.0).reshape(6, 8)
    v = {k: plugin.lower_reference(e) for k, e in _ref_views(xh, yh, ((3, 3), (4, 4))).items()}
    assert v["cat"].shape == (6, 16) and v["cat"].chunks == ((3, 3), (4,) * 4)
    assert v["stk"].shape == (2, 6, 8) and v["stk"].chunks == ((1, 1), (3, 3), (4, 4))
    assert v["exp"].shape == (1, 6, 1, 8) and v["sqz"].shape == (6, 1, 8)
    assert v["bto"].shape == (3, 1, 8) and v["sl"].shape == (2,)
    def nancumsum(x, axis=None, dtype=None): ...
    for func, kind, nan in ((np.cumsum, "cumsum", False), (np.cumprod, "cumprod", False), (nancumsum, "cumsum", True)):
        for method in ("sequential", "blelloch"):
            c = plugin.lower_reference(_ref_cum(xh, ((3, 3), (4, 4)), func, 1, method))
            assert type(c).__name__ == "CumReduction" and c.operand("kind") == kind and c.operand("nan") is nan
            assert c.chunks == ((3, 3), (4, 4)) and c.dtype == np.float64
    c = plugin.lower_reference(_ref_cum(xh.astype("i4"), ((3, 3), (4, 4)), np.cumsum, 0, dtype="i8"))
    assert c.dtype == np.int64
    def cummax(x, axis=None): ...
    with pytest.raises(NotImplementedError, match="cumulative"):
        plugin.lower_reference(_ref_cum(xh, ((3, 3), (4, 4)), cummax, 0))
    x3 = ref_from_array(np.arange(120.0).reshape(6, 5, 4), ((3, 3), (2, 2, 1), (2, 2)))
    rs = plugin.lower_reference(node("Reshape", array=x3, _shape=(30, 4), chunks=((5,) * 6, (2, 2))))
    assert type(rs).__name__ == "Reshape" and rs.shape == (30, 4) and rs.chunks == ((5,) * 6, (2, 2))
    assert type(rs.operand("array")).__name__ == "Rechunk" and rs.operand("array").chunks == ((1,) * 6, (5,), (2, 2))
    ar = plugin.lower_reference(node("Arange", start=3, stop=40, step=4, chunks=((4, 4, 2),), like=None, dtype=np.dtype("i8")))
    assert ar.shape == (10,) and ar.chunks == ((4, 4, 2),) and ar.dtype == np.int64
    assert np.array_equal(np.concatenate([ar.operand("get_block")((k,)) for k in range(3)]), np.arange(3, 40, 4))
    ls = plugin.lower_reference(node("Linspace", start=0.0, stop=1.0, num=11, endpoint=True, chunks=((6, 5),), dtype=np.dtype("f8")))
    np.testing.assert_allclose(np.concatenate([ls.operand("get_block")((k,)) for k in range(2)]), np.linspace(0, 1, 11), rtol=1e-15)


def test_get_walks_graphs_and_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("covered by the gpu tests")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        plugin.get({"a": 1}, "a")


# ----------------------------------------------------------------------------- build container: real registrations
@pytest.mark.skipif(not HAVE_REF, reason="needs the reference checkout (build container only)")
def test_register_against_the_reference_modules():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _refshim

    _refshim.install()
    done = plugin.register()
    from dask_array import _chunk_types, _core_utils, _dispatch
    from dask_array_b200 import DeviceChunk, _eager

    assert _chunk_types.is_valid_chunk_type(DeviceChunk)                                  # B1
    assert _core_utils.concatenate_lookup.dispatch(DeviceChunk) is _eager.concatenate     # B1'
    assert _core_utils.tensordot_lookup.dispatch(DeviceChunk) is _eager.tensordot
    assert _dispatch.einsum_lookup.dispatch(DeviceChunk) is _eager.einsum
    assert _dispatch.divide_lookup.dispatch(DeviceChunk) is _eager.divide
    assert _dispatch.numel_lookup.dispatch(DeviceChunk) is _eager.numel
    assert _dispatch.nannumel_lookup.dispatch(DeviceChunk) is not _dispatch._nannumel
    # numpy registrations are untouched
    assert _dispatch.divide_lookup.dispatch(np.ndarray) is _dispatch._divide
    assert type(done["backend"]).__name__ == "B200BackendEntrypoint"
    assert done["backend"].default_bit_generator is np.random.PCG64
    plugin.register()                                                                     # idempotent
    assert _chunk_types._HANDLED_CHUNK_TYPES.count(DeviceChunk) == 1
    # the reference's numel on a DeviceChunk-shaped object: shape arithmetic only, host result
    fake = types.SimpleNamespace(shape=(6, 5))
    assert np.array_equal(_eager.numel(fake, axis=(0,), keepdims=True), _dispatch._numel(fake, axis=(0,), keepdims=True))


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_get_runs_reference_shaped_graph_on_device_chunks():
    """B3: a materialised mean(axis=0) graph -- leaf blocks, mean_chunk per block, mean_agg over nested lists --
    in legacy tuple form AND through GraphNode-like objects; every function runs on DeviceChunk."""
    from oracle import reference as ref
    from dask_array_b200 import DeviceChunk

    rng = np.random.default_rng(3)
    xh = rng.random((96, 64))
    seen = []

    def mean_chunk(b):
        seen.append(type(b))
        return ref.mean_chunk(b, np.dtype("f8"), (0,), True)

    def mean_agg(pairs):
        return ref.mean_agg(pairs, np.dtype("f8"), (0,), keepdims=False)

    dsk = {}
    for i in range(3):
        for j in range(2):
            dsk[("x", i, j)] = xh[32 * i:32 * i + 32, 32 * j:32 * j + 32]
            dsk[("c", i, j)] = (mean_chunk, ("x", i, j))
    for j in range(2):
        dsk[("m", j)] = (mean_agg, [("c", i, j) for i in range(3)])
    got = plugin.get(dsk, [("m", 0), ("m", 1)])
    assert all(t is DeviceChunk for t in seen) and len(seen) == 6
    np.testing.assert_allclose(np.concatenate(got), xh.mean(axis=0), rtol=1e-12)

    class Task:                                        # dask._task_spec.Task: .dependencies + __call__(values)
        def __init__(self, key, func, *deps):
            self.key, self.func, self.deps, self.dependencies = key, func, deps, frozenset(deps)

        def __call__(self, values):
            return self.func(*[values[d] for d in self.deps])

    class Data:
        dependencies = frozenset()

        def __init__(self, v):
            self.v = v

        def __call__(self, values=()):
            return self.v

    g = {("x", 0): Data(xh[:48]), ("x", 1): Data(xh[48:])}
    g[("s", 0)] = Task(("s", 0), lambda b: np.sum(b, axis=0, keepdims=True), ("x", 0))
    g[("s", 1)] = Task(("s", 1), lambda b: np.sum(b, axis=0, keepdims=True), ("x", 1))
    g["tot"] = Task("tot", lambda a, b: np.concatenate([a, b], axis=0).sum(axis=0), ("s", 0), ("s", 1))
    g["alias"] = "tot"
    out = plugin.get(g, "alias", to_host=False)
    assert isinstance(out, DeviceChunk)
    np.testing.assert_allclose(out.to_numpy(), xh.sum(axis=0), rtol=1e-12)


@pytest.mark.gpu
def test_reference_arg_path_runs_on_device_chunks():
    """B1 leftovers: arg_chunk (unravel_index / ravel_multi_index / ``arg[:] = ...``), _arg_combine (np.ogrid
    take-along-axis, ``vals.ravel()[local]``) on DeviceChunk == the same functions on NumPy, bit for bit."""
    from oracle import reference as ref
    from dask_array_b200 import DeviceChunk

    rng = np.random.default_rng(4)
    xh = np.floor(rng.random((40, 36)) * 20)           # many ties
    blocks = {(i, j): xh[20 * i:20 * i + 20, 12 * j:12 * j + 12] for i in range(2) for j in range(3)}
    dev = {k: DeviceChunk.from_numpy(np.ascontiguousarray(v)) for k, v in blocks.items()}
    for func, argfunc in ((ref.chunk_max, ref.argmax_kd), (ref.chunk_min, ref.argmin_kd)):
        # axis=1: offsets along the axis, combine over the three column blocks of a block row
        for i in range(2):
            parts_h = [ref.arg_chunk(func, argfunc, blocks[(i, j)], (1,), 12 * j) for j in range(3)]
            parts_d = [ref.arg_chunk(func, argfunc, dev[(i, j)], (1,), 12 * j) for j in range(3)]
            assert all(isinstance(p, dict) and isinstance(p["arg"], DeviceChunk) for p in parts_d)
            cat_h = ref.concatenate2(parts_h, axes=[1])
            cat_d = {k: np.concatenate([p[k] for p in parts_d], axis=1) for k in ("vals", "arg")}
            want = ref.arg_agg(argfunc, cat_h, (1,), keepdims=False)
            got = ref.arg_agg(argfunc, cat_d, (1,), keepdims=False)
            assert np.array_equal(got.to_numpy(), want)
            comb = ref.arg_combine(argfunc, cat_d, (1,))
            assert np.array_equal(comb["arg"].to_numpy(), ref.arg_combine(argfunc, cat_h, (1,))["arg"])
        # axis=None: the ravel path with per-block N-d offsets
        parts_h, parts_d = [], []
        for (i, j), b in blocks.items():
            info = ((20 * i, 12 * j), xh.shape)
            parts_h.append(ref.arg_chunk(func, argfunc, b, (0, 1), info))
            parts_d.append(ref.arg_chunk(func, argfunc, dev[(i, j)], (0, 1), info))
        for ph, pd_ in zip(parts_h, parts_d):
            assert np.array_equal(pd_["arg"].to_numpy(), ph["arg"]) and np.array_equal(pd_["vals"].to_numpy(), ph["vals"])
        cat_h = np.concatenate([p.reshape(1, 1) for p in parts_h], axis=1)
        cat_d = {k: np.concatenate([p[k] for p in parts_d], axis=1) for k in ("vals", "arg")}
        want = ref.arg_agg(argfunc, cat_h, (0, 1), keepdims=False)
        got = ref.arg_agg(argfunc, cat_d, (0, 1), keepdims=False)
        assert int(got.to_numpy()) == int(want)


@pytest.mark.gpu
def test_plugin_compute_of_a_reference_tree_matches_numpy():
    xh = np.random.default_rng(5).random((128, 96), dtype=np.float32)
    x, y, (s, m, p) = ref_chain(xh, (32, 32))
    want = np.sin(xh) * 2 + xh**2
    fused = node("FusedBlockwise", exprs=(y, m, s, p), chunks=y.chunks, dtype=np.dtype("f4"))
    np.testing.assert_allclose(plugin.compute(fused), want, rtol=3e-7)
    mean = node("Mean", array=y, axis=(0,), keepdims=False, dtype=np.dtype("f4"), split_every=None, aggregate=None)
    np.testing.assert_allclose(plugin.compute(mean), want.astype(np.float64).mean(axis=0), rtol=1e-5)
    var = node("Var", array=y, axis=None, keepdims=False, dtype=np.dtype("f4"), split_every=None,
               aggregate=functools.partial(lambda *a, **k: None, ddof=0), shape=(), ndim=0, chunks=())

    def safe_sqrt(a): ...
    std = ref_elemwise(safe_sqrt, var)
    np.testing.assert_allclose(plugin.compute(std), want.astype(np.float64).std(), rtol=1e-5)
    t = node("Transpose", array=x, axes=(1, 0), shape=(96, 128), ndim=2, chunks=(x.chunks[1], x.chunks[0]), dtype=x.dtype)
    amax = node("Max", array=t, axis=(1,), keepdims=False, dtype=np.dtype("f4"), split_every=None, aggregate=None)
    assert np.array_equal(plugin.compute(amax), xh.T.max(axis=1))


@pytest.mark.gpu
def test_chunk_level_tensordot_einsum_and_exact_gemm():
    """tensordot_lookup / einsum_lookup implementations on DeviceChunk (B1'), incl. the fp64 / integer GEMM."""
    from dask_array_b200 import DeviceChunk, _eager

    rng = np.random.default_rng(6)
    a, b = rng.random((5, 24, 7)), rng.random((7, 24, 3))
    A, B = DeviceChunk.from_numpy(a), DeviceChunk.from_numpy(b)
    got = _eager.tensordot(A, B, axes=((1, 2), (1, 0)))
    np.testing.assert_allclose(got.to_numpy(), np.tensordot(a, b, axes=((1, 2), (1, 0))), rtol=1e-12)
    assert got.dtype == np.float64
    ai, bi = rng.integers(-50, 50, (33, 70)).astype(np.int32), rng.integers(-50, 50, (70, 21)).astype(np.int32)
    gi = np.matmul(DeviceChunk.from_numpy(ai), DeviceChunk.from_numpy(bi))
    assert gi.dtype == np.int32 and np.array_equal(gi.to_numpy(), ai @ bi)
    af, bf = rng.random((130, 72), dtype=np.float32) - 0.5, rng.random((72, 260), dtype=np.float32) - 0.5
    gf = np.matmul(DeviceChunk.from_numpy(af), DeviceChunk.from_numpy(bf)).to_numpy()
    a64, b64 = af.astype(np.float64), bf.astype(np.float64)
    assert np.all(np.abs(gf - a64 @ b64) <= 1e-5 * (np.abs(a64) @ np.abs(b64)))
    for subs, ops in (("ij,jk->ik", (a[0], b[:, :3, 0].T[:3].T if False else rng.random((7, 4)))),
                      ("ijk,kjl->il", (a, b)), ("ij->ji", (a[0],)), ("ij->", (a[0],)), ("ij,ij->", (a[0], a[0])),
                      ("i,j->ij", (a[0, :, 0], a[0, 0])), ("ij,jk,kl->il", (rng.random((6, 5)), rng.random((5, 4)), rng.random((4, 3))))):
        want = np.einsum(subs, *ops)
        got = _eager.einsum(subs, *[DeviceChunk.from_numpy(np.ascontiguousarray(o)) for o in ops])
        np.testing.assert_allclose(got.to_numpy(), want, rtol=1e-12, err_msg=subs)
    with pytest.raises(NotImplementedError, match="batch"):
        _eager.einsum("bij,bjk->bik", A, DeviceChunk.from_numpy(rng.random((5, 7, 2))))


def test_get_graph_walking_on_host_objects(monkeypatch):
    """The scheduler's graph handling without a GPU (device upload patched out): legacy tuples with nested key
    lists and inline tasks, GraphNode-like values, aliases, literals, nested result keys, and a chain deep enough
    to break a recursive walker."""
    monkeypatch.setattr(plugin, "_gpu", lambda: True)
    monkeypatch.setattr(plugin, "to_device", lambda x: x)
    add = lambda a, b: a + b                                        # noqa: E731
    dsk = {"a": 1, "b": (add, "a", 10), "c": (sum, ["a", "b", (add, "b", 1)]), "alias": "c",
           ("x", 0): np.arange(4.0), ("x", 1): (np.multiply, ("x", 0), 2.0)}
    assert plugin.get(dsk, "alias") == 1 + 11 + 12
    out = plugin.get(dsk, [["a", "b"], [("x", 1)]])
    assert out[0] == [1, 11] and np.array_equal(out[1][0], np.arange(4.0) * 2)

    class Node:
        def __init__(self, fn, *deps):
            self.fn, self.deps, self.dependencies = fn, deps, frozenset(deps)

        def __call__(self, values):
            return self.fn(*[values[d] for d in self.deps])

    g = {"n0": Node(lambda: 0)}
    for i in range(1, 5000):                                        # deeper than Python's recursion limit
        g[f"n{i}"] = Node(lambda v: v + 1, f"n{i - 1}")
    assert plugin.get(g, "n4999") == 4999
    with pytest.raises(KeyError):
        plugin.get({"a": (add, "missing-is-a-literal", 1)}, "nope")


@pytest.mark.gpu
def test_plugin_compute_of_reference_contractions_and_windows():
    rng = np.random.default_rng(9)
    ah, bh = rng.random((64, 48)), rng.random((48, 32))
    got = plugin.compute(_ref_contraction("matmul", ah, bh, ((32, 32), (16,) * 3), ((16,) * 3, (32,))), optimize=False)
    np.testing.assert_allclose(got, ah @ bh, rtol=1e-12)
    got = plugin.compute(_ref_contraction("matmul", ah, bh, ((32, 32), (48,)), ((48,), (32,))), optimize=False)
    np.testing.assert_allclose(got, ah @ bh, rtol=1e-12)
    ih, jh = rng.integers(-9, 9, (32, 24)), rng.integers(-9, 9, (24, 16))
    got = plugin.compute(_ref_contraction("matmul", ih, jh, ((16, 16), (8,) * 3), ((8,) * 3, (16,))), optimize=False)
    assert got.dtype == np.int64 and np.array_equal(got, ih @ jh)
    th, uh = rng.random((8, 6, 10)), rng.random((10, 6, 4))
    got = plugin.compute(_ref_contraction("tensordot", th, uh, ((4, 4), (6,), (5, 5)), ((5, 5), (6,), (4,)),
                                          axes=((2, 1), (0, 1))), optimize=False)
    np.testing.assert_allclose(got, np.tensordot(th, uh, axes=((2, 1), (0, 1))), rtol=1e-12)
    xh = rng.random((40, 64), dtype=np.float32)
    view = np.lib.stride_tricks.sliding_window_view
    got = plugin.compute(_ref_window(xh, ((20, 20), (32, 32)), 5, 1, "max"), optimize=False)
    assert np.array_equal(got, view(xh, 5, axis=1).max(axis=-1))
    got = plugin.compute(_ref_window(xh, ((20, 20), (32, 32)), 7, 0, "sum", dtype="f4"), optimize=False)
    np.testing.assert_allclose(got, view(xh, 7, axis=0).sum(axis=-1, dtype=np.float64), rtol=1e-5)
    xn = xh.copy(); xn[3, 10:14] = np.nan
    mv = node("MovingWindowReduction", array=ref_from_array(xn, ((20, 20), (32, 32))), window=4, min_count=2,
              sliding_axis=1, reducer="move_sum", dtype=np.dtype("f4"))
    got = plugin.compute(mv, optimize=False)
    want = np.full(xn.shape, np.nan, dtype=np.float64)
    for t in range(xn.shape[1]):
        w = xn[:, max(0, t - 3):t + 1].astype(np.float64)
        ok = (~np.isnan(w)).sum(axis=1)
        want[:, t] = np.where(ok >= 2, np.nansum(w, axis=1), np.nan)
    np.testing.assert_allclose(got, want, rtol=1e-5, equal_nan=True)


@pytest.mark.gpu
def test_plugin_compute_of_reference_views_and_cumulative():
    rng = np.random.default_rng(10)
    xh, yh = rng.random((6, 8)), rng.random((6, 8))
    v = _ref_views(xh, yh, ((3, 3), (4, 4)))
    want = dict(cat=np.concatenate([xh, yh], axis=1), stk=np.stack([xh, yh]), exp=xh[None, :, None, :],
                sqz=xh[:, None, :], bto=np.broadcast_to(xh[0:1], (3, 1, 8)), sl=xh[3::2, 5])
    for k, e in v.items():
        assert np.array_equal(plugin.compute(e, optimize=False), want[k]), k
    def nancumsum(x, axis=None, dtype=None): ...
    xn = xh.copy(); xn[2, 3] = np.nan
    for func, ref_fn, src in ((np.cumsum, np.cumsum, xh), (np.cumprod, np.cumprod, xh), (nancumsum, np.nancumsum, xn)):
        for axis in (0, 1):
            for method in ("sequential", "blelloch"):
                got = plugin.compute(_ref_cum(src, ((3, 3), (4, 4)), func, axis, method), optimize=False)
                np.testing.assert_allclose(got, ref_fn(src, axis=axis), rtol=1e-12)
    ih = rng.integers(-5, 5, (6, 8)).astype("i4")
    got = plugin.compute(_ref_cum(ih, ((3, 3), (4, 4)), np.cumsum, 0, dtype="i8"), optimize=False)
    assert got.dtype == np.int64 and np.array_equal(got, np.cumsum(ih, axis=0, dtype="i8"))
