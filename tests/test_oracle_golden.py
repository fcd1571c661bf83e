"""The oracle (oracle/reference.py) against golden outputs of the reference's OWN functions
(tests/golden/hotpath.*, produced by tests/golden/generate.py from /root/reference).
Bit-exact: both sides are NumPy on the same inputs in the same order."""
import json
import os

import ast

import numpy as np
import pytest

from oracle import reference as ref

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "hotpath.npz"))
with open(os.path.join(HERE, "golden", "hotpath.json")) as f:
    J = json.load(f)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b, equal_nan=True)


def blocks_of(x, chunks):
    return ref.Blocked.from_array(x, chunks).blocks


def nested(parts, axis):
    fixed = {d: ([0, 1] if d in axis else d_) for d, d_ in ((0, 0), (1, 1))}

    def rec(d, prefix):
        if d == 2:
            return parts[prefix]
        if d in axis:
            return [rec(d + 1, prefix + (i,)) for i in fixed[d]]
        return rec(d + 1, prefix + (fixed[d],))
    return rec(0, ())


@pytest.mark.parametrize("tag,acc", [("f4", "f4"), ("i4", "f8"), ("f8", "f8")])
@pytest.mark.parametrize("axis", [(0,), (1,), (0, 1)])
def test_mean_kernels(tag, acc, axis):
    x = G[f"mean.{tag}.x"]
    at = "".join(map(str, axis))
    bl = blocks_of(x, ((5, 7), (4, 6)))
    parts = {bid: ref.mean_chunk(b, np.dtype(acc), axis, True) for bid, b in bl.items()}
    for bid, p in parts.items():
        for k in ("n", "total"):
            same(p[k], G[f"mean.{tag}.ax{at}.chunk{bid[0]}{bid[1]}.{k}"])
    fixed_other = [d for d in (0, 1) if d not in axis]
    parts_fixed = {bid: p for bid, p in parts.items()}
    if fixed_other:          # the golden fixes the non-reduced block index at its own position id
        d = fixed_other[0]
        parts_fixed = {bid: p for bid, p in parts.items() if bid[d] == d}
    lst = nested(parts_fixed, axis)
    comb = ref.mean_combine(lst, np.dtype(acc), axis, True)
    for k in ("n", "total"):
        same(comb[k], G[f"mean.{tag}.ax{at}.combine.{k}"])
    same(ref.mean_agg(lst, np.dtype(acc), axis, False), G[f"mean.{tag}.ax{at}.agg"])


@pytest.mark.parametrize("tag,acc", [("f4", "f4"), ("f8", "f8"), ("i4", "f8")])
@pytest.mark.parametrize("axis", [(0,), (1,), (0, 1)])
def test_moment_kernels(tag, acc, axis):
    x = G[f"var.{tag}.x"]
    at = "".join(map(str, axis))
    bl = blocks_of(x, ((5, 7), (4, 6)))
    parts = {bid: ref.moment_chunk(b, np.dtype(acc), axis, True) for bid, b in bl.items()}
    for bid, p in parts.items():
        for k in ("n", "total", "M"):
            same(p[k], G[f"var.{tag}.ax{at}.chunk{bid[0]}{bid[1]}.{k}"])
    other = [d for d in (0, 1) if d not in axis]
    if other:
        parts = {bid: p for bid, p in parts.items() if bid[other[0]] == other[0]}
    lst = nested(parts, axis)
    comb = ref.moment_combine(lst, np.dtype(acc), axis)
    for k in ("n", "total", "M"):
        same(comb[k], G[f"var.{tag}.ax{at}.combine.{k}"])
    for ddof in (0, 1):
        same(ref.moment_agg(lst, np.dtype(acc), axis, False, ddof=ddof), G[f"var.{tag}.ax{at}.agg.ddof{ddof}"])


@pytest.mark.parametrize("nm", ["max", "min"])
@pytest.mark.parametrize("axis", [(0,), (1,), (0, 1)])
def test_arg_kernels(nm, axis):
    x = G["arg.x"]
    at = "".join(map(str, axis))
    chunks = ((4, 5), (6, 8))
    func, argfunc = (np.max, ref.argmax_kd) if nm == "max" else (np.min, ref.argmin_kd)
    starts = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in chunks]
    parts = {}
    for bid, b in blocks_of(x, chunks).items():
        off = tuple(int(starts[d][i]) for d, i in enumerate(bid))
        info = (off, x.shape) if len(axis) == 2 else off[axis[0]]
        parts[bid] = ref.arg_chunk(func, argfunc, b, axis, info)
        same(parts[bid]["vals"], G[f"arg.{nm}.ax{at}.chunk{bid[0]}{bid[1]}.vals"])
        same(parts[bid]["arg"], G[f"arg.{nm}.ax{at}.chunk{bid[0]}{bid[1]}.arg"])
    other = [d for d in (0, 1) if d not in axis]
    if other:
        parts = {bid: p for bid, p in parts.items() if bid[other[0]] == other[0]}
    data = ref.concatenate2(nested(parts, axis), axes=sorted(axis))
    comb = ref.arg_combine(argfunc, data, axis)
    same(comb["vals"], G[f"arg.{nm}.ax{at}.combine.vals"])
    same(comb["arg"], G[f"arg.{nm}.ax{at}.combine.arg"])
    same(ref.arg_agg(argfunc, data, axis, False), G[f"arg.{nm}.ax{at}.agg"])


def test_small_kernels():
    y = G["minmax.x"]
    same(ref.chunk_min(y, axis=(1,), keepdims=True), G["minmax.min1"])
    same(ref.chunk_max(y, axis=(0,), keepdims=True), G["minmax.max0"])
    same(np.array(ref.numel(y, axis=(0,), keepdims=True, dtype="f4")), G["numel.ax0"])
    same(np.array(ref.numel(y, axis=(0, 1), keepdims=True, dtype="f8")), G["numel.all"])
    a, b = G["cat2.a"], G["cat2.b"]
    same(ref.concatenate2([a, b], axes=[0]), G["cat2.ax0"])
    same(ref.concatenate2([[a, b], [b, a]], axes=[0, 1]), G["cat2.ax01"])
    blk = ref.Blocked({(0, 0): a, (0, 1): b, (1, 0): b, (1, 1): a}, ((2, 2), (3, 3)))
    same(blk.to_array(), G["cat3"])
    same(np.matmul(G["matmul.a"], G["matmul.b"])[..., np.newaxis, :], G["matmul.out"])
    bt = ref.broadcast_trick(1, (3, 4), (3, 4), np.float64).blocks[(0, 0)]
    assert list(bt.shape) == J["broadcast_trick"]["shape"]
    assert list(bt.strides) == J["broadcast_trick"]["strides"] == [0, 0]
    assert float(bt[0, 0]) == J["broadcast_trick"]["value"]


def test_split_every_and_tree_nesting():
    def parse(k):
        se, ax = k.split(",", 1)
        se = None if se == "None" else eval(se)
        return se, eval(ax)
    for k, want in J["split_every"].items():
        se, ax = parse(k)
        got = ref.normalize_split_every(se, ax)
        assert {str(a): n for a, n in got.items()} == want, k
    for tag, L in J["partial_reduce_layers"].items():
        nb = tuple(L["numblocks"])
        split = {int(k): v for k, v in L["split_every"].items()}
        blocks = {bid: ("x",) + bid for bid in np.ndindex(*nb)}
        x = ref.Blocked(blocks, tuple((1,) * n for n in nb))

        def as_json(v):
            return [as_json(i) for i in v] if isinstance(v, list) else list(v)
        res = ref.partial_reduce(x, as_json, split, L["keepdims"])
        got = sorted([list(k), v] for k, v in res.blocks.items())
        assert got == sorted(L["tasks"]), tag


def test_rechunk_intersections_and_plan():
    for tag, item in J["intersect_chunks"].items():
        old = tuple(tuple(c) for c in item["old"])
        new = tuple(tuple(c) for c in item["new"])
        o2n = ref.old_to_new(old, new)
        import itertools
        got = []
        for nbid in itertools.product(*[range(len(c)) for c in new]):
            per_dim = [o2n[d][i] for d, i in enumerate(nbid)]
            pieces = []
            for combo in itertools.product(*per_dim):
                pieces.append([[i, s.start, s.stop] for (i, s) in combo])
            got.append(pieces)
        assert got == item["pieces"], tag
    # the planner's stages for config 4 (documented in DESIGN.md; values are stage independent)
    assert J["plan_rechunk"]["c4_f8"][0][1] == [1024] * 16
    assert J["plan_rechunk"]["c4_f4"][0][1] == [2048] * 8
    assert J["getitem_small_is_copy"] is True


def test_oracle_end_to_end_matches_numpy():
    rng = np.random.default_rng(7)
    x = rng.random((50, 37)).astype(np.float32)
    b = ref.Blocked.from_array(x, (16, 10))
    np.testing.assert_allclose(ref.da_mean(b, axis=0), x.mean(axis=0), rtol=1e-6)
    np.testing.assert_allclose(ref.da_std(b), x.std(), rtol=1e-5)
    np.testing.assert_allclose(ref.da_var(b, axis=1, ddof=1), x.var(axis=1, ddof=1), rtol=1e-5)
    assert np.array_equal(ref.da_argmax(b, axis=1), x.argmax(axis=1))
    assert ref.da_argmin(b) == x.argmin()
    assert np.array_equal(ref.da_min(b, axis=0), x.min(axis=0))
    assert np.array_equal(ref.da_sum(ref.Blocked.from_array((x * 100).astype(np.int32), (16, 10))),
                          (x * 100).astype(np.int32).sum())
    r = ref.rechunk(b, ((50,), (5,) * 7 + (2,)))
    assert np.array_equal(r.to_array(), x) and r.blocks[(0, 7)].shape == (50, 2)
    t = ref.elemwise(np.add, ref.transpose(ref.Blocked.from_array(x[:32, :32], (16, 16))),
                     ref.Blocked.from_array(x[:32, :32], (16, 16)))
    assert np.array_equal(t.to_array(), x[:32, :32].T + x[:32, :32])
    a = rng.random((24, 20)).astype(np.float32)
    c = rng.random((20, 12)).astype(np.float32)
    m = ref.matmul(ref.Blocked.from_array(a, (8, 5)), ref.Blocked.from_array(c, (5, 6)))
    np.testing.assert_allclose(m.to_array(), a @ c, rtol=1e-5)
    # README example
    ones = ref.broadcast_trick(1, (1000, 1000), (100, 100), np.float64)
    y = ref.elemwise(np.add, ones, ref.transpose(ones))
    assert ref.da_sum(y) == 2_000_000.0
    assert np.array_equal(ref.getitem_slices(y, (slice(0, 100), slice(0, 100))).to_array(), np.full((100, 100), 2.0))


# ----------------------------------------------------------------------------- chunk unification
def _unify_golden():
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "unify.json")) as f:
        return json.load(f)


def test_blockdim_helpers_match_reference():
    g = _unify_golden()
    for sets, want in g["coarse_blockdim"]:
        assert list(ref.coarse_blockdim([tuple(s) for s in sets])) == want, sets
    for sets, want in g["common_blockdim"]:
        assert list(ref.common_blockdim([tuple(s) for s in sets])) == want, sets
    for src, dst, want in g["moved_fraction"]:
        assert ref.moved_fraction(tuple(src), tuple(dst)) == want, (src, dst)


@pytest.mark.parametrize("case", sorted(_unify_golden()["unify"]))
def test_unify_chunks_matches_reference(case):
    rec = _unify_golden()["unify"][case]
    operands = [(tuple(s), tuple(map(tuple, ch)), np.dtype(d).itemsize) for s, ch, d in rec["operands"]]
    chunkss, targets, changed = ref.unify_chunks(operands)
    assert {str(k): list(v) for k, v in chunkss.items()} == rec["chunkss"]
    assert [[list(c) for c in t] for t in targets] == rec["result_chunks"]
    assert changed == rec["changed"]


# ----------------------------------------------------------------------------- cumulative scans
def _cum_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "cumulative.npz"))


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in _cum_golden().files}))
def test_cumulative_matches_reference_graph(case):
    g = _cum_golden()
    chunks, axis, kind, nan = ast.literal_eval(str(g[case + "/meta"][0]))
    xh, want = g[case + "/x"], g[case + "/result"]
    got = ref.da_cumulative(ref.Blocked.from_array(xh, chunks), axis, kind, nan).to_array()
    assert got.dtype == want.dtype
    assert np.array_equal(got, want, equal_nan=True)


# ----------------------------------------------------------------------------- overlap
def test_overlap_known_answer_from_the_reference_docstring():
    """`_overlap.py:935-962`: the documented output of overlap(depth={0: 2, 1: 1}, boundary={0: 100, 1: 'reflect'})."""
    x = np.arange(64).reshape((8, 8))
    got = ref.overlap(ref.Blocked.from_array(x, (4, 4)), {0: 2, 1: 1}, {0: 100, 1: "reflect"})
    assert got.chunks == ((8, 8), (6, 6))
    cols = (0, 0, 1, 2, 3, 4, 3, 4, 5, 6, 7, 7)
    rows = [[100] * 12] * 2 + [[r * 8 + c for c in cols] for r in (0, 1, 2, 3, 4, 5, 2, 3, 4, 5, 6, 7)] + [[100] * 12] * 2
    assert np.array_equal(got.to_array(), np.array(rows))


def test_overlap_host_helpers_match_reference_doctests():
    import dask_array_b200 as da
    assert da.overlap.ensure_minimum_chunksize(10, (20, 20, 1)) == (20, 11, 10)      # _overlap.py:849-852
    assert da.overlap.ensure_minimum_chunksize(3, (1, 1, 3)) == (5,)
    x = da.from_array(np.zeros((8, 8)), chunks=(4, 4))
    assert da.overlap.overlap(x, depth={0: 2, 1: 1}, boundary={0: 100, 1: "reflect"}).chunks == ((8, 8), (6, 6))
    assert da.overlap.overlap(x, depth=1, boundary="none").chunks == ((5, 5), (5, 5))
    assert da.overlap.trim_internal(da.overlap.overlap(x, depth=1, boundary="none"), {0: 1, 1: 1}).chunks == x.chunks


# ----------------------------------------------------------------------------- topk
def _topk_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "topk.npz"))


@pytest.mark.parametrize("case", sorted({k.split("/")[0] for k in _topk_golden().files}))
def test_topk_matches_reference_chunk_functions(case):
    g = _topk_golden()
    chunks, k, axis = ast.literal_eval(str(g[case + "/meta"][0]))
    xh = g[case + "/x"]
    got = ref.da_topk(ref.Blocked.from_array(xh, chunks), k, axis).to_array()
    assert np.array_equal(got, g[case + "/topk"], equal_nan=True)
    if case + "/argtopk" in g.files:
        gi = ref.da_argtopk(ref.Blocked.from_array(xh, chunks), k, axis).to_array()
        assert np.array_equal(gi, g[case + "/argtopk"])
