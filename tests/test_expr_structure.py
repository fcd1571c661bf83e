"""Structure of optimised expression trees (CPU): the fusion groups, tree depths and rewrite
rules the reference documents (README.md:52-57, tests/test_collection.py:996-1135,
tests/test_reductions.py:646-681), reproduced by the host-side front-end."""
import numpy as np

import dask_array_b200 as da
from dask_array_b200._blockwise import FusedBlockwise, FusedPlan
from dask_array_b200._rechunk import TasksRechunk, old_to_new
from dask_array_b200._reductions import PartialReduce, normalize_split_every
from oracle import reference as ref


def labels(expr):
    out = [expr._tree_label()]
    for d in expr.dependencies():
        out += labels(d)
    return out


def test_readme_example_tree():
    x = da.ones((1000, 1000), chunks=(100, 100))
    opt = (x + x.T)[:100, :100].optimize().expr
    # README: FusedBlockwise(add, transpose) <- Ones(100, 100); `x` is evicted from the group by
    # the a + a.T conflict rule (_blockwise.py:1342-1402) and the slice reached the leaf.
    assert isinstance(opt, FusedBlockwise)
    assert [e._tree_label() for e in opt.exprs] == ["Elemwise(add)", "Transpose(1, 0)"]
    (dep,) = opt.dependencies()
    assert dep.shape == (100, 100) and "Ones" in dep._tree_label()


def test_chunk_step_fuses_with_chain_and_tree_depth():
    x = da.random.default_rng(0).random((32768, 32768), dtype=np.float32, chunks=(4096, 4096))
    y = da.sin(x) * 2 + x**2
    m = y.mean(axis=0).optimize().expr
    assert isinstance(m, PartialReduce) and m.operand("final") and m.operand("split_every") == {0: 16}
    (f,) = m.dependencies()
    assert isinstance(f, FusedBlockwise) and len(f.exprs) == 5 and f.exprs[0]._tree_label().startswith("mean_chunk")
    s = y.std().optimize().expr                      # sqrt <- aggregate <- partial <- fused chunk step
    agg = s.dependencies()[0]
    part = agg.dependencies()[0]
    assert agg.operand("split_every") == {0: 4, 1: 4} and not part.operand("final")
    assert part.chunks == ((1, 1), (1, 1))
    plan = FusedPlan(part.dependencies()[0])
    assert len(plan.program.inputs) == 1 and "sin" in plan.program.key()


def test_partial_reduce_groups_match_oracle_nesting():
    x = da.from_array(np.zeros((50, 30)), chunks=(10, 10))
    e = x.sum().optimize().expr
    while not (isinstance(e, PartialReduce) and not e.operand("final")):
        e = e.dependencies()[0]
    got = {k: m for k, m in e.groups()}
    blocks = {bid: bid for bid in np.ndindex(5, 3)}
    want = ref.partial_reduce(ref.Blocked(blocks, ((1,) * 5, (1,) * 3)), lambda lst: lst, {0: 4, 1: 4}, True)

    def flat(v):
        return [t for i in v for t in flat(i)] if isinstance(v, list) else [v]
    assert {k: flat(v) for k, v in want.blocks.items()} == got
    assert normalize_split_every(None, (0, 1, 2)) == ref.normalize_split_every(None, (0, 1, 2))


def test_transpose_conflict_and_same_mapping():
    a = da.from_array(np.zeros((8, 8)), chunks=4)
    opt = (a + a.T).optimize().expr                   # leaf read through two mappings -> two kernel inputs
    plan = FusedPlan(opt)
    assert len(plan.leaves) == 2 and {m for _, m in plan.leaves} == {(0, 1), (1, 0)}
    plan2 = FusedPlan((a + a * 2).optimize().expr)    # same mapping twice -> one input
    assert len(plan2.leaves) == 1
    assert (a.T.T).optimize().expr._name == a.optimize().expr._name


def _opaque(shape, chunks, token):
    """A leaf that cannot absorb a rechunk (per-block host callbacks), like the reference's opaque IO."""
    return da.from_host_blocks(lambda bid: None, shape, chunks, np.float64, token=token)


def test_rechunk_rules_and_pieces():
    x = _opaque((16, 16), (16, 2), "rr")
    assert x.rechunk((16, 2)).optimize().expr._name == x.optimize().expr._name      # no-op removed
    r = x.rechunk((4, 8)).rechunk((2, 16)).optimize().expr                            # double rechunk collapses
    assert isinstance(r, TasksRechunk) and r.chunks == ((2,) * 8, (16,))
    assert not isinstance(r.operand("array"), TasksRechunk)
    assert len(r.pieces((0, 0))) == 8
    o2n = old_to_new(((4, 4, 3), (2, 2, 2)), ((2, 6, 3), (6,)))
    assert o2n[0][1] == [(0, slice(2, 4)), (1, slice(0, 4))]


def _walk(e, acc=None):
    acc = [] if acc is None else acc
    acc.append(e)
    for d in e.dependencies():
        _walk(d, acc)
    return acc


def test_rechunk_pushdown_mirrors_reference_rules():
    """dask_array/tests/test_rechunk_pushdown.py: :132 (into FromArray), :168 (through elemwise),
    :487-528 (through transpose), :682 (lower-inserted rechunks), :605-680 (never into shared nodes)."""
    from dask_array_b200._expr import FromArray
    from dask_array_b200._blockwise import Elemwise
    from dask_array_b200._rechunk import Rechunk, pushdown_rechunks
    data = np.arange(100 * 50, dtype=np.float64).reshape(100, 50)
    darr = da.from_array(data, chunks=(25, 25))
    s = pushdown_rechunks(darr.rechunk((50, 50)).expr.simplify())
    assert isinstance(s, FromArray) and s.chunks == ((50, 50), (50,))
    s = pushdown_rechunks(darr.rechunk({0: 50}).expr.simplify())
    assert isinstance(s, FromArray) and s.chunks == ((50, 50), (25, 25))
    # through elemwise (incl. a broadcast operand: its extent-1 / missing dims keep their chunks)
    x, v = da.from_array(data, chunks=(25, 25)), da.from_array(data[0], chunks=10)
    s = pushdown_rechunks(((x + 1) * v).rechunk((20, 5)).expr.simplify())
    assert isinstance(s, Elemwise) and not any(type(n) is Rechunk for n in _walk(s))
    assert {n.chunks for n in _walk(s) if isinstance(n, FromArray)} == {((20,) * 5, (5,) * 10), ((5,) * 10,)}
    # through transpose: input axis i takes the chunks of the output axis it lands on
    t = _opaque((2, 3, 4), (1, 1, 2), "tp")
    got = pushdown_rechunks(t.transpose((2, 0, 1)).rechunk((2, 1, 3)).expr.simplify())
    want = t.rechunk((1, 3, 2)).transpose((2, 0, 1)).expr.simplify()
    assert got._name == want._name
    # rechunks inserted by chunk unification at lowering are absorbed by the reads too
    a, b = da.from_array(data[:22, :22], chunks=(11, 4)), da.from_array(data[:22, :22], chunks=(4, 11))
    assert not any("Rechunk" in type(n).__name__ for n in _walk((a + b).optimize().expr))
    # shared nodes are left alone: one read, the chain is not duplicated
    y = (x + 1) * 2
    z = y.sum() + y.rechunk((50, 50)).sum()
    opt = z.optimize(fuse=False).expr
    assert len({n._name for n in _walk(opt) if isinstance(n, FromArray)}) == 1
    assert len({n._name for n in _walk(opt) if isinstance(n, Elemwise)}) == 3      # add, mul, top-level add
    z2 = x.sum() + x.rechunk((50, 50)).sum()
    opt2 = z2.optimize(fuse=False).expr
    assert len({n._name for n in _walk(opt2) if isinstance(n, FromArray)}) == 1
    assert any("Rechunk" in type(n).__name__ for n in _walk(opt2))


def test_elemwise_dtypes_follow_numpy_nep50():
    i4 = da.from_array(np.zeros((4,), np.int32), chunks=2)
    f4 = da.from_array(np.zeros((4,), np.float32), chunks=2)
    assert (f4 * 2).dtype == np.float32 and (i4 / i4).dtype == np.float64 and (i4**2).dtype == np.int32
    assert (i4 < 3).dtype == np.bool_ and (f4 + i4).dtype == np.float64 and (f4 * 2.5).dtype == np.float32
    assert i4.sum().dtype == np.int64 and i4.mean().dtype == np.float64 and f4.var().dtype == np.float32
    assert i4.argmax().dtype == np.int64


def test_unaligned_chunks_get_rechunked():
    a = _opaque((12, 12), (4, 12), "ua")
    b = _opaque((12, 12), (6, 12), "ub")
    assert any("Rechunk" in l for l in labels((a + b).optimize().expr))


def test_diamond_group_is_rebuilt_consistently():
    """x + x*2 with x = abs(T(rechunk(...))): the fused group is a DAG; after substitution no
    un-fused Elemwise/Transpose may remain anywhere in the tree (found by the fuzz test)."""
    from dask_array_b200._blockwise import Elemwise, Transpose
    a = da.from_array(np.zeros((81, 51)), chunks=(20, 17))
    x = abs((abs(a) * 3).rechunk((48, 15)).T)
    opt = (x + x * 2).optimize().expr

    def bare(e):
        return int(isinstance(e, (Elemwise, Transpose))) + sum(bare(d) for d in e.dependencies())
    assert bare(opt) == 0
    assert len(FusedPlan(opt).leaves) == 1


def test_fused_groups_respect_the_kernel_input_limit():
    """A kernel's descriptor has B2_MAX_IN input slots: wider expressions are cut into several groups,
    each swallowing as many inputs as fit (found by the fuzz test)."""
    from dask_array_b200 import _lib
    from dask_array_b200._blockwise import FusedBlockwise
    arrs = [_opaque((40, 40), (10, 10), f"lim{i}") for i in range(11)]
    e = arrs[0]
    for a in arrs[1:]:
        e = e * 2 + a
    fused = {n._name: n for n in _walk(e.optimize().expr) if isinstance(n, FusedBlockwise)}
    sizes = sorted(len(FusedPlan(f).leaves) for f in fused.values())
    assert sizes == [6, 6] and max(sizes) <= _lib.B2_MAX_IN
    t = ((arrs[0] + arrs[1]) + (arrs[2] + arrs[3])) + ((arrs[4] + arrs[5]) + (arrs[6] + arrs[7].T))
    fused = {n._name: n for n in _walk(t.optimize().expr) if isinstance(n, FusedBlockwise)}
    assert all(len(FusedPlan(f).leaves) <= _lib.B2_MAX_IN for f in fused.values()) and len(fused) == 2
