"""``normalize_chunks`` (dask_array/_core_utils.py:731-885) -- the chunk conventions every creation function
goes through.  Fixed cases are the reference's own doctest examples plus outputs recorded from the reference here;
the randomised equivalence runs against the reference's function itself (build container only, through
tests/golden/_refshim.py)."""
import os
import random
import sys

import numpy as np
import pytest

from dask_array_b200._expr import normalize_chunks

HAVE_REF = os.path.exists("/root/reference/dask_array/_core_utils.py")

CASES = [
    # (chunks, shape, dtype, expected) -- the doctest examples of the reference (:744-823)
    ((2, 2), (5, 6), None, ((2, 2, 1), (2, 2, 2))),
    (((2, 2, 1), (2, 2, 2)), (5, 6), None, ((2, 2, 1), (2, 2, 2))),
    ([[2, 2], [3, 3]], (4, 6), None, ((2, 2), (3, 3))),
    (10, (30, 5), None, ((10, 10, 10), (5,))),
    ((-1,), (10,), None, ((10,),)),
    ((3, None), (9, 9), None, ((3, 3, 3), (9,))),
    (("auto",), (20,), "uint8", None),             # with limit=5 below
    ("auto", (2, 3), "int32", ((2,), (3,))),
    ("1kiB", (2000,), "float32", ((256, 256, 256, 256, 256, 256, 256, 208),)),
    ((), (), None, ()),
    ((1,), (), None, ()),
    ((), (0, 0), None, ((0,), (0,))),
    # recorded from the reference in the build container
    ("auto", (32768, 32768), "float32", ((5792,) * 5 + (3808,),) * 2),
    (("auto", 4096), (32768, 32768), "float32", ((8192,) * 4, (4096,) * 8)),
    ({0: 10}, (25, 7), "float32", ((10, 10, 5), (7,))),
    ((5, 0, 5), (10,), None, ((5, 0, 5),)),
    (4, (10, 0, 5), None, ((4, 4, 2), (0,), (4, 1))),
    (np.array([2, 3]), (10, 10), None, ((2,) * 5, (3, 3, 3, 1))),
    ("auto", (0, 5), "float64", ((0,), (5,))),
]


@pytest.mark.parametrize("chunks,shape,dtype,want", CASES)
def test_fixed_cases(chunks, shape, dtype, want):
    if want is None:
        assert normalize_chunks(chunks, shape, dtype=dtype, limit=5) == ((5, 5, 5, 5),)
    else:
        assert normalize_chunks(chunks, shape, dtype=dtype) == want


def _unrle(runs):
    return tuple(v for v, n in runs for _ in range(n))


def _golden():
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "chunks.json")) as f:
        return json.load(f)


def _outcome(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception as e:      # noqa: BLE001 -- the error TYPE is part of the recorded behaviour
        return type(e).__name__


def test_golden_cases_recorded_from_the_reference():
    """tests/golden/chunks.json (generate_chunks.py): 1100 seeded requests answered by the reference's own
    ``normalize_chunks`` (with and without ``previous_chunks``) and ``_balance_chunksizes``."""
    from dask_array_b200._rechunk import balance_chunksizes

    g = _golden()
    assert len(g["normalize"]) == 400 and len(g["previous"]) == 400 and len(g["balance"]) == 300
    for case in g["normalize"] + g["previous"]:
        req = tuple(case["chunks"]) if isinstance(case["chunks"], list) else case["chunks"]
        prev = tuple(_unrle(p) for p in case["previous"]) if "previous" in case else None
        want = case["out"] if isinstance(case["out"], str) else tuple(_unrle(c) for c in case["out"])
        got = _outcome(normalize_chunks, req, tuple(case["shape"]), dtype=case["dtype"], limit=case["limit"],
                       previous_chunks=prev)
        assert got == want, case
    for case in g["balance"]:
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert balance_chunksizes(_unrle(case["chunks"])) == _unrle(case["out"]), case


def test_errors():
    with pytest.raises(ValueError, match="same length"):
        normalize_chunks((2, 2, 2), (4, 4))
    with pytest.raises(ValueError, match="add up"):
        normalize_chunks(((2, 3), (4,)), (4, 4))
    with pytest.raises(ValueError, match="Empty tuples"):
        normalize_chunks(((), (4,)), (0, 4))
    with pytest.raises(ValueError, match="String values"):
        normalize_chunks((5, "auto"), (10,))
    with pytest.raises(ValueError, match="byte unit"):
        normalize_chunks("12", (10,), dtype="f4")
    with pytest.raises(ValueError, match="consistent"):
        normalize_chunks(("1MiB", "2MiB"), (4000, 4000), dtype="f4")
    with pytest.raises(TypeError, match="dtype must be known"):
        normalize_chunks("auto", (10,))
    with pytest.raises(ValueError, match="chunks="):
        normalize_chunks(None, (10,))


def test_creation_defaults_follow_the_reference():
    """chunks="auto" is the default of from_array / ones / random (io/_from_array.py:134, creation/_ones_zeros.py,
    random/_expr.py:86-90): 128 MiB blocks, so the block structure -- and with it the per-block random streams --
    is the reference's."""
    import dask_array_b200 as da

    assert da.ones((40000, 40000)).chunks[0] == (4096,) * 9 + (3136,)
    assert da.random.default_rng(0).random((32768, 32768), dtype=np.float32).numblocks == (6, 6)
    assert da.from_array(np.zeros((100, 100))).chunks == ((100,), (100,))
    assert da.from_array(np.int_(3), chunks=(1,)).chunks == ()
    x = da.ones((10, 10), chunks=5)
    assert x.rechunk({0: -1}).chunks == ((10,), (5, 5)) and x.rechunk((2, -1)).chunks == ((2,) * 5, (10,))
    assert x.rechunk({-1: 2, 0: None}).chunks == ((5, 5), (2,) * 5) and x.rechunk((None, 10)).chunks == ((5, 5), (10,))
    with pytest.raises(ValueError, match="out of bounds"):
        x.rechunk({2: 1})


def test_rechunk_auto_scales_the_current_blocks():
    """``rechunk("auto")`` (``Rechunk.chunks`` _rechunk.py:690-718 -> ``auto_chunks(previous_chunks=x.chunks)``):
    expected values recorded from the reference's ``normalize_chunks`` in the build container."""
    import dask_array_b200 as da

    x = da.ones((32768, 32768), dtype="f4", chunks=(4096, 256))
    # growing keeps the aspect ratio of the median block (16 : 1) and merges existing blocks: 20480 x 1280 = 100 MiB
    assert x.rechunk().chunks == ((20480, 12288), (1280,) * 25 + (768,))
    assert x.rechunk(("auto", -1)).chunks == ((1024,) * 32, (32768,))
    assert x.rechunk("1MiB").chunks == ((2048,) * 16, (128,) * 256)       # shrinking: regular edges, same aspect ratio
    assert x.rechunk({0: "auto"}, block_size_limit=2**20).chunks == ((1024,) * 32, (256,) * 128)
    y = da.ones((1000,), chunks=((300, 300, 300, 100),))
    assert y.rechunk(250, balance=True).chunks == ((250,) * 4,)
    assert y.rechunk(400, balance=True).chunks == ((500, 500),)


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container)")
def test_auto_from_previous_and_balance_match_the_reference():
    import warnings

    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _refshim

    _refshim.install()
    from dask_array._core_utils import normalize_chunks as ref
    from dask_array._rechunk import _balance_chunksizes as ref_balance
    from dask_array_b200._rechunk import balance_chunksizes

    rng = random.Random(7)

    def blocks(n):
        if rng.random() < 0.5:
            c = min(n, rng.choice([1, 2, 5, 10, 64, 100, 1000, 4096, n]))
            full, rest = divmod(n, c)
            return (c,) * full + ((rest,) if rest else ())
        parts, rem = [], n
        while rem > 0:
            t = rng.randint(1, max(1, min(rem, rng.choice([3, 50, 2000, 100000]))))
            parts.append(t)
            rem -= t
        return tuple(parts)

    for _ in range(1500):
        nd = rng.randint(1, 4)
        shape = tuple(rng.choice([1, 7, 100, 5000, 40000, 123457]) for _ in range(nd))
        prev = tuple(blocks(n) for n in shape)
        chunks = tuple(rng.choice(["auto", "auto", "auto", -1, None, 3, 64, "2MiB"]) for _ in range(nd))
        chunks = tuple(prev[i] if c is None else c for i, c in enumerate(chunks))
        if rng.random() < 0.2:
            chunks = rng.choice(["auto", "64MiB", "10kiB"])
        dt = rng.choice(["f4", "f8", "i1", "c16"])
        limit = rng.choice([None, None, None, "1MiB", 5000, 10**9, 100])

        def call(fn, dtype):
            try:
                return fn(chunks, shape, limit=limit, dtype=dtype, previous_chunks=prev)
            except Exception as e:          # noqa: BLE001
                return type(e).__name__
        assert call(ref, np.dtype(dt)) == call(normalize_chunks, dt), (chunks, shape, prev, dt, limit)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(1500):
            ch = blocks(rng.choice([5, 17, 100, 1000, 4097])) if rng.random() < 0.6 else \
                tuple(rng.randint(1, rng.choice([3, 20, 500])) for _ in range(rng.randint(1, 12)))
            assert tuple(int(v) for v in ref_balance(ch)) == balance_chunksizes(ch), ch


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container)")
def test_randomised_equivalence_with_the_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _refshim

    _refshim.install()
    from dask_array._core_utils import normalize_chunks as ref

    rng = random.Random(1)
    for _ in range(2000):
        nd = rng.randint(0, 4)
        shape = tuple(rng.choice([0, 1, 7, 100, 5000, 40000, 123457]) for _ in range(nd))
        chunks = tuple(rng.choice(["auto", "auto", -1, None, 1, 3, 64, 1000, "2MiB", (1,)]) for _ in range(nd))
        chunks = tuple((shape[i],) if c == (1,) else c for i, c in enumerate(chunks))
        if rng.random() < 0.2:
            chunks = rng.choice(["auto", 5, "64MiB", -1, {0: 3}])
        dt = rng.choice(["f4", "f8", "i1", "i8", "c16"])
        limit = rng.choice([None, None, "1MiB", 5000, 10**9])

        def call(fn, dtype):
            try:
                return fn(chunks, shape, limit=limit, dtype=dtype)
            except Exception as e:          # noqa: BLE001 -- the error TYPE is part of the behaviour compared
                return type(e).__name__
        assert call(ref, np.dtype(dt)) == call(normalize_chunks, dt), (chunks, shape, dt, limit)


def _host_values(a):
    e = a.expr
    get = e.operand("get_block")
    return np.concatenate([get((k,)) for k in range(len(e.chunks[0]))])


@pytest.mark.parametrize("args,kw", [((10,), {}), ((2, 20, 3), {"chunks": 4}), ((0, 1, 0.1), {"chunks": 3}),
                                      ((77, 130, 1), {"chunks": 5}), ((10, 0, -2), {"chunks": 2}),
                                      ((5,), {"dtype": "f4", "chunks": 2}), ((0,), {}), ((1, 31.3, 2.5), {"chunks": 7}),
                                      ((2**63 - 10000, 2**63 - 1, 100), {"chunks": 30})])
def test_arange_blocks(args, kw):
    """creation/_arange.py:102-123 (tests/test_creation.py arange cases): every block is generated from its own
    start / stop; the concatenation is NumPy's arange."""
    import dask_array_b200 as da

    a = da.arange(*args, **kw)
    want = np.arange(*args, dtype=kw.get("dtype"))
    assert a.dtype == want.dtype and a.shape == want.shape
    assert sum(a.chunks[0]) == want.size and (want.size == 0 or max(a.chunks[0]) <= kw.get("chunks", want.size))
    np.testing.assert_allclose(_host_values(a), want, rtol=1e-15)
    assert da.arange(*args, **kw).name == a.name                        # deterministic names


@pytest.mark.parametrize("args,kw", [((6, 49), {"chunks": 5, "num": 13}), ((1.4, 4.9), {"chunks": 5, "num": 13}),
                                      ((0, 1), {"num": 50, "endpoint": False, "chunks": 7}),
                                      ((6, 49), {"chunks": 5, "num": 13, "dtype": int}), ((0, 1), {"num": 1}), ((0, 1), {"num": 0})])
def test_linspace_blocks(args, kw):
    import dask_array_b200 as da

    a = da.linspace(*args, **kw)
    want = np.linspace(*args, **{k: v for k, v in kw.items() if k != "chunks"})
    assert a.dtype == want.dtype and a.shape == want.shape
    np.testing.assert_allclose(_host_values(a), want, rtol=1e-13)
    _, step = da.linspace(*args, retstep=True, **kw)
    if want.size > 1:
        assert step == pytest.approx(np.linspace(*args, retstep=True, **{k: v for k, v in kw.items() if k != "chunks"})[1])
