"""``reshape`` (manipulation/_reshape.py): the chunk planner against golden cases recorded from the reference's own
``reshape_rechunk`` (tests/golden/reshape.json, generate_reshape.py) and -- in the build container -- against that
function itself on random cases; the block pairing (input block k viewed with output block k's shape) against
``np.reshape`` on the host, using the expression's own ``source`` / ``block_shape``."""
import itertools
import json
import os
import random
import sys

import numpy as np
import pytest

import dask_array_b200 as da
from dask_array_b200._reshape import reshape_plan

HAVE_REF = os.path.exists("/root/reference/dask_array/manipulation/_reshape.py")


def _tt(x):
    return tuple(tuple(c) for c in x)


def test_golden_plans():
    with open(os.path.join(os.path.dirname(__file__), "golden", "reshape.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 300
    for c in cases:
        args = (tuple(c["inshape"]), tuple(c["outshape"]), _tt(c["inchunks"]))
        if c["out"] == "NotImplementedError":
            with pytest.raises(NotImplementedError):
                reshape_plan(*args)
        else:
            assert reshape_plan(*args) == (_tt(c["out"][0]), _tt(c["out"][1])), c


def test_reference_docstring_cases():
    # manipulation/_reshape.py:23-31
    x = da.ones((6, 5, 4), chunks=(3, 2, 2))
    assert x.reshape((3, 2, 5, 4)).shape == (3, 2, 5, 4)
    assert x.reshape((30, 4)).chunks == ((5,) * 6, (2, 2))
    with pytest.raises(NotImplementedError, match="unevenly"):
        x.reshape((4, 5, 6))
    with pytest.raises(ValueError, match="unchanged"):
        x.reshape((7, 5, 4))
    with pytest.raises(ValueError, match="one unknown"):
        x.reshape((-1, -1, 4))
    assert x.reshape(-1, 4).shape == (30, 4) and x.reshape((6, 5, 4)) is x
    assert da.reshape(x, 120).shape == (120,)
    one = da.ones((6, 4), chunks=(6, 4)).reshape((2, 3, 4))
    assert type(one.expr).__name__ == "Reshape" and one.chunks == ((2,), (3,), (4,))       # one block: viewed directly
    assert da.ones((4, 4, 4), chunks=(2, 2, 2)).reshape((16, 4), merge_chunks=False).chunks == ((2,) * 8, (2, 2))


def _host_reshape(arr):
    """Evaluate ``arr`` (an Array whose expression is Reshape(Rechunk?(FromArray))) on the host, block by block, with
    the expression's own pairing."""
    node = arr.expr
    assert type(node).__name__ == "Reshape"
    src = node.operand("array")
    leaf = src
    while type(leaf).__name__ != "FromArray":
        leaf = leaf.operand("array")
    xh = leaf.operand("array")
    out = np.empty(node.shape, dtype=xh.dtype)
    for bid in itertools.product(*[range(n) for n in node.numblocks]):
        ibid = node.source(bid)
        start, shape = src.block_start(ibid), src.block_shape(ibid)
        blk = xh[tuple(slice(s, s + n) for s, n in zip(start, shape))]
        ostart, oshape = node.block_start(bid), node.block_shape(bid)
        out[tuple(slice(s, s + n) for s, n in zip(ostart, oshape))] = np.ascontiguousarray(blk).reshape(oshape)
    return out


@pytest.mark.parametrize("seed", range(6))
def test_block_pairing_reproduces_numpy_reshape(seed):
    rng = random.Random(seed)
    done = 0
    while done < 60:
        groups, inshape, outshape = rng.randint(1, 3), [], []
        for _ in range(groups):
            n = rng.choice([2, 4, 6, 8, 12, 16, 24])
            fs = [1, 1, 1]
            m, p = n, 2
            while m > 1:
                while m % p == 0:
                    fs[rng.randrange(3)] *= p
                    m //= p
                p += 1
            fs = [f for f in fs if f > 1] or [1]
            mode = rng.random()
            if mode < 0.3:
                inshape.append(n); outshape.append(n)
            elif mode < 0.65:
                inshape += fs; outshape.append(n)
            else:
                inshape.append(n); outshape += fs
            if rng.random() < 0.2:
                outshape.insert(rng.randrange(len(outshape) + 1), 1)
        inshape, outshape = tuple(inshape), tuple(outshape)
        if inshape == outshape:
            continue
        chunks = tuple(rng.choice([1, 2, 3, 5, n]) for n in inshape)
        xh = np.arange(int(np.prod(inshape))).reshape(inshape)
        try:
            y = da.from_array(xh, chunks=chunks).reshape(outshape)
        except NotImplementedError:
            continue
        assert y.shape == outshape and y.dtype == xh.dtype
        assert np.array_equal(_host_reshape(y), xh.reshape(outshape)), (inshape, outshape, chunks)
        done += 1


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build container)")
def test_randomised_equivalence_with_the_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _refshim

    _refshim.install()
    from dask_array.manipulation._reshape import reshape_rechunk as ref
    from golden_reshape_cases import random_case

    rng = random.Random(99)
    for _ in range(4000):
        inshape, outshape, inchunks = random_case(rng)
        try:
            want = ref(inshape, outshape, inchunks)[:2]
        except NotImplementedError:
            want = "NotImplementedError"
        except IndexError:
            continue                       # the reference's own crash on (1,) -> (1, 1, 1): planned fine here
        try:
            got = reshape_plan(inshape, outshape, inchunks)
        except NotImplementedError:
            got = "NotImplementedError"
        assert got == want, (inshape, outshape, inchunks)
