"""Golden fixtures for the reshape chunk planner: runs the reference's OWN ``reshape_rechunk``
(``/root/reference/dask_array/manipulation/_reshape.py:38-127``), unmodified, through ``_refshim`` on seeded random
requests and records ``(inchunks, outchunks)`` or "NotImplementedError".

Run by hand in the build container (never on the GPU box):  python tests/golden/generate_reshape.py
Writes tests/golden/reshape.json.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()
from dask_array.manipulation._reshape import reshape_rechunk  # noqa: E402
from golden_reshape_cases import random_case  # noqa: E402


def main():
    rng = random.Random(4242)
    cases = []
    while len(cases) < 400:
        inshape, outshape, inchunks = random_case(rng)
        try:
            a, b, _, _ = reshape_rechunk(inshape, outshape, inchunks)
            out = [[list(c) for c in a], [list(c) for c in b]]
        except NotImplementedError:
            out = "NotImplementedError"
        except IndexError:
            continue            # the reference's own crash on (1,) -> (1, 1, 1)
        cases.append(dict(inshape=list(inshape), outshape=list(outshape), inchunks=[list(c) for c in inchunks], out=out))
    path = os.path.join(HERE, "reshape.json")
    with open(path, "w") as f:
        json.dump(cases, f, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
