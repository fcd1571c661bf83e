"""Golden fixtures for host-side structure logic: the reference's OWN ``_slice_1d``
(``slicing/_utils.py:279-440``), ``_overlap_internal_chunks`` / ``ensure_minimum_chunksize`` /
``coerce_depth`` / ``coerce_boundary`` (``_overlap.py:29-50, 836-883, 1303-1362``), run unmodified through
``_refshim``.  Run by hand in the build container:  python tests/golden/generate_structure.py
Writes tests/golden/structure.json.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

_refshim.install()
from dask_array.slicing._utils import _slice_1d  # noqa: E402


def _reference_functions(path, names):
    """The module imports the whole collection layer (needs the real ``dask``); the helpers wanted here
    are pure functions, so their source is taken from the file verbatim (ast) and executed alone."""
    import ast
    import types
    from numbers import Integral, Number

    src = open(path).read()
    tree = ast.parse(src)
    ns = {"Integral": Integral, "Number": Number, "np": __import__("numpy")}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return types.SimpleNamespace(**{n: ns[n] for n in names})


O = _reference_functions(_refshim.REFERENCE_ROOT + "/dask_array/_overlap.py",
                         ["_overlap_internal_chunks", "ensure_minimum_chunksize", "coerce_depth", "coerce_depth_type",
                          "coerce_boundary"])


def enc_slice(v):
    return [v.start, v.stop, v.step] if isinstance(v, slice) else int(v)


def main():
    rng = random.Random(0)
    out = {"slice_1d": [], "overlap_chunks": [], "min_chunksize": [], "coerce": []}
    cases = [(100, [60, 40], slice(None, None, None)), (100, [20] * 5, slice(0, 35)), (100, [20, 10, 10, 10, 25, 25], slice(10, 35)),
             (100, [15, 14, 13, 58], slice(10, 41, 3)), (100, [20] * 5, slice(0, 100, 40)), (100, [20] * 5, 25),
             (100, [20] * 5, slice(100, 0, -3)), (100, [20] * 5, slice(100, 12, -3)), (100, [20] * 5, slice(100, -12, -3))]
    for _ in range(120):
        n = rng.randint(1, 80)
        cuts = sorted(rng.sample(range(1, n), min(n - 1, rng.randint(0, 5)))) if n > 1 else []
        lengths = [b - a for a, b in zip([0] + cuts, cuts + [n])]
        step = rng.choice([-7, -3, -2, -1, 1, 2, 3, 5, 11])
        start = rng.choice([None, rng.randint(-n - 3, n + 3)])
        stop = rng.choice([None, rng.randint(-n - 3, n + 3)])
        cases.append((n, lengths, slice(start, stop, step)))
    for n, lengths, index in cases:
        if isinstance(index, slice):            # the callers normalise first (normalize_index / slice.indices)
            s0, s1, st = index.indices(n)
            if len(range(s0, s1, st)) == 0:
                continue                        # empty selections never reach _slice_1d with these arguments
            if st < 0 and s1 < 0:
                s1 = None if index.stop is None or index.stop < -n else s1
            norm = slice(s0, s1, st)
            # the reference is called with what its own normalisation (sanitize / posify) produces
            got = _slice_1d(n, lengths, norm if st > 0 else slice(s0, None if s1 is None else s1, st))
            out["slice_1d"].append({"n": n, "lengths": lengths, "index": [index.start, index.stop, index.step],
                                    "blocks": [[int(k), enc_slice(v)] for k, v in got.items()]})
        else:
            got = _slice_1d(n, lengths, index)
            out["slice_1d"].append({"n": n, "lengths": lengths, "index": int(index),
                                    "blocks": [[int(k), enc_slice(v)] for k, v in got.items()]})
    for chunks, axes in [(((4, 4), (4, 4)), {0: 2, 1: 1}), (((10, 10, 10, 7), (25, 25)), {0: 3}), (((5,), (6, 6, 6)), {0: 2, 1: (1, 2)}),
                         (((8, 9, 10),), {0: (0, 3)}), (((6, 6), (6, 6), (12,)), {0: 1, 1: 2, 2: 5})]:
        got = O._overlap_internal_chunks(chunks, axes)
        out["overlap_chunks"].append({"chunks": [list(c) for c in chunks], "axes": {str(k): (list(v) if isinstance(v, tuple) else v) for k, v in axes.items()},
                                      "result": [list(c) for c in got]})
    for size, chunks in [(10, (20, 20, 1)), (3, (1, 1, 3)), (10, (20, 20, 1, 6)), (5, (3, 3, 3, 3)), (4, (10, 1, 1, 1, 10)), (2, (2, 2, 2)), (7, (1, 20, 1))]:
        out["min_chunksize"].append({"size": size, "chunks": list(chunks), "result": list(O.ensure_minimum_chunksize(size, chunks))})
    # 2000 random (size, chunks) cases of the same helper (ValueError recorded as null): the product's
    # rewrite is compared with these DATA, never with reference source executed at test time
    r2 = random.Random(3)
    for _ in range(2000):
        chunks = tuple(r2.randint(1, 30) for _ in range(r2.randint(1, 8)))
        size = r2.randint(1, 25)
        try:
            res = list(O.ensure_minimum_chunksize(size, chunks))
        except ValueError:
            res = None
        out["min_chunksize"].append({"size": size, "chunks": list(chunks), "result": res})
    for ndim, depth, boundary in [(2, 1, "reflect"), (3, {0: 2, 2: (1, 3)}, {1: "periodic"}), (2, (1, 2), None), (1, None, 5)]:
        d = O.coerce_depth(ndim, depth)
        b = O.coerce_boundary(ndim, boundary)
        out["coerce"].append({"ndim": ndim, "depth": repr(depth), "boundary": repr(boundary),
                              "depth_out": {str(k): (list(v) if isinstance(v, tuple) else v) for k, v in d.items()},
                              "boundary_out": {str(k): v for k, v in b.items()}})
    with open(os.path.join(HERE, "structure.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(len(out["slice_1d"]), "slice cases;", out["overlap_chunks"][2], out["min_chunksize"][2])


if __name__ == "__main__":
    main()
