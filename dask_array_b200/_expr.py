"""Expression tree of the B200 backend (host side).

A small mirror of the reference's ``ArrayExpr`` system (``dask_array/_expr.py:74-583``) --
only what the hot path of SURVEY.md section 8 needs: an expression records WHAT to compute
(operands, chunks, dtype); ``optimize()`` runs ``simplify -> lower -> fuse`` like
``_expr.py:506-515`` / ``_materialize.py:34-104``; the executor (``_executor.py``) then launches one
kernel per fused expression and device.  The upstream ``dask`` rewrite engine is not
available (not installed), so the handful of rewrite rules the configs exercise are
implemented directly on these classes.
"""
from __future__ import annotations

import hashlib
import itertools
import math
from numbers import Number

import numpy as np


CHUNK_SIZE_BYTES = 128 * 2**20          # ``array.chunk-size`` default ("128MiB", dask 2025.12 array schema)

_BYTE_UNITS = {"": 1, "b": 1, "kb": 10**3, "mb": 10**6, "gb": 10**9, "tb": 10**12, "pb": 10**15,
               "kib": 2**10, "mib": 2**20, "gib": 2**30, "tib": 2**40, "pib": 2**50,
               "k": 10**3, "m": 10**6, "g": 10**9, "t": 10**12, "p": 10**15,
               "ki": 2**10, "mi": 2**20, "gi": 2**30, "ti": 2**40, "pi": 2**50}


def _parse_bytes(text) -> int:
    """``dask.utils.parse_bytes``: "128MiB" -> 134217728, "1e6 kB" -> 10**9, 123 -> 123."""
    if isinstance(text, (int, float, np.integer, np.floating)):
        return int(text)
    t = str(text).replace(" ", "")
    k = len(t)
    while k and t[k - 1].isalpha():
        k -= 1
    number, unit = t[:k] or "1", t[k:].lower()
    if unit not in _BYTE_UNITS:
        raise ValueError(f"Could not interpret {unit!r} as a byte unit")
    return int(float(number) * _BYTE_UNITS[unit])


def _regular(n, c):
    """``blockdims_from_blockshape`` for one axis: ``c`` repeated, the remainder last; a zero-length axis is (0,)."""
    n, c = int(n), int(c)
    if n == 0:
        return (0,)
    if c <= 0:
        raise ValueError(f"chunk size must be positive, got {c}")
    full, rest = divmod(n, c)
    return (c,) * full + ((rest,) if rest else ())


CHUNK_SIZE_TOLERANCE = 1.25             # ``array.chunk-size-tolerance`` default (dask 2025.12)


def _auto_from_previous(chunks, shape, limit, itemsize, previous):
    """The ``previous_chunks`` branch of ``auto_chunks`` (``_core_utils.py:584-660``), the one ``rechunk("auto")``
    takes: grow or shrink the "auto" axes from the MEDIAN of their current block lengths by a common factor so that
    a block holds ``limit`` bytes.  Growing keeps existing block boundaries (neighbouring blocks are merged up to the
    proposed length); shrinking -- or growing an axis whose largest block already exceeds the proposal by the
    tolerance -- picks a regular edge rounded to the axis' dominant block length.  An axis that outgrows its length
    becomes one block and the factor is redistributed over the remaining axes."""
    autos = {i for i, c in enumerate(chunks) if isinstance(c, str)}
    fixed = math.prod(c if isinstance(c, Number) else max(c) for c in chunks if not isinstance(c, str))
    median = {a: np.median(previous[a]) for a in autos}

    def factor(sizes):
        return limit / itemsize / fixed / math.prod(max(r) if isinstance(r, tuple) else r for r in sizes.values() if r)

    mult = factor(median)
    shrinking = mult < 1
    # when shrinking, the running sizes ARE the result: every pass refines them and the factor is recomputed from them
    result = median if shrinking else {}
    dominant = []
    for i, n in enumerate(shape):
        counts = {}
        for c in previous[i]:
            counts[c] = counts.get(c, 0) + 1
        mode, count = max(counts.items(), key=lambda kv: kv[1])
        dominant.append(mode if mode > 1 and count >= len(previous[i]) / 2 else n)
    again = True
    while again:
        live = len(autos)
        again = False
        for a in sorted(autos):
            proposed = median[a] * mult ** (1 / live)
            ceiling = proposed * CHUNK_SIZE_TOLERANCE ** (1 / live)
            if proposed > shape[a]:
                # the axis is exhausted: one block, and the others share what is left of the factor
                chunks[a] = shape[a]
                autos.remove(a)
                del median[a]
                fixed *= shape[a]
                result[a] = (shape[a],)
                again = True
            elif shrinking or max(previous[a]) > ceiling:
                step = dominant[a]
                result[a] = max(1, int(proposed)) if proposed <= step else proposed // step * step
                if proposed < 1:
                    again = True
                    autos.discard(a)
            else:
                merged, run = [], 0
                for c in previous[a]:
                    if run + c <= proposed:
                        run += c
                    else:
                        if run > 0:
                            merged.append(run)
                        run = c
                if run > 0:
                    merged.append(run)
                result[a] = tuple(merged)
        if again or shrinking:
            before, mult = mult, factor(median)
            again = again or mult != before
    for a, v in result.items():
        chunks[a] = v if v else 0
    return tuple(chunks)


def _auto_sizes(chunks, shape, limit, dtype, previous=None):
    """``auto_chunks`` (``_core_utils.py:524-677``).  Without ``previous``: every "auto" axis gets the same edge
    length such that one block is ``limit`` bytes; axes shorter than that become one block and the edge is
    recomputed for the rest (:662-677).  With ``previous``: ``_auto_from_previous``."""
    chunks = list(chunks)
    if dtype is None:
        raise TypeError("dtype must be known for auto-chunking")
    dtype = np.dtype(dtype)
    if dtype.hasobject:
        raise NotImplementedError("Can not use auto rechunking with object dtype")
    if dtype.itemsize == 0:
        raise ValueError("auto-chunking with dtype.itemsize == 0 is not supported, please pass in `chunks` explicitly")
    limit = max(1, CHUNK_SIZE_BYTES if limit is None else _parse_bytes(limit))
    if previous:
        return _auto_from_previous(chunks, shape, limit, dtype.itemsize, previous)
    while True:
        autos = [i for i, c in enumerate(chunks) if isinstance(c, str)]
        if not autos:
            return tuple(chunks)
        fixed = math.prod(c if isinstance(c, Number) else max(c) for c in chunks if not isinstance(c, str))
        edge = (limit / dtype.itemsize / fixed) ** (1 / len(autos))
        small = [i for i in autos if shape[i] < edge]
        if small:
            for i in small:
                chunks[i] = (int(shape[i]),)
            continue
        for i in autos:
            # ``round_to`` (:507-521): edge <= axis length here, so the edge itself with a ragged last block
            chunks[i] = max(1, int(edge))
        return tuple(chunks)


def normalize_chunks(chunks, shape, dtype=None, limit=None, previous_chunks=None):
    """``normalize_chunks`` (``_core_utils.py:731-885``) for known shapes: a block edge for every axis (int), one
    per axis (tuple of ints; -1 / None = the whole axis), explicit block lengths per axis (zero-length blocks
    allowed), ``{axis: size}``, "auto" and byte-size strings ("1kiB") -- edges chosen so a block holds
    ``array.chunk-size`` (128 MiB) of ``dtype`` -- and the 0-d / zero-size conventions (``(1,)`` on ``()`` -> ``()``,
    ``()`` on ``(0, 0)`` -> ``((0,), (0,))``).  ``previous_chunks=`` (the current blocks of an array being
    re-chunked) makes "auto" scale those instead of starting from a cube."""
    shape = tuple(int(n) for n in shape)
    if chunks is None:
        raise ValueError("You must specify a chunks= keyword argument.")
    if isinstance(chunks, np.ndarray):
        chunks = chunks.tolist()
    if isinstance(chunks, list):
        chunks = tuple(chunks)
    if isinstance(chunks, (Number, str)):
        chunks = (chunks,) * len(shape)
    if isinstance(chunks, dict):
        chunks = tuple(chunks.get(i, None) for i in range(len(shape)))
    if not chunks and shape and all(n == 0 for n in shape):
        chunks = ((0,),) * len(shape)
    if len(shape) == 1 and len(chunks) > 1 and all(isinstance(c, (Number, str)) for c in chunks):
        if any(isinstance(c, str) for c in chunks):
            raise ValueError(f"String values are not supported inside explicit chunk tuples. Got chunks={chunks}")
        chunks = (chunks,)
    if not shape:
        return ()                                                # ``normalize_chunks((1,), ())`` -> ``()``
    if len(chunks) != len(shape):
        raise ValueError(f"Chunks and shape must be of the same length/dimension. Got chunks={chunks}, shape={shape}")
    chunks = tuple(n if c is None or (isinstance(c, Number) and c == -1) else c for c, n in zip(chunks, shape))
    for c in chunks:
        if isinstance(c, str) and c != "auto":
            text = c.replace(" ", "")
            if not text or not text[-1].isalpha():
                raise ValueError("String chunk sizes must be 'auto' or byte sizes with a byte unit like 'B', 'MB', "
                                 f"or 'MiB'. Got {c!r}")
            parsed = _parse_bytes(c)
            if parsed < 0:
                raise ValueError(f"String chunk byte sizes must not be negative. Got {c!r}")
            if limit is None:
                limit = parsed
            elif parsed != _parse_bytes(limit):
                raise ValueError(f"Only one consistent value of limit or chunk is allowed. Used {parsed} != {limit}")
    chunks = tuple("auto" if isinstance(c, str) else c for c in chunks)
    if any(isinstance(c, str) for c in chunks):
        prev = None
        if previous_chunks is not None:
            prev = tuple(_regular(n, c) if isinstance(c, Number) else tuple(c) for n, c in zip(shape, previous_chunks))
        chunks = _auto_sizes(chunks, shape, limit, dtype, prev)
    out = []
    for c, n in zip(chunks, shape):
        if isinstance(c, (tuple, list)):
            if not c:
                raise ValueError("Empty tuples are not allowed in chunks. Express zero length dimensions with 0(s) in chunks")
            if sum(c) != n:
                raise ValueError(f"Chunks do not add up to shape. Got chunks={chunks}, shape={shape}")
            out.append(tuple(int(v) for v in c))
        else:
            out.append(_regular(n, c))
    return tuple(out)


def _tok(obj) -> str:
    if isinstance(obj, ArrayExpr):
        return obj._name
    if isinstance(obj, np.ndarray):
        h = hashlib.sha1()
        h.update(str((obj.shape, obj.dtype.str)).encode())
        h.update(np.ascontiguousarray(obj).view(np.uint8).reshape(-1)[: 1 << 20].tobytes())
        h.update(str(id(obj)).encode() if obj.nbytes > (1 << 20) else b"")
        return "nd" + h.hexdigest()[:12]
    if isinstance(obj, (tuple, list)):
        return "(" + ",".join(_tok(o) for o in obj) + ")"
    if isinstance(obj, dict):
        return "{" + ",".join(f"{_tok(k)}:{_tok(v)}" for k, v in sorted(obj.items(), key=lambda kv: str(kv[0]))) + "}"
    if callable(obj):
        return _tok_callable(obj)
    return repr(obj)


def _tok_callable(fn) -> str:
    """Identity of a user function inside an expression name: module-level importable callables (NumPy
    ufuncs, library functions) by their qualified name; lambdas / closures / partials by their code object,
    constants, defaults and closure contents -- two different lambdas never share a name, the same source
    evaluated twice does (the reference tokenises ``func`` the same way and treats ``token=`` as a prefix,
    ``core/_blockwise_funcs.py``)."""
    import functools

    if isinstance(fn, functools.partial):
        return f"partial({_tok_callable(fn.func)},{_tok(fn.args)},{_tok(fn.keywords)})"
    code = getattr(fn, "__code__", None)
    name = f"{getattr(fn, '__module__', '')}.{getattr(fn, '__qualname__', getattr(fn, '__name__', type(fn).__name__))}"
    if code is None:
        return f"fn:{name}" if hasattr(fn, "__name__") else f"fn:{name}@{id(fn):x}"
    h = hashlib.sha1()
    h.update(code.co_code)
    h.update(repr(code.co_consts).encode())
    h.update(repr(code.co_names).encode())
    h.update(repr(getattr(fn, "__defaults__", None)).encode())
    for cell in getattr(fn, "__closure__", None) or ():
        try:
            h.update(_tok(cell.cell_contents).encode())
        except ValueError:                      # empty cell
            h.update(b"<empty>")
    return f"fn:{name}#{h.hexdigest()[:12]}"


class ArrayExpr:
    """Base class.  Sub-classes list ``_parameters``; operands are positional."""

    _parameters: list = []
    _defaults: dict = {}
    _is_blockwise_fusable = False      # _blockwise.py:186-209

    def __init__(self, *operands, **kw):
        ops = list(operands)
        for p in self._parameters[len(ops):]:
            if p in kw:
                ops.append(kw.pop(p))
            elif p in self._defaults:
                ops.append(self._defaults[p])
            else:
                raise TypeError(f"{type(self).__name__} missing operand {p!r}")
        if kw:
            raise TypeError(f"unexpected operands {sorted(kw)}")
        self.operands = ops
        self._cache = {}

    def operand(self, name):
        return self.operands[self._parameters.index(name)]

    def __getattr__(self, name):
        if name.startswith("_") and name not in ("_name",):
            raise AttributeError(name)
        params = type(self)._parameters
        if name in params:
            return self.operands[params.index(name)]
        raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")

    # ---- identity
    @property
    def _prefix(self) -> str:
        return type(self).__name__.lower()

    @property
    def _name(self) -> str:
        if "name" not in self._cache:
            h = hashlib.sha1(("|".join([type(self).__name__] + [_tok(o) for o in self.operands])).encode())
            self._cache["name"] = f"{self._prefix}-{h.hexdigest()[:16]}"
        return self._cache["name"]

    def dependencies(self) -> list:
        return [o for o in self.operands if isinstance(o, ArrayExpr)]

    # ---- array metadata
    @property
    def chunks(self) -> tuple:
        raise NotImplementedError

    @property
    def dtype(self) -> np.dtype:
        raise NotImplementedError

    @property
    def shape(self) -> tuple:
        return tuple(sum(c) for c in self.chunks)

    @property
    def ndim(self) -> int:
        return len(self.chunks)

    @property
    def numblocks(self) -> tuple:
        return tuple(len(c) for c in self.chunks)

    @property
    def size(self) -> int:
        return math.prod(self.shape)

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    def block_ids(self):
        return itertools.product(*[range(n) for n in self.numblocks])

    def block_shape(self, bid) -> tuple:
        return tuple(self.chunks[d][i] for d, i in enumerate(bid))

    def block_start(self, bid) -> tuple:
        return tuple(sum(self.chunks[d][:i]) for d, i in enumerate(bid))

    # ---- rewriting
    def map_children(self, fn):
        """Rebuild with ``fn`` applied to every child expression."""
        return self.substitute_operands([fn(o) if isinstance(o, ArrayExpr) else o for o in self.operands])

    def substitute_operands(self, new_ops):
        if all(a is b for a, b in zip(new_ops, self.operands)):
            return self
        return type(self)(*new_ops)

    def _simplify_down(self):
        """Return a replacement expression or None (reference: per-class ``_simplify_down``)."""
        return None

    def _lower(self):
        return None

    def simplify(self):
        return _rewrite(self, "_simplify_down")

    def _lower_rechunk(self):
        return None

    def lower_completely(self):
        return _rewrite(_rewrite(self, "_lower"), "_lower_rechunk")

    def optimize(self, fuse: bool = True):
        """simplify -> rechunk pushdown -> lower (chunk unification inserts rechunks) -> pushdown again
        (``test_lower_inserted_rechunk_pushes_into_from_array``) -> Rechunk -> TasksRechunk -> fuse."""
        from ._blockwise import optimize_blockwise_fusion
        from ._rechunk import pushdown_rechunks

        expr = pushdown_rechunks(self.simplify())
        expr = pushdown_rechunks(_rewrite(expr, "_lower")).simplify()
        expr = _rewrite(expr, "_lower_rechunk").simplify()
        return optimize_blockwise_fusion(expr) if fuse else expr

    # ---- display (README ``pprint``)
    def _tree_label(self) -> str:
        return type(self).__name__

    def tree_repr(self, indent=0) -> str:
        lines = ["  " * indent + self._tree_label()]
        for d in self.dependencies():
            lines.append(d.tree_repr(indent + 1))
        return "\n".join(lines)

    def pprint(self):
        print(self.tree_repr())

    def __repr__(self):
        return f"<{type(self).__name__} {self._name} shape={self.shape} dtype={self.dtype}>"


def _rewrite(expr: ArrayExpr, hook: str) -> ArrayExpr:
    """Bottom-up fixpoint of a per-class rewrite hook (the role of dask._expr's
    simplify / lower loops, ``_materialize.py:34-47``)."""
    memo = {}

    def visit(node):
        if node._name in memo:
            return memo[node._name]
        cur = node
        for _ in range(64):
            cur = cur.map_children(visit)
            out = getattr(cur, hook)()
            if out is None or out._name == cur._name:
                break
            cur = out
        memo[node._name] = cur
        return cur

    return visit(expr)


# ----------------------------------------------------------------------------- leaves
class FromArray(ArrayExpr):
    """Host NumPy array cut into blocks (``io/_from_array.py:60``).  Blocks are uploaded to
    the owning GPU when first needed and stay resident."""

    _parameters = ["array", "chunks_"]

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return self.operand("array").dtype

    def _tree_label(self):
        return f"FromArray{self.shape}"


class HostBlocks(ArrayExpr):
    """Per-block host arrays (``get_block(block id) -> ndarray``), e.g. the output of a host
    RNG or loader -- the role of ``from_map`` / ``from_delayed`` leaves (``io/``).  Only the
    blocks this rank owns are ever requested."""

    _parameters = ["get_block", "chunks_", "dtype_", "token"]

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    @property
    def _name(self):
        return f"hostblocks-{self.operand('token')}"

    def _tree_label(self):
        return f"HostBlocks{self.shape}"


class Resident(ArrayExpr):
    """Blocks already on the device(s) under their original keys -- what ``persist()``
    leaves behind (``_collection.py:285-300`` -> ``FromGraph`` ``io/_from_graph.py:12``)."""

    _parameters = ["store", "chunks_", "dtype_", "token"]

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    @property
    def _name(self):
        return f"resident-{self.operand('token')}"

    def _tree_label(self):
        return f"Resident{self.shape}"


class Random(ArrayExpr):
    """``random/_expr.py:63-250``: per-block generators seeded from
    ``SeedSequence(seed).spawn(nblocks)`` (:29-32, :97-126).  Host RNG, staged once."""

    _parameters = ["seed", "distribution", "shape_", "chunks_", "dtype_", "args"]
    _defaults = {"args": ()}

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    def block_seeds(self):
        """``_spawn_bitgens`` (``random/_expr.py:29-32``): the children the generator's live SeedSequence
        spawned for THIS draw.  ``seed`` = (entropy, spawn_key, children spawned before this draw)."""
        entropy, spawn_key, first = self.operand("seed")
        nb = self.numblocks
        return np.random.SeedSequence(entropy, spawn_key=tuple(spawn_key), n_children_spawned=first).spawn(math.prod(nb) or 1)

    def host_block(self, bid) -> np.ndarray:
        nb = self.numblocks
        flat = int(np.ravel_multi_index(bid, nb)) if nb else 0
        child = self.block_seeds()[flat]
        gen = np.random.Generator(np.random.PCG64(child))
        dist = self.operand("distribution")
        shape = self.block_shape(bid)
        if dist == "random":
            return gen.random(shape, dtype=self.dtype)
        if dist == "standard_normal":
            return gen.standard_normal(shape, dtype=self.dtype)
        if dist == "integers":
            lo, hi = self.operand("args")
            return gen.integers(lo, hi, size=shape, dtype=self.dtype)
        raise NotImplementedError(f"random distribution {dist!r}")

    def _tree_label(self):
        return f"Random{self.shape}"


class BroadcastTrick(ArrayExpr):
    """Ones / Zeros / Full (``creation/_ones_zeros.py:17-137``): a constant leaf.  Fusable:
    inside a fused kernel it is an immediate; materialised on its own it is a fill."""

    _parameters = ["value", "shape_", "chunks_", "dtype_"]
    _is_blockwise_fusable = True

    @property
    def chunks(self):
        return self.operand("chunks_")

    @property
    def dtype(self):
        return np.dtype(self.operand("dtype_"))

    def _tree_label(self):
        v = self.operand("value")
        nm = {0: "Zeros", 1: "Ones"}.get(v if isinstance(v, (int, float)) else None, "Full")
        return f"{nm}{self.shape}"
