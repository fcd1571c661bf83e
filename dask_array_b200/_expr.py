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

import numpy as np


def normalize_chunks(chunks, shape):
    """Regular-chunk subset of ``_core_utils.py:731 normalize_chunks``."""
    if isinstance(chunks, (int, np.integer)):
        chunks = (int(chunks),) * len(shape)
    if len(chunks) != len(shape):
        raise ValueError(f"chunks {chunks} do not match shape {shape}")
    out = []
    for c, n in zip(chunks, shape):
        if isinstance(c, (tuple, list)):
            if sum(c) != n:
                raise ValueError(f"chunks {c} do not add up to {n}")
            out.append(tuple(int(v) for v in c))
        else:
            c = int(c)
            if c == -1 or c >= n:
                out.append((int(n),))
            else:
                if c <= 0:
                    raise ValueError(f"bad chunk size {c}")
                full, rest = divmod(n, c)
                out.append((c,) * full + ((rest,) if rest else ()))
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
