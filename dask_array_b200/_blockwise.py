"""Blockwise expressions, the fusion pass and the fused-kernel program builder.

Mirrors ``dask_array/_blockwise.py``: ``Elemwise`` (:837), the fusion pass
``optimize_blockwise_fusion_array`` (:1405-1571) with its conflict rule
``_remove_conflicting_exprs`` (:1342-1402), and ``FusedBlockwise`` (:1574-1738).  Where the
reference's ``FusedBlockwise._task`` builds a ``Task.fuse`` sub-graph that NumPy executes
operator by operator, ``FusedBlockwise.build_program`` here emits ONE kernel program
(``_codegen.Program``) whose inputs are the group's external dependencies, each with the
dimension map (``_symbolic_mapping`` :1298-1339) that says which block, and which
transposed / broadcast view of it, every output block reads.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

from . import _codegen as cg
from ._expr import ArrayExpr, BroadcastTrick


def _is_scalar(x) -> bool:
    return isinstance(x, (bool, int, float, np.generic)) or (isinstance(x, np.ndarray) and x.ndim == 0)


def unified_chunks(args):
    """Output chunks of an element-wise op and the chunks every array operand must have for it
    (``Blockwise.chunks`` -> ``unify_chunks_expr``, ``_blockwise.py:94-98``, ``_expr.py:723-905``).
    Returns ``(out_chunks, {operand name: target chunks})``; an operand used twice votes once, as in
    the reference (``seen`` keys, ``_expr.py:783-786``)."""
    from ._unify import unify

    arrs = {}
    for a in args:
        if isinstance(a, ArrayExpr):
            arrs.setdefault(a._name, a)
    uniq = list(arrs.values())
    out, targets = unify([(a.shape, a.chunks, a.dtype.itemsize) for a in uniq])
    return out, {a._name: t for a, t in zip(uniq, targets)}


def broadcast_chunks(args):
    """Output chunks of an element-wise op (NumPy right-aligned broadcasting on the block grid; 1-block
    dims of extent 1 broadcast, ``_blockwise.py:1243 _broadcast_block_id``)."""
    return unified_chunks(args)[0]


class Elemwise(ArrayExpr):
    """One NumPy ufunc / ``operator.*`` per block (``_blockwise.py:837-1074``).
    ``op`` is the canonical NumPy ufunc name; ``args`` mix expressions and Python scalars."""

    _parameters = ["op", "args", "kwargs"]
    _defaults = {"kwargs": ()}
    _is_blockwise_fusable = True

    def dependencies(self):
        return [a for a in self.operand("args") if isinstance(a, ArrayExpr)]

    def map_children(self, fn):
        return self._map_args(fn)

    @property
    def args(self):
        return self.operand("args")

    @property
    def chunks(self):
        if "chunks" not in self._cache:
            self._cache["chunks"] = broadcast_chunks(self.args)
        return self._cache["chunks"]

    @property
    def dtype(self):
        """``Elemwise._info`` (:928-966): run the op on 1-element dummies + raw scalars."""
        if "dtype" not in self._cache:
            kw = dict(self.operand("kwargs"))
            refs = []
            for a in self.args:
                if isinstance(a, ArrayExpr):
                    refs.append(cg.Ref("in", 0, a.dtype))
                elif isinstance(a, (np.generic, np.ndarray)):
                    refs.append(cg.Ref("const", -1, np.asarray(a).dtype, np.asarray(a)[()]))
                else:
                    refs.append(cg.Ref("const", -1, None, a))
            self._cache["dtype"] = cg.infer_dtype(self.operand("op"), refs, **kw)
        return self._cache["dtype"]

    def _tree_label(self):
        return f"Elemwise({self.operand('op')})"

    # expression operands live inside the ``args`` tuple: rewrite through it
    def _map_args(self, fn):
        new = tuple(fn(a) if isinstance(a, ArrayExpr) else a for a in self.args)
        if all(a is b for a, b in zip(new, self.args)):
            return self
        return Elemwise(self.operand("op"), new, self.operand("kwargs"))

    def _lower(self):
        """``Elemwise._lower`` (:1003): unify operand chunks, inserting Rechunk where needed."""
        from ._rechunk import Rechunk

        arrs = [a for a in self.args if isinstance(a, ArrayExpr)]
        if len(arrs) < 2:
            return None
        _, targets = unified_chunks(self.args)
        if all(targets[a._name] == a.chunks for a in arrs):
            return None
        return self._map_args(lambda a: a if targets[a._name] == a.chunks else Rechunk(a, targets[a._name]))


class Transpose(ArrayExpr):
    """``manipulation/_transpose.py:14-75``: a per-block ``np.transpose`` VIEW plus the same
    permutation of the block grid."""

    _parameters = ["array", "axes"]
    _is_blockwise_fusable = True

    @property
    def chunks(self):
        c = self.operand("array").chunks
        return tuple(c[a] for a in self.operand("axes"))

    @property
    def dtype(self):
        return self.operand("array").dtype

    def _simplify_down(self):
        arr, axes = self.operand("array"), tuple(self.operand("axes"))
        if axes == tuple(range(len(axes))):
            return arr
        if isinstance(arr, Transpose):                     # T(T(x)) -> one permutation (:77-110)
            inner = arr.operand("axes")
            return Transpose(arr.operand("array"), tuple(inner[a] for a in axes))
        return None

    def _tree_label(self):
        return f"Transpose{tuple(self.operand('axes'))}"


def _visit_children(expr, fn):
    return expr.map_children(fn)


# ----------------------------------------------------------------------------- fusion
class FusedBlockwise(ArrayExpr):
    """A group of block-aligned expressions executed as one kernel per device
    (``_blockwise.py:1574-1738``).  ``exprs[0]`` is the root (the group's output)."""

    _parameters = ["exprs"]

    @property
    def exprs(self):
        return self.operand("exprs")

    @property
    def root(self):
        return self.exprs[0]

    @property
    def chunks(self):
        return self.root.chunks

    @property
    def dtype(self):
        return self.root.dtype

    @property
    def _name(self):
        return "fused-" + self.root._name

    def dependencies(self):
        inside = {e._name for e in self.exprs}
        seen, out = set(), []
        for e in self.exprs:
            for d in e.dependencies():
                if d._name not in inside and d._name not in seen:
                    seen.add(d._name)
                    out.append(d)
        return out

    def substitute_operands(self, new_ops):
        return self if new_ops[0] is self.exprs else FusedBlockwise(new_ops[0])

    def _tree_label(self):
        names = ", ".join(e._tree_label() for e in self.exprs)
        return f"FusedBlockwise[{names}]"


def _symbolic_dep_maps(expr, my_map):
    """{dep name: [dimension maps]} -- which ROOT output dim each dim of a dependency follows
    (``_symbolic_mapping``, :1298-1339).  ``None`` marks a dim that follows no root dim."""
    out = defaultdict(list)
    if isinstance(expr, Transpose):
        axes = expr.operand("axes")
        inv = [0] * len(axes)
        for i, a in enumerate(axes):
            inv[a] = i
        dep = expr.operand("array")
        out[dep._name].append(tuple(my_map[inv[j]] for j in range(len(axes))))
    else:
        n = expr.ndim
        for dep in expr.dependencies():
            off = n - dep.ndim
            out[dep._name].append(tuple(my_map[off + d] for d in range(dep.ndim)))
    return out


def _remove_conflicting(group):
    """``_remove_conflicting_exprs`` (:1342-1402): an expression reached through two different
    block mappings (``a + a.T``) cannot live inside the group."""
    if len(group) <= 1:
        return group
    names = {e._name for e in group}
    by_name = {e._name: e for e in group}
    root = group[0]
    maps = {root._name: tuple(range(root.ndim))}
    conflicts = set()
    for e in group:
        if e._name not in maps:
            continue
        for dep_name, lst in _symbolic_dep_maps(e, maps[e._name]).items():
            if dep_name not in names:
                continue
            for m in lst:
                if dep_name in maps:
                    if maps[dep_name] != m:
                        conflicts.add(dep_name)
                else:
                    maps[dep_name] = m
    if not conflicts:
        return group
    remaining = names - conflicts
    reach, stack = {root._name}, [root]
    while stack:
        e = stack.pop()
        for d in e.dependencies():
            if d._name in remaining and d._name not in reach:
                reach.add(d._name)
                stack.append(by_name[d._name])
    return [e for e in group if e._name in reach]


def _external_reads(group) -> int:
    """Upper bound of the kernel inputs of a group: edges from members to array expressions outside it
    (constant leaves inside the group are immediates; a dependency read through two block mappings
    counts twice, as the program builder counts it)."""
    names = {e._name for e in group}
    return sum(1 for e in group for d in e.dependencies() if d._name not in names)


def _reachable(group):
    """Members still connected to the root through members."""
    if not group:
        return group
    names = {e._name for e in group}
    by = {e._name: e for e in group}
    reach, stack = {group[0]._name}, [group[0]]
    while stack:
        e = stack.pop()
        for d in e.dependencies():
            if d._name in names and d._name not in reach:
                reach.add(d._name)
                stack.append(by[d._name])
    return [e for e in group if e._name in reach]


def optimize_blockwise_fusion(expr):
    """``optimize_blockwise_fusion_array`` (:1405-1571): roots are fusable nodes without
    fusable dependents; a dependency joins a group when all its dependents are inside it."""
    fusable = lambda e: getattr(e, "_is_blockwise_fusable", False)
    seen, stack = set(), [expr]
    dependents, dependencies, by_name = defaultdict(set), {}, {}
    order = []
    while stack:
        node = stack.pop()
        if node._name in seen:
            continue
        seen.add(node._name)
        order.append(node)
        by_name[node._name] = node
        if fusable(node):
            dependencies.setdefault(node._name, set())
            dependents.setdefault(node._name, set())
        for dep in node.dependencies():
            stack.append(dep)
            by_name[dep._name] = dep
            dependents[dep._name].add(node._name)
            if fusable(dep) and fusable(node):
                dependencies[node._name].add(dep._name)
    roots = [by_name[k] for k in dependencies
             if not any(fusable(by_name[d]) for d in dependents.get(k, ()))]
    replacements, assigned = {}, set()
    while roots:
        root = roots.pop()
        if root._name in assigned:
            continue
        group, gstack, in_group = [], [root], set()
        while gstack:
            node = gstack.pop()
            if node._name in in_group or node._name in assigned:
                continue
            in_group.add(node._name)
            group.append(node)
            for dep_name in sorted(dependencies.get(node._name, ())):
                dep = by_name[dep_name]
                inside = in_group | {s._name for s in gstack}
                if dependents[dep_name] <= inside:
                    gstack.append(dep)
                elif dep_name not in {r._name for r in roots}:
                    roots.append(dep)
        group = _remove_conflicting(group)
        # a kernel takes at most B2_MAX_IN array inputs: while the group reads more, the operand of the
        # root with the largest sub-tree is cut out and becomes a group (a kernel, a materialised
        # intermediate) of its own
        while len(group) > 1 and _external_reads(group) > cg.MAX_INPUTS:
            names = {e._name for e in group}
            by = {e._name: e for e in group}

            def subtree(e, acc):
                if e._name in acc:
                    return acc
                acc.add(e._name)
                for d in e.dependencies():
                    if d._name in names:
                        subtree(by[d._name], acc)
                return acc
            # the member whose sub-tree swallows the most inputs while still fitting one kernel
            def reads(e):
                sub = subtree(e, set())
                return sum(1 for n in sub for d in by[n].dependencies() if d._name not in sub)
            cands = [(reads(e), len(subtree(e, set())), e._name, e) for e in group[1:]]
            cands = [c for c in cands if 1 < c[0] <= cg.MAX_INPUTS]
            if not cands:
                break
            victim = max(cands)[3]
            cut = subtree(victim, set())
            group = [e for e in group if e._name not in cut]
            # members still reachable from the root only (a cut sub-tree may have shared nodes with the rest)
            group = _remove_conflicting(_reachable(group))
            if victim._name not in {r._name for r in roots}:
                roots.append(victim)
        # every launch is a "fused" launch, also a group of one (single Elemwise / Transpose)
        replacements[group[0]._name] = FusedBlockwise(tuple(group))
        assigned.update(e._name for e in group)
    if not replacements:
        return expr

    memo = {}

    def rebuild(node):
        if node._name in memo:
            return memo[node._name]
        if node._name in replacements:
            fb = replacements[node._name]
            inside = {e._name: e for e in fb.exprs}
            rebuilt = {}

            def rebuild_inner(e, _inside=inside, _rebuilt=rebuilt):
                # children first (the group is a DAG, e.g. x + x*2 shares x), keyed by the OLD name
                if e._name not in _rebuilt:
                    _rebuilt[e._name] = e.map_children(
                        lambda o: rebuild_inner(o) if o._name in _inside else rebuild(o))
                return _rebuilt[e._name]

            for inner in fb.exprs:
                rebuild_inner(inner)
            out = FusedBlockwise(tuple(rebuilt[e._name] for e in fb.exprs))
        else:
            out = _visit_children(node, rebuild)
        memo[node._name] = out
        return out

    return rebuild(expr)


# ----------------------------------------------------------------------------- program builder
class FusedPlan:
    """Kernel program of one FusedBlockwise plus, per kernel input, the external
    dependency it reads and the dimension map relating it to the root's dims."""

    def __init__(self, fused: FusedBlockwise):
        from ._reductions import ChunkReduce

        self.fused = fused
        self.program = cg.Program()
        self.leaves = []            # [(dep expr, dimmap)] in kernel-input order
        self._leaf_index = {}
        inside = {e._name: e for e in fused.exprs}
        root = fused.root
        self.reduce = root if isinstance(root, ChunkReduce) else None
        # the chain is evaluated on the root's INPUT grid for a reduction chunk step
        top = root.operand("array") if self.reduce is not None else root
        self.eval_expr = top
        memo = {}

        def emit(e, dmap):
            key = (e._name, dmap)
            if key in memo:
                return memo[key]
            if e._name not in inside:
                ref = self._leaf(e, dmap)
                memo[key] = ref
                return ref
            if isinstance(e, BroadcastTrick):
                ref = self.program.typed_const(e.operand("value"), e.dtype)
            elif isinstance(e, Transpose):
                axes = e.operand("axes")
                inv = [0] * len(axes)
                for i, a in enumerate(axes):
                    inv[a] = i
                ref = emit(e.operand("array"), tuple(dmap[inv[j]] for j in range(len(axes))))
            elif isinstance(e, Elemwise):
                n = e.ndim
                refs = []
                for a in e.args:
                    if isinstance(a, ArrayExpr):
                        off = n - a.ndim
                        refs.append(emit(a, tuple(dmap[off + d] for d in range(a.ndim))))
                    else:
                        refs.append(self.program.const(a))
                ref = self.program.op(e.operand("op"), *refs, **dict(e.operand("kwargs")))
                if ref.dtype != e.dtype:
                    ref = self.program.op("astype", ref, dtype=e.dtype)
            else:
                raise NotImplementedError(f"{type(e).__name__} cannot be fused into a B200 kernel")
            memo[key] = ref
            return ref

        out = emit(top, tuple(range(top.ndim)))
        if out.kind != "tmp":
            # a bare leaf / constant at the top still needs one op so that there is a result
            if out.kind == "const":
                out = self.program.op("positive", self.program.typed_const(out.value, top.dtype))
            else:
                out = self.program.op("positive", out)
        self.program.set_output(out)

    @classmethod
    def from_reference(cls, fused):
        """Plan of a REFERENCE ``FusedBlockwise`` (``dask_array/_blockwise.py:1574-1728``): see
        ``plugin.fused_plan_from_reference``."""
        from .plugin import fused_plan_from_reference

        return fused_plan_from_reference(fused)

    def _leaf(self, dep, dmap):
        key = (dep._name, dmap)
        if key not in self._leaf_index:
            self._leaf_index[key] = self.program.add_input(dep.dtype)
            self.leaves.append((dep, dmap))
        return self._leaf_index[key]

    def leaf_block_id(self, leaf_idx, out_bid):
        """Block of the dependency that output block ``out_bid`` reads
        (``_compute_block_ids`` :1712-1728 + ``_broadcast_block_id`` :1243)."""
        dep, dmap = self.leaves[leaf_idx]
        nb = dep.numblocks
        return tuple(out_bid[dmap[d]] if nb[d] > 1 else 0 for d in range(dep.ndim))

    def leaf_strides(self, leaf_idx, chunk, ndim_out):
        """Element strides of ``chunk`` (a block of the dependency) along each root dim:
        transposes become swapped strides, broadcasts become 0."""
        dep, dmap = self.leaves[leaf_idx]
        st = [0] * ndim_out
        for d in range(dep.ndim):
            if chunk.shape[d] != 1:
                st[dmap[d]] = chunk.strides[d]
        return tuple(st)
