"""Executor: runs an optimised expression tree on the GPU(s).

Takes the place of the scheduler + task graph of the reference (``dask.threaded.get`` chosen by
``Array.__dask_scheduler__``, ``_collection.py:111``; graph emission in every ``_layer()``):
instead of one Python task per block there is ONE kernel launch per (expression, device)
that covers every resident block.

Placement: blocks are dealt block-cyclically,
``owner(block) = ravel(block id) mod world``; one process per GPU (``torch.distributed``).
Per-block partials of a tree reduction are tiny, so they are all-gathered and every rank
folds the tree redundantly (results replicated); a rechunk / transposed read that crosses
the partition exchanges the needed blocks over NCCL before the local gather.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch

from . import _codegen as cg
from . import _lib
from . import _peer
from . import _runtime as rt
from ._blockwise import FusedBlockwise, FusedPlan
from ._device import DeviceChunk, alloc_bytes, contiguous_strides as _contig
from ._expr import ArrayExpr
from ._rechunk import TasksRechunk
from ._reductions import REDOPS, ArgChunk, CumReduction, PartialReduce


import os as _os

# B2_NVTX: 0 = off, 1 (default) = one range per expression while it is planned and first launched,
# 2 = the replayed tape carries the same ranges (two host calls per expression and replay)
_NVTX = int(_os.environ.get("B2_NVTX", "1") or 0)


class BlockStore:
    """Blocks of one expression held by this rank.  ``kind``: array | mean | moment | arg."""

    def __init__(self, expr, kind="array", replicated=False):
        self.expr = expr
        self.kind = kind
        self.replicated = replicated
        self.blocks = {}          # bid -> DeviceChunk | dict
        self.keepalive = []       # launch tables / pointer tables the stream may still read
        self.slab = None          # partial blocks laid out by _exchange.partials_layout (send buffer of the all-gather)


class World:
    """Process group facts.  world == 1 needs no torch.distributed at all."""

    def __init__(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            self.rank, self.size = dist.get_rank(), dist.get_world_size()
        else:
            self.rank, self.size = 0, 1

    def owner(self, expr, bid) -> int:
        return owner_of(expr, bid, self.size)


def _fill_identity(out, kind):
    """Write the identity of fold ``kind`` into the partial ``out`` (zero-size blocks)."""
    dt = out.dtype
    if kind in ("prod", "all"):
        val = 1
    elif kind in ("min", "nanmin", "max", "nanmax"):
        low = kind in ("max", "nanmax")
        if dt.kind == "f" or dt.name == "bfloat16":
            val = -np.inf if low else np.inf
        elif dt.kind == "b":
            val = not low
        else:
            info = np.iinfo(dt)
            val = info.min if low else info.max
    else:
        val = 0
    DeviceChunk.from_numpy(np.full(out.shape, val, dtype=dt), out=out)


class Executor:
    def __init__(self, world: World | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("dask_array_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.world = world or World()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.results: dict[str, BlockStore] = {}
        self.tape = []            # every device action in issue order: replaying it repeats the step
        self.collectives = []     # tape indices of the actions that synchronise with other ranks
        self.tape_span = {}       # expression name -> (first, end) tape indices of its own device actions

    def _do(self, fn, collective: bool = False):
        """Run a device action now and record it on the tape.  ``collective``: the action synchronises with the
        other ranks (peer barrier / all-gather, NCCL call): such actions keep their global order on every rank,
        so a tape that holds any is never split over concurrent streams."""
        fn()
        if collective:
            self.collectives.append(len(self.tape))
        self.tape.append(fn)

    # ------------------------------------------------------------------ driver
    def run(self, expr: ArrayExpr) -> BlockStore:
        if expr._name in self.results:
            return self.results[expr._name]
        for dep in expr.dependencies():
            self.run(dep)
        fn = getattr(self, "_run_" + type(expr).__name__, None)
        if fn is None:
            raise NotImplementedError(
                f"{type(expr).__name__} has no B200 execution path (optimise the expression first; "
                "there is no CPU fallback)")
        # one NVTX range per expression (planning + first launch) and, through the tape markers, per replay:
        # the timeline of a profiler shows which fused expression a kernel belongs to
        label = f"b2:{type(expr).__name__}:{expr._name[-12:]}"
        first = len(self.tape)
        if _NVTX:
            torch.cuda.nvtx.range_push(label)
        try:
            store = fn(expr)
        finally:
            if _NVTX:
                torch.cuda.nvtx.range_pop()
        self.tape_span[expr._name] = (first, len(self.tape))      # tape entries this expression itself added
        if _NVTX >= 2 and len(self.tape) > first:
            self.tape.insert(first, lambda label=label: torch.cuda.nvtx.range_push(label))
            self.tape.append(torch.cuda.nvtx.range_pop)
        self.results[expr._name] = store
        return store

    def mine(self, expr, bid) -> bool:
        return self.world.owner(expr, bid) == self.world.rank

    # ------------------------------------------------------------------ leaves
    def _run_FromArray(self, expr):
        st = BlockStore(expr)
        arr = expr.operand("array")
        for bid in expr.block_ids():
            if not self.mine(expr, bid):
                continue
            start, shape = expr.block_start(bid), expr.block_shape(bid)
            sl = tuple(slice(s, s + n) for s, n in zip(start, shape))
            host = arr[sl]
            chunk = DeviceChunk.empty(host.shape, host.dtype, self.device)
            self._do(lambda h=host, c=chunk: DeviceChunk.from_numpy(h, self.device, out=c))
            st.blocks[bid] = chunk
        return st

    def _run_HostBlocks(self, expr):
        st = BlockStore(expr)
        get = expr.operand("get_block")
        for bid in expr.block_ids():
            if not self.mine(expr, bid):
                continue
            host = get(bid)
            if host.shape != expr.block_shape(bid) or host.dtype != expr.dtype:
                raise ValueError(f"host block {bid} is {host.shape}/{host.dtype}, expected "
                                 f"{expr.block_shape(bid)}/{expr.dtype}")
            chunk = DeviceChunk.empty(host.shape, host.dtype, self.device)
            self._do(lambda h=host, c=chunk: DeviceChunk.from_numpy(h, self.device, out=c))
            st.blocks[bid] = chunk
        return st

    def _run_Random(self, expr):
        st = BlockStore(expr)
        for bid in expr.block_ids():
            if self.mine(expr, bid):
                st.blocks[bid] = DeviceChunk.from_numpy(expr.host_block(bid), self.device)
        return st

    def _run_Resident(self, expr):
        return expr.operand("store")

    def _run_BroadcastTrick(self, expr):
        """A constant leaf that was NOT fused (e.g. evicted by the ``a + a.T`` conflict rule):
        ONE allocation and ONE fill launch for all resident blocks, carved into per-block views."""
        st = BlockStore(expr)
        ids = [bid for bid in expr.block_ids() if self.mine(expr, bid)]
        item = expr.dtype.itemsize
        sizes = [-(-math.prod(expr.block_shape(bid)) * item // 256) * 256 for bid in ids]
        total = sum(sizes)
        buf = alloc_bytes(total, self.device)
        off = 0
        for bid, nb in zip(ids, sizes):
            st.blocks[bid] = DeviceChunk(buf, expr.block_shape(bid), expr.dtype, offset=off // item)
            off += nb
        if total:
            whole = DeviceChunk(buf, (total // item,), expr.dtype)
            self._do(lambda c=whole, v=expr.operand("value"): rt.fill(c, v))
        return st

    # ------------------------------------------------------------------ views
    def _run_SliceSlicesIntegers(self, expr):
        src = self.results[expr.operand("array")._name]
        st = BlockStore(expr)
        x = expr.operand("array")
        moves = []
        for bid in expr.block_ids():
            ibid, idx = expr.source(bid)
            if self.world.size > 1 and not src.replicated and self.world.owner(x, ibid) != self.world.owner(expr, bid):
                moves.append((bid, self.world.owner(x, ibid), (lambda blk, idx=idx: blk[idx]), ibid))
                continue
            if ibid in src.blocks and (src.replicated or self.mine(expr, bid)):
                st.blocks[bid] = src.blocks[ibid][idx]
        st.replicated = src.replicated
        if moves:
            _push_views(self, expr, st, src, moves)
        return st

    def _view_store(self, expr, src_of, view_of):
        """Zero-copy structural expressions: each output block is a view of one input block."""
        st = BlockStore(expr)
        rep = None
        moves = {}
        for bid in expr.block_ids():
            store, x, ibid = src_of(bid)
            rep = store.replicated if rep is None else (rep and store.replicated)
            if self.world.size > 1 and not store.replicated and self.world.owner(x, ibid) != self.world.owner(expr, bid):
                moves.setdefault(id(store), (store, []))[1].append(
                    (bid, self.world.owner(x, ibid), (lambda blk, bid=bid: view_of(blk, bid)), ibid))
                continue
            blk = store.blocks.get(ibid)
            if blk is not None and (store.replicated or self.mine(expr, bid)):
                st.blocks[bid] = view_of(blk, bid)
        st.replicated = bool(rep)
        for store, mv in moves.values():
            _push_views(self, expr, st, store, mv)
        return st

    def _run_ExpandDims(self, expr):
        x = expr.operand("array")
        src = self.results[x._name]
        return self._view_store(expr, lambda bid: (src, x, expr.source(bid)), lambda blk, bid: expr.view(blk))

    def _run_TrimInternal(self, expr):
        x = expr.operand("array")
        src = self.results[x._name]
        return self._view_store(expr, lambda bid: (src, x, bid), lambda blk, bid: expr.view(blk, bid))

    def _run_OverlapInternal(self, expr):
        """Halo exchange = the rechunk executor on ``OverlapInternal.pieces``: one gather per device,
        neighbour rims stored into peer memory when the neighbour block lives on another GPU."""
        return self._run_TasksRechunk(expr)

    def _run_MapBlocks(self, expr):
        """``func(*blocks)`` per block on the chunk type (NEP-13 / NEP-18 launches).  The result is
        copied into a block allocated once, so that a replayed tape keeps feeding the same memory to
        the launches recorded after it."""
        from . import _eager

        func = expr.operand("func")
        template, kw = expr.operand("kwargs")
        kwargs = dict(kw)
        stores = [self.results[a._name] for a in expr.operand("arrays")]
        st = BlockStore(expr)
        work = []
        for bid in expr.block_ids():
            if not self.mine(expr, bid):
                continue
            ins = []
            for a, s_ in zip(expr.operand("arrays"), stores):
                blk = s_.blocks.get(bid[:a.ndim])            # trailing output dims are new axes (one block)
                if blk is None:
                    raise RuntimeError(f"map_blocks: block {bid} of an argument is not resident on rank {self.world.rank}")
                ins.append(blk)
            work.append((bid, expr.block_shape(bid), ins))
        ident = cg.Program()
        ident.set_output(ident.op("astype", ident.add_input(expr.dtype), dtype=expr.dtype))

        def run():
            for bid, shape, ins in work:
                it = iter(ins)
                args = [next(it) if t is None else t for t in template]
                res = func(*args, **kwargs)
                if not isinstance(res, DeviceChunk):
                    raise TypeError(f"map_blocks: {getattr(func, '__name__', func)!r} returned {type(res).__name__}, "
                                    "not a DeviceChunk (no host fallback: use NumPy functions the chunk type implements)")
                if res.shape != tuple(shape) or res.dtype != expr.dtype:
                    raise ValueError(f"map_blocks: block result {res.shape} {res.dtype} does not match the declared "
                                     f"chunks / dtype {tuple(shape)} {expr.dtype}")
                prev = st.blocks.get(bid)
                if any(res.buf is b.buf for b in ins):
                    # a VIEW of an argument block (slices, sliding windows ...): pointer-stable across replays,
                    # so it is the output block itself -- nothing is materialised
                    if prev is not None and (prev.ptr, prev.strides) != (res.ptr, res.strides):
                        raise RuntimeError("map_blocks: a view result changed between replays")
                    st.blocks[bid] = res
                    continue
                out = prev if prev is not None else DeviceChunk.empty(shape, expr.dtype, self.device)
                st.blocks[bid] = out                          # allocated once: replays refill the same memory
                if out.size:
                    blk = rt.BlockArgs(shape=res.shape, inputs=[(res.ptr, res.strides)], out0=out.ptr)
                    keep = [res]
                    for launch in rt.fused_launches(ident, _lib.RED_NONE, (), [blk]):
                        launch.run()
                        keep.append(launch)
                    out._keep = keep
        self._do(run)
        return st

    def _run_Ravel(self, expr):
        x = expr.operand("array")
        src = self.results[x._name]
        return self._view_store(expr, lambda bid: (src, x, expr.source(bid)), lambda blk, bid: expr.view(blk))

    def _run_Reshape(self, expr):
        """Every output block is the same-rank input block viewed with its own shape (``ReshapeLowered._layer``)."""
        x = expr.operand("array")
        src = self.results[x._name]
        return self._view_store(expr, lambda bid: (src, x, expr.source(bid)), lambda blk, bid: expr.view(blk, bid))

    def _run_Squeeze(self, expr):
        x = expr.operand("array")
        src = self.results[x._name]
        return self._view_store(expr, lambda bid: (src, x, expr.source(bid)), lambda blk, bid: expr.view(blk))

    def _run_BroadcastTo(self, expr):
        x = expr.operand("array")
        src = self.results[x._name]
        return self._view_store(expr, lambda bid: (src, x, expr.source(bid)), lambda blk, bid: expr.view(blk, bid))

    def _run_Concatenate(self, expr):
        arrs = expr.operand("arrays")
        stores = [self.results[a._name] for a in arrs]

        def src_of(bid):
            k, ibid = expr.source(bid)
            return stores[k], arrs[k], ibid
        return self._view_store(expr, src_of, lambda blk, bid: blk)

    # ------------------------------------------------------------------ fused blockwise
    def _run_FusedBlockwise(self, expr: FusedBlockwise):
        # the plan (kernel program + leaf map) is rebuilt per run: it references the leaves -- host arrays,
        # resident device blocks -- which a process-wide cache would pin for ever; the expensive part, the
        # compiled kernel, is cached by structure in _runtime
        plan = FusedPlan(expr)
        red = plan.reduce
        top = plan.eval_expr
        kind = red.operand("kind") if red is not None else None
        store_kind = {"mean": "mean", "var": "moment"}.get(kind, "array")
        st = BlockStore(expr, store_kind)
        deps = [self.results[dep._name] for dep, _ in plan.leaves]
        out_ids = [bid for bid in expr.block_ids() if self.mine(expr, bid)]
        # ---- blocks of dependencies living on other GPUs (e.g. x.T + x across the partition)
        extra = self._exchange_for_fused(plan, deps, out_ids) if self.world.size > 1 else {}
        blocks, block_owner = [], []
        axes = red.operand("axis") if red is not None else ()
        acc_dtype = None
        if red is not None:
            # every partial block of this rank inside ONE slab (the send buffer of the tree's all-gather)
            st.slab, fields = alloc_partials(self, expr, store_kind)
            if kind == "var":
                acc_dtype = np.float32 if plan.program.out_dtype == np.float32 else np.float64
            else:
                acc_dtype = red.dtype
        if kind in ("min", "max", "nanmin", "nanmax") and math.prod(top.shape[a] for a in axes) == 0:
            ufunc = {"min": "minimum", "max": "maximum", "nanmin": "fmin", "nanmax": "fmax"}[kind]
            raise ValueError(f"zero-size array to reduction operation {ufunc} which has no identity")
        for bid in out_ids:
            shape = top.block_shape(bid) if red is not None else expr.block_shape(bid)
            if red is not None:
                f = fields[bid]
                out = f["total"] if kind == "mean" else f[""]
                st.blocks[bid] = {"total": out, "n": math.prod(shape[a] for a in axes)} if kind == "mean" else out
                if math.prod(shape) == 0:
                    # a zero-size block contributes the fold's identity (``chunk_min`` / ``chunk_max`` return an
                    # empty partial that vanishes in the aggregate's concatenate, _common.py:92-105)
                    if out.size:
                        _fill_identity(out, kind)
                    continue
            elif math.prod(shape) == 0:
                st.blocks[bid] = self._empty_result(expr, bid, store_kind)
                continue
            ins = []
            remote_owner = self.world.rank
            for k, (dep, _) in enumerate(plan.leaves):
                lbid = plan.leaf_block_id(k, bid)
                src = deps[k].blocks.get(lbid)
                if src is None:
                    src = extra.get((dep._name, lbid))
                    remote_owner = getattr(getattr(src, "buf", None), "owner", remote_owner)
                if src is None:
                    raise RuntimeError(f"block {lbid} of {dep._name} is not resident on rank {self.world.rank}")
                ins.append((src.ptr, plan.leaf_strides(k, src, len(shape))))
            if red is None:
                block_owner.append(remote_owner)
                out = DeviceChunk.empty(expr.block_shape(bid), expr.dtype, self.device)
                st.blocks[bid] = out
            blocks.append(rt.BlockArgs(shape=shape, inputs=ins, out0=out.ptr))
        bar = extra.get("__barrier__")
        keep_order = False
        if bar is not None:
            self._do(bar, collective=True)         # the owners have produced the blocks this rank reads in place
            if red is None:
                blocks, keep_order = _interleave_remote_reads(
                    blocks, block_owner, self.world.rank, self.world.size,
                    [d.itemsize for d in plan.program.inputs], expr.dtype.itemsize), True
        if blocks:
            for launch in rt.fused_launches(plan.program, REDOPS[kind] if red is not None else _lib.RED_NONE,
                                            axes, blocks, acc_dtype=acc_dtype, keep_order=keep_order):
                self._do(launch.run)
                st.keepalive.append(launch)
            st.keepalive.append(extra)
        if bar is not None:
            self._do(bar, collective=True)         # ... and every reader is done before an owner may reuse them
        return st

    def _empty_result(self, expr, bid, kind):
        shape = expr.block_shape(bid)
        if kind == "moment":
            c = DeviceChunk.empty(tuple(shape) + (3,), np.float64, self.device)
            c.buf.zero_()
            return c
        c = DeviceChunk.empty(shape, expr.dtype, self.device)
        c.buf.zero_()
        return {"total": c, "n": 0} if kind == "mean" else c

    # ------------------------------------------------------------------ arg reductions
    def _run_ArgChunk(self, expr: ArgChunk):
        from . import _codegen as cg

        x = expr.operand("array")
        src = self.results[x._name]
        kind, axis, ravel = expr.operand("kind"), expr.operand("axis"), expr.operand("ravel")
        prog = cg.Program()
        prog.set_output(prog.op("positive", prog.add_input(x.dtype)))
        st = BlockStore(expr, "arg")
        blocks = []
        st.slab, fields = alloc_partials(self, expr, "arg")
        for bid in x.block_ids():
            if not self.mine(x, bid):
                continue
            c = src.blocks[bid]
            if c.size == 0:
                # np.argmax of an empty block raises inside the reference's ``arg_chunk`` too (_common.py:730-779)
                raise ValueError(f"attempt to get {kind} of an empty sequence (block {bid} has shape {c.shape})")
            vals, arg = fields[bid]["vals"], fields[bid]["arg"]
            start = x.block_start(bid)
            kw = {}
            if ravel:
                kw["arg_ravel"] = (c.shape, start, x.shape)
            else:
                kw["arg_offset"] = start[axis[0]]
            blocks.append(rt.BlockArgs(shape=c.shape, inputs=[(c.ptr, c.strides)], out0=vals.ptr, out1=arg.ptr, **kw))
            st.blocks[bid] = {"vals": vals, "arg": arg}
        if blocks:
            for launch in rt.fused_launches(prog, REDOPS[kind], axis, blocks):
                self._do(launch.run)
                st.keepalive.append(launch)
        return st

    # ------------------------------------------------------------------ top-k
    def _run_TopK(self, expr):
        """``topk`` / ``argtopk`` (``routines/_topk.py``): per chain of blocks along the axis, level 0 sorts
        the segments of every block row and keeps their best k (value, global index) candidates; further
        levels do the same on the candidate rows until one segment is left.  One launch per block at level 0,
        one per level afterwards."""
        import ctypes as C

        x = expr.operand("array")
        multi = self.world.size > 1
        src = self.results[x._name]
        k, axis, want_arg = expr.operand("k"), expr.operand("axis"), expr.operand("arg")
        largest, kabs = k > 0, abs(k)
        item = x.dtype.itemsize
        seg_max = 4096 if item <= 4 else 2048
        ntot = x.shape[axis]
        if min(kabs, ntot) > seg_max // 2 and ntot > seg_max:
            raise NotImplementedError(f"topk with k > {seg_max // 2} over an axis longer than {seg_max}")
        code = _lib.dtype_code(x.dtype)
        st = BlockStore(expr)
        nd = x.ndim
        perm = tuple(d for d in range(nd) if d != axis) + (axis,)
        ident = cg.Program()
        ident.set_output(ident.op("astype", ident.add_input(x.dtype), dtype=x.dtype))

        def pow2(v):
            p = 2
            while p < v:
                p *= 2
            return p

        def rows_last(chunk):
            """(rows, n) contiguous view of a block with the axis moved last (copied once if needed)."""
            moved = chunk.transpose(perm)
            if not moved.is_contiguous:
                out = DeviceChunk.empty(moved.shape, moved.dtype, self.device)
                if moved.size:
                    for launch in rt.fused_launches(ident, _lib.RED_NONE, (), [rt.BlockArgs(
                            shape=moved.shape, inputs=[(moved.ptr, moved.strides)], out0=out.ptr)]):
                        self._do(launch.run)
                        st.keepalive.append(launch)
                moved = out
            return moved

        def launch(src_ptr, rows, n, pitch, seg, vals, idx, out_pitch, in_idx, off):
            args = (code, src_ptr, rows, n, pitch, seg, kabs, int(largest), vals, idx, out_pitch, in_idx, off)
            self._do(lambda: _lib.check(_lib.lib.b2_topk_rows(*args, rt.current_stream_ptr())))

        others = [range(nb) for d, nb in enumerate(x.numblocks) if d != axis]
        for cid in itertools.product(*others):
            bids = [cid[:axis] + (i,) + cid[axis:] for i in range(x.numblocks[axis])]
            oshape = tuple(n for d, n in enumerate(x.block_shape(bids[0])) if d != axis)
            rows = math.prod(oshape)
            out_bid = cid[:axis] + (0,) + cid[axis:]
            keep = min(kabs, ntot)
            if rows == 0 or ntot == 0:
                st.blocks[out_bid] = DeviceChunk.empty(expr.block_shape(out_bid), expr.dtype, self.device)
                continue
            # ---- level 0: every block of the chain contributes min(k, seg) candidates per segment
            plan, width = [], 0
            for bid in bids:
                n_i = x.block_shape(bid)[axis]
                if n_i == 0:
                    continue
                seg = min(seg_max, pow2(n_i))
                nseg = -(-n_i // seg)
                kk = min(kabs, seg)
                plan.append((bid, n_i, seg, width))
                width += nseg * kk
            vals = DeviceChunk.empty((rows, width), x.dtype, self.device)
            idx = DeviceChunk.empty((rows, width), np.int64, self.device)
            if multi and not src.replicated:
                # every rank fills the candidate columns of ITS blocks; the rest is zero, so a byte-wise
                # all-reduce(SUM) completes both tables exactly (no arithmetic on the values) and every rank
                # runs the remaining levels redundantly: the result is replicated
                self._do(lambda v=vals, i=idx: (v.buf.zero_(), i.buf.zero_()))
            for bid, n_i, seg, col in plan:
                if multi and not src.replicated and not self.mine(x, bid):
                    continue
                blk = rows_last(src.blocks[bid])
                st.keepalive.append(blk)
                launch(blk.ptr, rows, n_i, n_i, seg, vals.ptr + col * item, idx.ptr + col * 8, width, None,
                       x.block_start(bid)[axis])
            if multi and not src.replicated:
                import torch.distributed as dist

                self._do(lambda v=vals, i=idx: (dist.all_reduce(v.buf), dist.all_reduce(i.buf)), collective=True)
            single = len(plan) == 1 and -(-plan[0][1] // plan[0][2]) == 1
            # ---- further levels on the candidate rows
            while not single:
                seg = min(seg_max, pow2(width))
                nseg = -(-width // seg)
                kk = min(kabs, seg)
                nv = DeviceChunk.empty((rows, nseg * kk), x.dtype, self.device)
                ni = DeviceChunk.empty((rows, nseg * kk), np.int64, self.device)
                launch(vals.ptr, rows, width, width, seg, nv.ptr, ni.ptr, nseg * kk, idx.ptr, 0)
                st.keepalive.extend([vals, idx])
                vals, idx, width = nv, ni, nseg * kk
                single = nseg == 1
            res = idx if want_arg else vals
            # (rows, width) with the best `keep` first -> block shaped like the input with the axis last ...
            strides = []
            acc = width
            for n in reversed(oshape):
                strides.append(acc)
                acc *= max(n, 1)
            view = DeviceChunk(res.buf, oshape + (keep,), res.dtype, strides=tuple(reversed(strides)) + (1,), offset=res.offset)
            inv = [0] * nd
            for pos, d in enumerate(perm):
                inv[d] = pos
            st.blocks[out_bid] = view.transpose(tuple(inv))         # ... and the axis back in its place (a view)
            st.keepalive.extend([vals, idx])
        st.replicated = multi
        return st

    # ------------------------------------------------------------------ cumulative scans
    def _run_CumReduction(self, expr: CumReduction):
        """cumsum / cumprod (``reductions/_cumulative.py:100-265``) in three steps, 3 N bytes:
        (1) per-segment totals with the ordinary reduction kernels (read N);
        (2) the scan of those totals along the axis -- a tiny launch of the scan kernel itself;
        (3) one scan pass writing ``carry (+|*) local_scan(x)`` (read N, write N).
        Segments are the blocks along the axis; a 1-D block is first viewed as rows of ``SEG`` elements
        so that a long vector still fills the GPU.  With several ranks the totals tables are completed
        by an all-reduce (each entry has one owner) and every rank scans them redundantly."""
        x = expr.operand("array")
        src = self.results[x._name]
        kind, axis, nan = expr.operand("kind"), expr.operand("axis"), expr.operand("nan")
        acc = expr.dtype
        if acc.itemsize not in (4, 8) or acc.kind not in "iuf":
            raise NotImplementedError(f"cumulative reduction into dtype {acc} is not supported by the B200 kernels")
        cum = _Cum(self, BlockStore(expr), acc, _lib.RED_SUM if kind == "cumsum" else _lib.RED_PROD)
        prog = cg.Program()
        ref = prog.add_input(x.dtype)
        if nan and x.dtype.kind == "f":
            ref = prog.op("where", prog.op("isnan", ref), prog.typed_const(cum.ident, x.dtype), ref)
        prog.set_output(prog.op("astype", ref, dtype=acc))
        cum.prog = prog
        if x.ndim == 1:
            cum.vector_blocks(x, src)
        else:
            cum.nd_blocks(x, src, axis)
        return cum.st

    # ------------------------------------------------------------------ tree levels
    def _gather_partials(self, expr, src: BlockStore, x):
        """The partial blocks a ``PartialReduce`` level folds on this rank.  Returns ``(blocks, local)``.
        ``local``: every group's members live on the rank that owns the group's output block (e.g.
        ``mean(axis=0)`` with the block columns dealt to the ranks): each owner folds its own groups, no
        exchange at all.  Otherwise the partials are all-gathered (peer memory) and every rank folds the
        whole level, so the results are replicated."""
        if self.world.size == 1 or src.replicated:
            return src.blocks, False
        if level_is_local(expr, self.world.size):
            return src.blocks, True
        return _allgather_blocks(self, src, x), False

    def _run_PartialReduce(self, expr: PartialReduce):
        x = expr.operand("array")
        src = self.results[x._name]
        kind, final = expr.operand("kind"), expr.operand("final")
        parts, local = self._gather_partials(expr, src, x)
        redop = REDOPS[kind]
        st = BlockStore(expr, src.kind if not final else "array", replicated=not local)
        groups, in_dt, out_dt, op = [], None, None, redop
        for key, members in expr.groups():
            if local and not self.mine(expr, key):
                continue
            blks = [parts[m] for m in members]
            first = blks[0]
            if src.kind == "mean":
                tot0 = first["total"]
                n = sum(b["n"] for b in blks)
                out = DeviceChunk.empty(tot0.shape if not final else self._final_shape(expr, tot0.shape),
                                        expr.dtype if final else tot0.dtype, self.device)
                groups.append(dict(parts=[b["total"].ptr for b in blks], nelem=tot0.size, out0=out.ptr,
                                   post=_lib.POST_MEAN if final else _lib.POST_NONE, count=n))
                in_dt, out_dt, op = tot0.dtype, (expr.dtype if final else tot0.dtype), _lib.RED_SUM
                st.blocks[key] = out if final else {"total": out, "n": n}
            elif src.kind == "moment":
                nelem = first.size // 3
                if final:
                    out = DeviceChunk.empty(self._final_shape(expr, first.shape[:-1]), expr.dtype, self.device)
                    groups.append(dict(parts=[b.ptr for b in blks], nelem=nelem, out0=out.ptr, post=_lib.POST_VAR,
                                       ddof=expr.operand("ddof")))
                else:
                    out = DeviceChunk.empty(first.shape, np.float64, self.device)
                    groups.append(dict(parts=[b.ptr for b in blks], nelem=nelem, out0=out.ptr))
                in_dt, out_dt, op = np.dtype(np.float64), (expr.dtype if final else np.dtype(np.float64)), _lib.RED_MOMENT
                st.blocks[key] = out
            elif src.kind == "arg":
                v0 = first["vals"]
                vals = DeviceChunk.empty(v0.shape, v0.dtype, self.device)
                arg = DeviceChunk.empty(self._final_shape(expr, v0.shape) if final else v0.shape, np.int64, self.device)
                groups.append(dict(parts=[b["vals"].ptr for b in blks], parts1=[b["arg"].ptr for b in blks],
                                   nelem=v0.size, out0=vals.ptr, out1=arg.ptr))
                in_dt = out_dt = v0.dtype
                st.blocks[key] = arg if final else {"vals": vals, "arg": arg}
                st.keepalive.append(vals)
            else:
                out = DeviceChunk.empty(self._final_shape(expr, first.shape) if final else first.shape,
                                        first.dtype, self.device)
                groups.append(dict(parts=[b.ptr for b in blks], nelem=first.size, out0=out.ptr))
                in_dt = out_dt = first.dtype
                st.blocks[key] = out
            st.keepalive.append(blks)
        if not groups:                      # this rank owns no output block of the level
            return st
        launch = rt.CombineGroupsLaunch(op, in_dt, out_dt, groups)       # ONE launch for the whole level
        self._do(launch.run)
        st.keepalive.append(launch)
        return st

    @staticmethod
    def _final_shape(expr: PartialReduce, kd_shape):
        if expr.operand("keepdims"):
            return tuple(kd_shape)
        se = expr.operand("split_every")
        return tuple(n for d, n in enumerate(kd_shape) if d not in se)

    # ------------------------------------------------------------------ rechunk
    def _run_TasksRechunk(self, expr: TasksRechunk):
        x = expr.operand("array")
        src = self.results[x._name]
        st = BlockStore(expr)
        item = expr.dtype.itemsize
        if self.world.size > 1 and not src.replicated and _peer.enabled():
            return _rechunk_push(self, expr, src, st)
        new_ids = [bid for bid in expr.block_ids() if self.mine(expr, bid)]
        remote = {}
        if self.world.size > 1 and not src.replicated:
            remote = _exchange_for_rechunk(self, expr, src, new_ids)
        copies = []
        for nbid in new_ids:
            out = DeviceChunk.empty(expr.block_shape(nbid), expr.dtype, self.device)
            st.blocks[nbid] = out
            for obid, sl, dsl in expr.pieces(nbid):
                blk = src.blocks.get(obid)
                if blk is not None:
                    piece = blk[sl]
                else:
                    piece = remote[(obid, nbid)]          # already cut to the piece on the sender
                copies.extend(_copy_descs(piece, out[dsl], item))
        launch = rt.GatherLaunch(copies)
        self._do(launch.run)
        st.keepalive.extend([launch, remote])
        return st

    # ------------------------------------------------------------------ sliding-window reductions
    def _run_WindowHalo(self, expr):
        """Blocks + the ``window - 1`` elements that follow them = the rechunk executor on ``WindowHalo.pieces``.
        On one GPU, a halo along the LAST axis gets blocks whose row pitch is rounded up to 16 bytes: the rows
        stay 16-byte aligned, so the gather keeps its TMA bulk path (an odd pitch drops it to 4-byte copies,
        3.1 ms instead of 1.3 ms for 4 GiB on B200)."""
        x = expr.operand("array")
        src = self.results[x._name]
        ax = expr.operand("axis")
        if self.world.size > 1 or ax != x.ndim - 1 or x.ndim < 2:
            return self._run_TasksRechunk(expr)
        st = BlockStore(expr)
        item = expr.dtype.itemsize
        copies = []
        for nbid in expr.block_ids():
            shape = expr.block_shape(nbid)
            q = max(1, 16 // item)
            pitch = -(-shape[-1] // q) * q
            rows = math.prod(shape[:-1])
            buf = alloc_bytes(max(rows * pitch * item, 1), self.device)
            strides, acc = [1], pitch
            for n in reversed(shape[:-1]):
                strides.append(acc)
                acc *= max(n, 1)
            out = DeviceChunk(buf, shape, expr.dtype, strides=tuple(reversed(strides)))
            st.blocks[nbid] = out
            for obid, sl, dsl in expr.pieces(nbid):
                copies.extend(_copy_descs(src.blocks[obid][sl], out[dsl], item))
        launch = rt.GatherLaunch(copies)
        self._do(launch.run)
        st.keepalive.append(launch)
        return st

    def _run_WindowReduce(self, expr):
        """``b2_window_reduce`` per block (``reductions/_sliding_window.py:96-160``)."""
        x = expr.operand("array")
        src = self.results[x._name]
        w, ax, kd = expr.operand("window"), expr.operand("axis"), expr.operand("keepdims")
        redop, mean = REDOPS[expr.operand("redop")], int(bool(expr.operand("mean")))
        st = BlockStore(expr)
        jobs, along = [], None
        for bid in x.block_ids():
            if not self.mine(x, bid):
                continue
            blk = src.blocks[bid]
            shape = list(blk.shape)
            dense = list(_contig(shape))
            pitch = 0
            if list(blk.strides) != dense:
                # the padded layout of _run_WindowHalo: dense except for the pitch of the last axis
                pitch = blk.strides[-2] if len(shape) >= 2 else 0
                want, acc = [1], pitch
                for n in reversed(shape[:-1]):
                    want.append(acc)
                    acc *= max(n, 1)
                if ax != len(shape) - 1 or list(blk.strides) != list(reversed(want)):
                    raise RuntimeError("window reduction needs contiguous (or last-axis padded) halo blocks")
            n = shape[ax] - (w - 1)
            oshape = shape[:ax] + [max(n, 0)] + shape[ax + 1:]
            out = DeviceChunk.empty(oshape, expr.dtype, self.device)
            obid = tuple(bid)
            if kd:
                wa = expr.operand("window_axis")
                obid = obid[:wa] + (0,) + obid[wa:]
                view = out.reshape(tuple(oshape[:wa]) + (1,) + tuple(oshape[wa:]))
            else:
                view = out
            st.blocks[obid] = view
            if out.size == 0:
                continue
            B = math.prod(shape[:ax])
            Cc = math.prod(shape[ax + 1:])
            job = _lib.WindowJob()
            job.src, job.dst = blk.ptr, out.ptr
            if Cc == 1:                                   # the sliding axis is the contiguous one: (rows, C)
                job.B, job.R, job.C, along_b = 1, B, shape[ax], 1
                job.src_pitch = pitch
            else:
                job.B, job.R, job.C, along_b = B, shape[ax], Cc, 0
            if along is not None and along != along_b:
                raise NotImplementedError("window reduction over blocks of mixed orientation")
            along = along_b
            jobs.append(job)
            st.keepalive.append(blk)
        if jobs:
            import ctypes as C

            arr = (_lib.WindowJob * len(jobs))(*jobs)
            d_jobs = alloc_bytes(C.sizeof(arr), self.device)
            code = _lib.dtype_code(x.dtype)
            self._do(lambda: _lib.check(_lib.lib.b2_window_reduce_batched(
                redop, code, arr, len(jobs), d_jobs.data_ptr(), w, along, mean, rt.current_stream_ptr())))
            st.keepalive.extend([arr, d_jobs])
        return st

    # ------------------------------------------------------------------ blocked matmul / tensordot
    def _run_BlockGEMM(self, expr):
        """One tcgen05 launch per output block; the k blocks (and, for fp32 operands, the six
        bf16 x 3 split products) are accumulated in TMEM (``b2_gemm_tn_pairs``)."""
        a, bt = expr.operand("a"), expr.operand("bt")
        sa, sb = self.results[a._name], self.results[bt._name]
        st = BlockStore(expr)
        nk = a.numblocks[1]
        mine = [bid for bid in expr.block_ids() if self.mine(expr, bid)]
        if self.world.size > 1:
            # the owner of output block (i, j) gathers row-panel i of a and row-panel j of bt;
            # the k-accumulation then stays local (TMEM)
            wanted = []
            for bid in expr.block_ids():
                r = self.world.owner(expr, bid)
                for k in range(nk):
                    if not sa.replicated:
                        wanted.append((r, a, (bid[0], k)))
                    if not sb.replicated:
                        wanted.append((r, bt, (bid[1], k)))
            got = _fetch_blocks(self, wanted, {a._name: sa, bt._name: sb})
            va, vb = BlockStore(a), BlockStore(bt)
            for (i, j) in mine:
                for k in range(nk):
                    blk = sa.blocks.get((i, k))
                    va.blocks[(i, k)] = blk if blk is not None else got[(a._name, (i, k))]
                    blk = sb.blocks.get((j, k))
                    vb.blocks[(j, k)] = blk if blk is not None else got[(bt._name, (j, k))]
            sa, sb = va, vb
            st.keepalive.append(got)
        problems = []
        for (i, j) in mine:
            out = DeviceChunk.empty(expr.block_shape((i, j)), expr.dtype, self.device)
            st.blocks[(i, j)] = out
            problems.append((out, [(sa.blocks[(i, k)], sb.blocks[(j, k)]) for k in range(nk)]))
        self._contract(st, problems, a.dtype)
        return st

    def _run_BlockContract(self, expr):
        """N-d ``tensordot`` (``linalg/_tensordot.py:45-136``): every operand block is matricised once (free
        axes x contracted axes, a copy only when the block is not already in that order), every output
        block is ONE accumulation over the contracted block indices -- the reference's (.., 1, ..) partials
        and their ``.sum(axis=left_axes)`` tree are never materialised."""
        a, b = expr.operand("a"), expr.operand("b")
        la, lb = expr.operand("la"), expr.operand("lb")
        sa, sb = self.results[a._name], self.results[b._name]
        if self.world.size > 1 and not (sa.replicated and sb.replicated):
            raise NotImplementedError("N-d tensordot across several GPUs (2-D matmul shards; gather the operands first)")
        fa = [d for d in range(a.ndim) if d not in la]
        fb = [d for d in range(b.ndim) if d not in lb]
        st = BlockStore(expr)
        ident = {}

        def matricize(store, x, bid, free, con, cache):
            if bid not in cache:
                blk = store.blocks[bid]
                moved = blk.transpose(tuple(free) + tuple(con))
                if not moved.is_contiguous or moved.ptr % 16:
                    out = DeviceChunk.empty(moved.shape, moved.dtype, self.device)
                    if moved.size:
                        prog = ident.get(x.dtype)
                        if prog is None:
                            prog = ident[x.dtype] = cg.Program()
                            prog.set_output(prog.op("astype", prog.add_input(x.dtype), dtype=x.dtype))
                        for launch in rt.fused_launches(prog, _lib.RED_NONE, (), [rt.BlockArgs(
                                shape=moved.shape, inputs=[(moved.ptr, moved.strides)], out0=out.ptr)]):
                            self._do(launch.run)
                            st.keepalive.append(launch)
                    moved = out
                m = math.prod(blk.shape[d] for d in free)
                k = math.prod(blk.shape[d] for d in con)
                cache[bid] = moved.reshape((m, k))
            return cache[bid]

        ca, cb = {}, {}
        kranges = [range(a.numblocks[d]) for d in la]
        problems = []
        for bid in expr.block_ids():
            if not self.mine(expr, bid):
                continue
            ia, jb = bid[:len(fa)], bid[len(fa):]
            pairs = []
            for kk in itertools.product(*kranges):
                abid = [0] * a.ndim
                for d, i in zip(fa, ia):
                    abid[d] = i
                for d, i in zip(la, kk):
                    abid[d] = i
                bbid = [0] * b.ndim
                for d, j in zip(fb, jb):
                    bbid[d] = j
                for d, i in zip(lb, kk):
                    bbid[d] = i
                pairs.append((matricize(sa, a, tuple(abid), fa, la, ca), matricize(sb, b, tuple(bbid), fb, lb, cb)))
            shape = expr.block_shape(bid)
            out = DeviceChunk.empty(shape, expr.dtype, self.device)
            st.blocks[bid] = out
            m = math.prod(shape[:len(fa)])
            problems.append((out.reshape((m, math.prod(shape[len(fa):]))), pairs))
        self._contract(st, problems, np.result_type(a.dtype, b.dtype) if a.dtype != b.dtype else a.dtype)
        st.keepalive.extend([ca, cb])
        return st

    def _contract(self, st, problems, in_dtype):
        """``out (M, N) = sum_p A_p (M, K_p) @ B_p (N, K_p)^T`` for every problem ``(out, [(A_p, B_p)])``.
        bf16 / fp32 operands whose contraction lengths are multiples of 8: ONE batched tcgen05 launch chain
        (fp32 through the bf16 x 3 split); everything else: the exact SIMT kernel, pair by pair."""
        import ctypes as C
        import os

        if not problems:
            return
        in_dtype = np.dtype(in_dtype)
        name = in_dtype.name
        tensor = name in ("float32", "bfloat16") and all(
            A.shape[1] % 8 == 0 and A.ptr % 16 == 0 and B.ptr % 16 == 0 and A.is_contiguous and B.is_contiguous
            and A.dtype == in_dtype and B.dtype == in_dtype for _, pairs in problems for A, B in pairs)
        if not tensor:
            if name == "bfloat16":
                raise NotImplementedError(
                    "bfloat16 contraction chunks must be multiples of 8 elements (16-byte TMA row strides): rechunk the contracted axis")
            return self._contract_exact(st, problems)
        fp32 = name == "float32"
        bf16 = np.dtype("uint16")          # raw 16-bit planes
        planes = {}

        def split(blk):
            key = (blk.ptr, blk.shape)
            if key not in planes:
                if not fp32:
                    planes[key] = (blk,)
                else:
                    hi, mid, lo = (DeviceChunk.empty(blk.shape, bf16, self.device) for _ in range(3))
                    self._do(lambda b=blk, h=hi, m=mid, l=lo: _lib.check(_lib.lib.b2_split3_bf16(
                        b.ptr, h.ptr, m.ptr, l.ptr, b.size, rt.current_stream_ptr())))
                    planes[key] = (hi, mid, lo)
            return planes[key]

        # products kept for fp32: hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi  (error ~2^-24)
        combos = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)] if fp32 else [(0, 0)]
        probs, keep = [], []
        for out, pairs in problems:
            M, N = out.shape
            if M == 0 or N == 0:
                continue
            pairs = [(A, B) for A, B in pairs if A.shape[1] > 0]
            if not pairs:
                self._do(lambda o=out: rt.fill(o, 0))
                continue
            ks = [A.shape[1] for A, _ in pairs]
            K = max(ks)
            Ap, Bp, Kp = [], [], []
            for (A, B), kk in zip(pairs, ks):
                pa, pb = split(A), split(B)
                for ca_, cb_ in combos:
                    Ap.append(pa[ca_].ptr)
                    Bp.append(pb[cb_].ptr)
                    Kp.append(kk)
            arrA = (C.c_void_p * len(Ap))(*Ap)
            arrB = (C.c_void_p * len(Bp))(*Bp)
            arrK = (C.c_int64 * len(Kp))(*Kp)
            keep.extend([arrA, arrB, arrK])
            p = _lib.GemmProblem()
            p.A, p.B = C.cast(arrA, C.c_void_p), C.cast(arrB, C.c_void_p)
            p.Kpair = C.cast(arrK, C.c_void_p).value if len(set(ks)) > 1 else None
            p.npairs, p.accumulate, p.lda, p.ldb = len(Ap), 0, K, K
            p.C, p.ldc, p.M, p.N, p.K = out.ptr, N, M, N, K
            probs.append(p)
        if not probs:
            return
        if os.environ.get("B2_GEMM_BATCHED", "1") == "0" and all(not p.Kpair for p in probs):   # diagnostic: one launch per output block
            for p in probs:
                self._do(lambda p=p: _lib.check(_lib.lib.b2_gemm_tn_pairs(
                    _lib.dtype_code("bfloat16"), p.A, p.B, p.npairs, p.lda, p.ldb, p.C, p.ldc, p.M, p.N, p.K, 0,
                    rt.current_stream_ptr())))
            st.keepalive.extend([keep, probs, planes])
            return
        # Long pair lists are issued as several batched launches that accumulate into C.  Same-box
        # A/B (bench --config c5, fp32 split = 48 pairs): 48 pairs/launch 693, 24 -> 763, 12 -> 993,
        # 6 -> 1058 tensor TFLOP/s: in a long launch the CTAs drift out of lockstep and stop sharing
        # operand tiles in L2; a kernel boundary re-aligns them.  (Cluster multicast is the real fix.)
        longest = max(p.npairs for p in probs)
        CH = int(os.environ.get("B2_GEMM_CHUNK", "12" if longest <= 12 else "6"))
        nchunks = max(-(-p.npairs // CH) for p in probs)
        for ci in range(nchunks):
            sub = []
            for p in probs:
                lo, hi = ci * CH, min((ci + 1) * CH, p.npairs)
                if lo >= hi:
                    continue
                q = _lib.GemmProblem()
                q.A, q.B = p.A + 8 * lo, p.B + 8 * lo
                q.npairs, q.accumulate, q.lda, q.ldb = hi - lo, 1 if ci > 0 else 0, p.lda, p.ldb
                q.C, q.ldc, q.M, q.N, q.K = p.C, p.ldc, p.M, p.N, p.K
                q.Kpair = (p.Kpair + 8 * lo) if p.Kpair else None
                sub.append(q)
            arr = (_lib.GemmProblem * len(sub))(*sub)
            need = C.c_size_t()
            _lib.check(_lib.lib.b2_gemm_tn_batched(_lib.dtype_code("bfloat16"), arr, len(sub), None, 0, C.byref(need), None))
            ws = alloc_bytes(need.value + 64, self.device)
            wptr = (ws.data_ptr() + 63) // 64 * 64
            self._do(lambda arr=arr, n=len(sub), wptr=wptr, nb=need.value: _lib.check(_lib.lib.b2_gemm_tn_batched(
                _lib.dtype_code("bfloat16"), arr, n, wptr, nb, None, rt.current_stream_ptr())))
            st.keepalive.extend([arr, ws])
        st.keepalive.extend([keep, planes])

    def _contract_exact(self, st, problems):
        """The coverage path of ``_contract``: ``b2_gemm_tn_simt`` per pair, accumulating in the element type
        (fp64 FMA, IEEE fp32, wrap-around integers -- ``np.matmul`` semantics per block)."""
        for out, pairs in problems:
            M, N = out.shape
            if M == 0 or N == 0:
                continue
            dt = out.dtype
            if dt.name not in ("float64", "float32", "int32", "uint32", "int64", "uint64"):
                raise NotImplementedError(f"matmul / tensordot with result dtype {dt} has no B200 kernel")
            first = True
            for A, B in pairs:
                K = A.shape[1]
                if K == 0:
                    continue
                if A.dtype != dt or B.dtype != dt or not A.is_contiguous or not B.is_contiguous:
                    raise NotImplementedError("exact contraction needs contiguous operands of the result dtype "
                                              "(tensordot() casts at the expression level)")
                code, acc = _lib.dtype_code(dt), 0 if first else 1
                self._do(lambda A=A, B=B, K=K, code=code, acc=acc, out=out, M=M, N=N: _lib.check(_lib.lib.b2_gemm_tn_simt(
                    code, A.ptr, K, B.ptr, K, out.ptr, N, M, N, K, acc, rt.current_stream_ptr())))
                st.keepalive.extend([A, B])
                first = False
            if first:
                self._do(lambda o=out: rt.fill(o, 0))

    # ------------------------------------------------------------------ multi-GPU exchange (fused)
    def _exchange_for_fused(self, plan, deps, out_ids):
        return _exchange_for_fused(self, plan, deps, out_ids)


# the multi-GPU plumbing and the cumulative-scan launch builder live in their own modules
from ._exchange import (  # noqa: E402,F401
    alloc_partials, level_is_local, mean_count, partials_layout, _allgather_blocks, _copy_descs, _exchange_for_fused, _exchange_for_rechunk, _fetch_blocks, _interleave_remote_reads,
    _peer_reads_for_fused, _push_views, _rechunk_push, gather_to_host, owner_of, plan_block_fetch, plan_fused_exchange,
    plan_fused_peer_reads, plan_rechunk_exchange, plan_rechunk_push,
)
from ._cumexec import _Cum  # noqa: E402,F401
