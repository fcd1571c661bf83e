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
from ._device import DeviceChunk, alloc_bytes
from ._expr import ArrayExpr, BroadcastTrick, FromArray, HostBlocks, Random, Resident
from ._rechunk import TasksRechunk
from ._reductions import REDOPS, ArgChunk, ChunkReduce, CumReduction, PartialReduce
from ._slicing import SliceSlicesIntegers


class BlockStore:
    """Blocks of one expression held by this rank.  ``kind``: array | mean | moment | arg."""

    def __init__(self, expr, kind="array", replicated=False):
        self.expr = expr
        self.kind = kind
        self.replicated = replicated
        self.blocks = {}          # bid -> DeviceChunk | dict
        self.keepalive = []       # launch tables / pointer tables the stream may still read


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


_PLAN_CACHE: dict = {}


class Executor:
    def __init__(self, world: World | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("dask_array_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.world = world or World()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.results: dict[str, BlockStore] = {}
        self.tape = []            # every device action in issue order: replaying it repeats the step

    def _do(self, fn):
        fn()
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
        store = fn(expr)
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
            out = DeviceChunk.empty(expr.block_shape(bid), expr.dtype, self.device)
            st.blocks[bid] = out
            ins = []
            for s_ in stores:
                blk = s_.blocks.get(bid)
                if blk is None:
                    raise RuntimeError(f"map_blocks: block {bid} of an argument is not resident on rank {self.world.rank}")
                ins.append(blk)
            work.append((out, ins))
        item = expr.dtype.itemsize

        def run():
            for out, ins in work:
                it = iter(ins)
                args = [next(it) if t is None else t for t in template]
                res = func(*args, **kwargs)
                if not isinstance(res, DeviceChunk):
                    raise TypeError(f"map_blocks: {getattr(func, '__name__', func)!r} returned {type(res).__name__}, "
                                    "not a DeviceChunk (no host fallback: use NumPy functions the chunk type implements)")
                if res.shape != out.shape or res.dtype != out.dtype:
                    raise ValueError(f"map_blocks: block result {res.shape} {res.dtype} does not match the declared "
                                     f"chunks / dtype {out.shape} {out.dtype}")
                if out.size:
                    g = rt.GatherLaunch(_copy_descs(res if res.ndim else res.reshape((1,)),
                                                    out if out.ndim else out.reshape((1,)), item))
                    g.run()
                    out._keep = (g, res)
        self._do(run)
        return st

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
        key = expr._name
        plan = _PLAN_CACHE.get(key)
        if plan is None:
            plan = _PLAN_CACHE[key] = FusedPlan(expr)
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
        for bid in out_ids:
            shape = top.block_shape(bid) if red is not None else expr.block_shape(bid)
            if math.prod(shape) == 0:
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
            out_shape = expr.block_shape(bid)
            if red is None:
                out = DeviceChunk.empty(out_shape, expr.dtype, self.device)
                blocks.append(rt.BlockArgs(shape=shape, inputs=ins, out0=out.ptr))
                st.blocks[bid] = out
            elif kind == "var":
                out = DeviceChunk.empty(tuple(out_shape) + (3,), np.float64, self.device)
                acc_dtype = np.float32 if plan.program.out_dtype == np.float32 else np.float64
                blocks.append(rt.BlockArgs(shape=shape, inputs=ins, out0=out.ptr))
                st.blocks[bid] = out
            elif kind == "mean":
                out = DeviceChunk.empty(out_shape, red.dtype, self.device)
                acc_dtype = red.dtype
                blocks.append(rt.BlockArgs(shape=shape, inputs=ins, out0=out.ptr))
                n = math.prod(shape[a] for a in axes)
                st.blocks[bid] = {"total": out, "n": n}
            else:
                out = DeviceChunk.empty(out_shape, red.dtype, self.device)
                acc_dtype = red.dtype
                blocks.append(rt.BlockArgs(shape=shape, inputs=ins, out0=out.ptr))
                st.blocks[bid] = out
        bar = extra.get("__barrier__")
        keep_order = False
        if bar is not None:
            self._do(bar)         # the owners have produced the blocks this rank reads in place
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
            self._do(bar)         # ... and every reader is done before an owner may reuse them
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
        for bid in x.block_ids():
            if not self.mine(x, bid):
                continue
            c = src.blocks[bid]
            oshape = expr.block_shape(bid)
            vals = DeviceChunk.empty(oshape, x.dtype, self.device)
            arg = DeviceChunk.empty(oshape, np.int64, self.device)
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

        if self.world.size > 1:
            raise NotImplementedError("topk across several GPUs (gather the array on one rank first)")
        x = expr.operand("array")
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
            for bid, n_i, seg, col in plan:
                blk = rows_last(src.blocks[bid])
                st.keepalive.append(blk)
                launch(blk.ptr, rows, n_i, n_i, seg, vals.ptr + col * item, idx.ptr + col * 8, width, None,
                       x.block_start(bid)[axis])
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
    def _gather_partials(self, src: BlockStore, x):
        """Make every partial block of ``x`` available on this rank (all-gather over NCCL)."""
        if self.world.size == 1 or src.replicated:
            return src.blocks
        return _allgather_blocks(self, src, x)

    def _run_PartialReduce(self, expr: PartialReduce):
        x = expr.operand("array")
        src = self.results[x._name]
        kind, final = expr.operand("kind"), expr.operand("final")
        parts = self._gather_partials(src, x)
        redop = REDOPS[kind]
        st = BlockStore(expr, src.kind if not final else "array", replicated=True)
        groups, in_dt, out_dt, op = [], None, None, redop
        for key, members in expr.groups():
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

    # ------------------------------------------------------------------ blocked matmul
    def _run_BlockGEMM(self, expr):
        """One tcgen05 launch per output block; the k blocks (and, for fp32 operands, the six
        bf16 x 3 split products) are accumulated in TMEM (``b2_gemm_tn_pairs``)."""
        import ctypes as C

        a, bt = expr.operand("a"), expr.operand("bt")
        sa, sb = self.results[a._name], self.results[bt._name]
        st = BlockStore(expr)
        nk_ = a.numblocks[1]
        mine = [bid for bid in expr.block_ids() if self.mine(expr, bid)]
        if self.world.size > 1:
            # the owner of output block (i, j) gathers row-panel i of a and row-panel j of bt;
            # the k-accumulation then stays local (TMEM)
            wanted = []
            for bid in expr.block_ids():
                r = self.world.owner(expr, bid)
                for k in range(nk_):
                    if not sa.replicated:
                        wanted.append((r, a, (bid[0], k)))
                    if not sb.replicated:
                        wanted.append((r, bt, (bid[1], k)))
            got = _fetch_blocks(self, wanted, {a._name: sa, bt._name: sb})
            va, vb = BlockStore(a), BlockStore(bt)
            for (i, j) in mine:
                for k in range(nk_):
                    blk = sa.blocks.get((i, k))
                    va.blocks[(i, k)] = blk if blk is not None else got[(a._name, (i, k))]
                    blk = sb.blocks.get((j, k))
                    vb.blocks[(j, k)] = blk if blk is not None else got[(bt._name, (j, k))]
            sa, sb = va, vb
            st.keepalive.append(got)
        fp32 = a.dtype == np.float32
        bf16 = np.dtype("uint16")          # raw 16-bit planes

        def planes(store, x):
            out = {}
            for bid, blk in store.blocks.items():
                if not blk.is_contiguous:
                    raise NotImplementedError("matmul operand blocks must be contiguous (persist / rechunk first)")
                if not fp32:
                    out[bid] = (blk,)
                    continue
                hi, mid, lo = (DeviceChunk.empty(blk.shape, bf16, self.device) for _ in range(3))
                self._do(lambda b=blk, h=hi, m=mid, l=lo: _lib.check(_lib.lib.b2_split3_bf16(
                    b.ptr, h.ptr, m.ptr, l.ptr, b.size, rt.current_stream_ptr())))
                out[bid] = (hi, mid, lo)
            return out

        pa, pb = planes(sa, a), planes(sb, bt)
        # products kept for fp32: hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi  (error ~2^-24)
        combos = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)] if fp32 else [(0, 0)]
        nk = a.numblocks[1]
        probs, keep = [], []
        if not mine:
            return st
        for (i, j) in mine:
            M, N = expr.block_shape((i, j))
            out = DeviceChunk.empty((M, N), np.float32, self.device)
            st.blocks[(i, j)] = out
            ks = [a.block_shape((i, k))[1] for k in range(nk)]
            if any(kk % 8 for kk in ks):
                raise NotImplementedError(
                    f"matmul contraction chunks {tuple(ks)} must be multiples of 8 elements (16-byte TMA row strides): rechunk the contracted axis")
            K = max(ks)
            A, B, Kp = [], [], []
            for k in range(nk):
                for ca, cb in combos:
                    A.append(pa[(i, k)][ca].ptr)
                    B.append(pb[(j, k)][cb].ptr)
                    Kp.append(ks[k])
            arrA = (C.c_void_p * len(A))(*A)
            arrB = (C.c_void_p * len(B))(*B)
            arrK = (C.c_int64 * len(Kp))(*Kp)
            keep.extend([arrA, arrB, arrK])
            p = _lib.GemmProblem()
            p.A, p.B = C.cast(arrA, C.c_void_p), C.cast(arrB, C.c_void_p)
            p.Kpair = C.cast(arrK, C.c_void_p).value if len(set(ks)) > 1 else None
            p.npairs, p.accumulate, p.lda, p.ldb = len(A), 0, K, K
            p.C, p.ldc, p.M, p.N, p.K = out.ptr, N, M, N, K
            probs.append(p)
        import os
        if os.environ.get("B2_GEMM_BATCHED", "1") == "0" and all(not p.Kpair for p in probs):   # diagnostic: one launch per output block
            for p in probs:
                self._do(lambda p=p: _lib.check(_lib.lib.b2_gemm_tn_pairs(
                    _lib.dtype_code("bfloat16"), p.A, p.B, p.npairs, p.lda, p.ldb, p.C, p.ldc, p.M, p.N, p.K, 0,
                    rt.current_stream_ptr())))
            st.keepalive.extend([keep, probs, pa, pb])
            return st
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
        st.keepalive.append(keep)
        st.keepalive.extend([pa, pb])
        return st

    # ------------------------------------------------------------------ multi-GPU exchange (fused)
    def _exchange_for_fused(self, plan, deps, out_ids):
        return _exchange_for_fused(self, plan, deps, out_ids)


class _Cum:
    """Launch builder of one cumulative reduction (see ``Executor._run_CumReduction``)."""

    SEG = 4096            # elements per virtual row of a 1-D block (16-32 KiB: one warp's worth of work)

    def __init__(self, ex: Executor, st: BlockStore, acc, redop):
        self.ex, self.st, self.acc, self.redop = ex, st, np.dtype(acc), redop
        self.ident = 0 if redop == _lib.RED_SUM else 1
        self.prog = None
        self.same = cg.Program()                   # identity chain on accumulator-typed tables
        self.same.set_output(self.same.op("positive", self.same.add_input(self.acc)))

    # ---- launches
    def _run(self, launches):
        for launch in launches:
            self.ex._do(launch.run)
            self.st.keepalive.append(launch)

    def totals(self, prog, blocks, axis):
        """reduce every block along ``axis`` into its ``out0`` (a row of a totals table)."""
        for group in self._by_alignment(blocks):
            self._run(rt.fused_launches(prog, self.redop, (axis,), group, acc_dtype=self.acc))

    def scan(self, prog, blocks, axis):
        for group in self._by_alignment(blocks):
            self._run(rt.scan_launches(prog, self.redop, axis, group, self.acc))

    @staticmethod
    def _by_alignment(blocks):
        """Ragged remainder rows go into their own launch so the full rows keep 16-byte vectors."""
        good = [b for b in blocks if b.shape[-1] % 4 == 0]
        odd = [b for b in blocks if b.shape[-1] % 4]
        return [g for g in (good, odd) if g]

    def table(self, shape):
        t = DeviceChunk(alloc_bytes(math.prod(shape) * self.acc.itemsize, self.ex.device, zero=True), shape, self.acc)
        if self.ex.world.size > 1:
            self.ex._do(lambda: t.buf.zero_())     # rows of other ranks must be zero before the all-reduce
        self.st.keepalive.append(t)
        return t

    def fill_identity(self, chunk):
        if self.ident != 0 or self.ex.world.size > 1:
            self.ex._do(lambda: rt.fill(chunk, self.ident))

    # ---- N-d blocks: the blocks along the axis are the segments, one totals table per chain
    def nd_blocks(self, x, src, axis):
        ex, st, acc = self.ex, self.st, self.acc
        nax = x.numblocks[axis]
        others = [range(n) for d, n in enumerate(x.numblocks) if d != axis]
        tables, tot, main = [], [], []
        for cid in itertools.product(*others):
            bids = [cid[:axis] + (i,) + cid[axis:] for i in range(nax)]
            kshape = tuple(n for d, n in enumerate(x.block_shape(bids[0])) if d != axis)
            ksize = math.prod(kshape)
            tab = self.table((nax,) + kshape) if nax > 1 and ksize else None
            if tab is not None:
                tables.append(tab)
            for i, bid in enumerate(bids):
                row = tab[i] if tab is not None else None
                if x.block_shape(bid)[axis] == 0:
                    if row is not None and (ex.mine(x, bid) or ex.world.size == 1) and self.ident != 0:
                        self.fill_identity(row)    # an empty block carries the identity (_cum_tail :28-39)
                    if ex.mine(x, bid):
                        st.blocks[bid] = DeviceChunk.empty(x.block_shape(bid), acc, ex.device)
                    continue
                if not ex.mine(x, bid):
                    continue
                c = src.blocks[bid]
                out = DeviceChunk.empty(c.shape, acc, ex.device)
                st.blocks[bid] = out
                if c.size == 0:
                    continue
                if row is not None and i < nax - 1:
                    tot.append(rt.BlockArgs(shape=c.shape, inputs=[(c.ptr, c.strides)], out0=row.ptr))
                carry = tab[i - 1].ptr if (tab is not None and i > 0) else 0
                main.append(rt.BlockArgs(shape=c.shape, inputs=[(c.ptr, c.strides)], out0=out.ptr, out1=carry))
        if tot:
            self.totals(self.prog, tot, axis)
        if tables:
            if ex.world.size > 1:
                _sum_tables(ex, tables, acc)
            # in place: row i becomes the total of blocks 0..i = the carry of block i + 1
            self.scan(self.same, [rt.BlockArgs(shape=t.shape, inputs=[(t.ptr, t.strides)], out0=t.ptr)
                                  for t in tables], 0)
        if main:
            self.scan(self.prog, main, axis)

    # ---- 1-D blocks: rows of SEG elements are the segments
    def vector_blocks(self, x, src):
        ex, st, acc, SEG = self.ex, self.st, self.acc, self.SEG
        item = x.dtype.itemsize
        rows, nseg = [], 0           # (block id, first element, row length, number of rows, first segment)
        for bid in x.block_ids():
            full, rem = divmod(x.block_shape(bid)[0], SEG)
            for first, length, n in ([(0, SEG, full)] if full else []) + ([(full * SEG, rem, 1)] if rem else []):
                rows.append((bid, first, length, n, nseg))
                nseg += n
        for bid in x.block_ids():
            if ex.mine(x, bid):
                st.blocks[bid] = DeviceChunk.empty(x.block_shape(bid), acc, ex.device)
        if nseg == 0:
            return
        # carries[k] = total of the segments before k (carries[0] = identity)
        carries = self.table((nseg + 1,)) if nseg > 1 else None
        tot, main = [], []
        for bid, first, length, n, seg0 in rows:
            if not ex.mine(x, bid):
                continue
            c, out = src.blocks[bid], st.blocks[bid]
            s0 = c.strides[0]
            view = [(c.ptr + first * s0 * item, (length * s0, s0))]
            if carries is not None:
                tot.append(rt.BlockArgs(shape=(n, length), inputs=view, out0=carries[seg0 + 1:].ptr))
            main.append(rt.BlockArgs(shape=(n, length), inputs=view, out0=out.ptr + first * acc.itemsize,
                                     out1=carries[seg0:].ptr if carries is not None else 0))
        if carries is not None:
            if tot:
                self.totals(self.prog, tot, 1)
            if ex.world.size > 1:
                _sum_tables(ex, [carries], acc)
            self.fill_identity(carries[0:1])
            self.scan_vector(carries[1:])
        if main:
            self.scan(self.prog, main, 1)

    def scan_vector(self, vec: DeviceChunk):
        """In-place inclusive scan of a contiguous accumulator-typed vector (segment totals)."""
        n, SEG, acc = vec.shape[0], self.SEG, self.acc
        if n <= 16 * SEG:            # one warp walks it
            self.scan(self.same, [rt.BlockArgs(shape=(1, n), inputs=[(vec.ptr, (n, 1))], out0=vec.ptr)], 1)
            return
        full, rem = divmod(n, SEG)
        nseg = full + (1 if rem else 0)
        carries = DeviceChunk(alloc_bytes((nseg + 1) * acc.itemsize, self.ex.device, zero=True), (nseg + 1,), acc)
        self.st.keepalive.append(carries)
        if self.ident != 0:
            self.ex._do(lambda: rt.fill(carries[0:1], self.ident))
        tot, main = [], []
        for first, length, rows, seg0 in [(0, SEG, full, 0)] + ([(full * SEG, rem, 1, full)] if rem else []):
            view = [(vec.ptr + first * acc.itemsize, (length, 1))]
            tot.append(rt.BlockArgs(shape=(rows, length), inputs=view, out0=carries[seg0 + 1:].ptr))
            main.append(rt.BlockArgs(shape=(rows, length), inputs=view, out0=vec.ptr + first * acc.itemsize,
                                     out1=carries[seg0:].ptr))
        self.totals(self.same, tot, 1)
        self.scan_vector(carries[1:])
        self.scan(self.same, main, 1)


def _sum_tables(ex: Executor, tables, acc):
    """Totals tables are zero where another rank owns the block: an all-reduce(SUM) completes them on
    every rank (exact: every entry has exactly one non-zero contribution)."""
    import torch.distributed as dist

    for t in tables:
        n = t.size
        if not n:
            continue
        tdt = {4: torch.int32, 8: torch.int64}[acc.itemsize] if acc.kind in "iu" else \
            {4: torch.float32, 8: torch.float64}[acc.itemsize]
        view = t.buf[t.offset * acc.itemsize: (t.offset + n) * acc.itemsize].view(tdt)
        ex._do(lambda v=view: dist.all_reduce(v))


def _copy_descs(src: DeviceChunk, dst: DeviceChunk, item: int):
    """2-D copy rectangles moving ``src`` into ``dst`` (same shape, arbitrary strides with a
    unit-stride innermost run)."""
    shape = [n for n in src.shape]
    if math.prod(shape) == 0:
        return []
    dims = [(n, s, d) for n, s, d in zip(shape, src.strides, dst.strides) if n != 1]
    if not dims:
        return [(src.ptr, dst.ptr, 1, item, item, item)]
    merged = []
    for n, s, d in dims:
        if merged and merged[-1][1] == s * n and merged[-1][2] == d * n:
            merged[-1] = (merged[-1][0] * n, s, d)
        else:
            merged.append((n, s, d))
    if merged[-1][1] != 1 or merged[-1][2] != 1:
        merged.append((1, 1, 1))        # e.g. a single column: rows of one element each
    n_in, s_in, d_in = merged[-1]
    outer = merged[:-1]
    if not outer:
        return [(src.ptr, dst.ptr, 1, n_in * item, n_in * item, n_in * item)]
    rows, s_row, d_row = outer[-1]
    lead = outer[:-1]
    out = []
    for idx in itertools.product(*[range(n) for n, _, _ in lead]):
        so = sum(i * s for i, (_, s, _) in zip(idx, lead))
        do = sum(i * d for i, (_, _, d) in zip(idx, lead))
        out.append((src.ptr + so * item, dst.ptr + do * item, rows, n_in * item, s_row * item, d_row * item))
    return out


# ----------------------------------------------------------------------------- NCCL plumbing
def _allgather_blocks(ex: Executor, src: BlockStore, x):
    """All-gather the (tiny) per-block partials of ``x`` so every rank can fold the tree."""
    import torch.distributed as dist

    W, me = ex.world.size, ex.world.rank
    ids = list(x.block_ids())

    def fields(b):
        if isinstance(b, dict):
            return [(k, v) for k, v in sorted(b.items()) if isinstance(v, DeviceChunk)]
        return [("", b)]

    # layout is derivable on every rank from shapes alone
    def proto(bid):
        kshape = x.block_shape(bid)
        if src.kind == "moment":
            return [("", tuple(kshape) + (3,), np.dtype(np.float64))]
        if src.kind == "mean":
            return [("total", kshape, x.dtype)]
        if src.kind == "arg":
            vdt = x.operand("array").dtype
            return [("arg", kshape, np.dtype(np.int64)), ("vals", kshape, vdt)]
        return [("", kshape, x.dtype)]

    def nbytes(bid):
        return sum(-(-math.prod(s) * d.itemsize // 16) * 16 for _, s, d in proto(bid))

    per_rank = [sum(nbytes(b) for b in ids if ex.world.owner(x, b) == r) for r in range(W)]
    cap = max(max(per_rank), 16)
    send = alloc_bytes(cap, ex.device)
    off = 0
    copies = []
    for bid in ids:
        if ex.world.owner(x, bid) != me:
            continue
        blk = src.blocks[bid]
        for (name, chunk), (_, shp, dt) in zip(fields(blk), proto(bid)):
            nb = math.prod(shp) * dt.itemsize
            if nb:
                copies.append((chunk.ptr, send.data_ptr() + off, 1, nb, nb, nb))
            off += -(-nb // 16) * 16
    g = rt.GatherLaunch(copies)
    ex._do(g.run)
    recv = alloc_bytes(cap * W, ex.device)
    ex._do(lambda: dist.all_gather_into_tensor(recv, send))
    out = {}
    offs = [0] * W
    for bid in ids:
        r = ex.world.owner(x, bid)
        parts = {}
        for name, shp, dt in proto(bid):
            nb = math.prod(shp) * dt.itemsize
            parts[name] = DeviceChunk(recv, shp, dt, offset=(r * cap + offs[r]) // dt.itemsize)
            offs[r] += -(-nb // 16) * 16
        if src.kind == "mean":
            red = x.root if isinstance(x, FusedBlockwise) else x        # the ChunkReduce
            axes = red.operand("axis")
            top = red.operand("array")
            n = math.prod(top.block_shape(bid)[a] for a in axes)
            out[bid] = {"total": parts["total"], "n": n}
        elif src.kind == "arg":
            out[bid] = parts
        else:
            out[bid] = parts[""]
    src.keepalive.extend([send, recv, g])
    return out


def _p2p_exchange(ex: Executor, sends, recvs):
    """sends: [(peer, tensor)], recvs: [(peer, tensor)] in a globally consistent order."""
    import torch.distributed as dist

    ops = [dist.P2POp(dist.isend, t, p) for p, t in sends] + [dist.P2POp(dist.irecv, t, p) for p, t in recvs]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def owner_of(expr, bid, world_size: int) -> int:
    """Block-cyclic placement: ``ravel(block id) mod world``."""
    if world_size == 1:
        return 0
    nb = expr.numblocks
    return (int(np.ravel_multi_index(bid, nb)) if nb else 0) % world_size


def plan_fused_exchange(plan: FusedPlan, replicated, W: int, me: int):
    """Pure (no device) schedule of the blocks a fused expression reads across the partition.
    Returns (send_items, recv_items): per peer, lists in one global canonical order --
    every rank derives the same order from the expression metadata, so sends and receives
    pair up without negotiation.  Items: (leaf index k, leaf block id, nbytes)."""
    expr = plan.fused
    wanted = {}
    for bid in expr.block_ids():
        r = owner_of(expr, bid, W)
        for k, (dep, _) in enumerate(plan.leaves):
            if replicated[k]:
                continue
            lbid = plan.leaf_block_id(k, bid)
            o = owner_of(dep, lbid, W)
            if o != r:
                wanted[(r, dep._name, lbid)] = (o, k)
    send_items = {p: [] for p in range(W)}
    recv_items = {p: [] for p in range(W)}
    for (r, name, lbid) in sorted(wanted):
        o, k = wanted[(r, name, lbid)]
        dep = plan.leaves[k][0]
        nb = math.prod(dep.block_shape(lbid)) * dep.dtype.itemsize
        if o == me:
            send_items[r].append((k, lbid, nb))
        if r == me:
            recv_items[o].append((k, lbid, nb))
    return send_items, recv_items


def plan_block_fetch(wanted, W: int, me: int):
    """Pure schedule for whole-block reads across the partition.  ``wanted``: iterable of
    (reader rank, dep expr, block id) over ALL ranks (every rank computes the same list).
    Returns (send_items, recv_items) per peer: (dep expr, block id, nbytes), canonical order."""
    uniq = {}
    for r, dep, bid in wanted:
        o = owner_of(dep, bid, W)
        if o != r:
            uniq[(r, dep._name, bid)] = (o, dep)
    send_items = {p: [] for p in range(W)}
    recv_items = {p: [] for p in range(W)}
    for (r, name, bid) in sorted(uniq):
        o, dep = uniq[(r, name, bid)]
        nb = math.prod(dep.block_shape(bid)) * dep.dtype.itemsize
        if o == me:
            send_items[r].append((dep, bid, nb))
        if r == me:
            recv_items[o].append((dep, bid, nb))
    return send_items, recv_items


def _fetch_blocks(ex: Executor, wanted, stores):
    """Execute a ``plan_block_fetch`` schedule over NCCL; ``stores``: {dep name: BlockStore}.
    Returns {(dep name, block id): DeviceChunk} for the blocks this rank received."""
    W, me = ex.world.size, ex.world.rank
    send_items, recv_items = plan_block_fetch(wanted, W, me)
    pad = lambda n: -(-n // 256) * 256
    sends, recvs, keep, out = [], [], [], {}
    for p in range(W):
        if send_items[p]:
            buf = alloc_bytes(sum(pad(nb) for *_, nb in send_items[p]), ex.device)
            off, copies = 0, []
            for dep, bid, nb in send_items[p]:
                blk = stores[dep._name].blocks[bid]
                flat = DeviceChunk(buf, blk.shape, blk.dtype, offset=off // blk.itemsize)
                copies.extend(_copy_descs(blk, flat, blk.itemsize))
                off += pad(nb)
            g = rt.GatherLaunch(copies)
            ex._do(g.run)
            keep.append(g)
            sends.append((p, buf))
        if recv_items[p]:
            buf = alloc_bytes(sum(pad(nb) for *_, nb in recv_items[p]), ex.device)
            off = 0
            for dep, bid, nb in recv_items[p]:
                out[(dep._name, bid)] = DeviceChunk(buf, dep.block_shape(bid), dep.dtype, offset=off // dep.dtype.itemsize)
                off += pad(nb)
            recvs.append((p, buf))
    if sends or recvs:
        ex._do(lambda: _p2p_exchange(ex, sends, recvs))
    out["__keep__"] = (keep, sends, recvs)
    return out


def plan_rechunk_exchange(expr: TasksRechunk, W: int, me: int):
    """Pure schedule of the rectangles a rechunk moves across the partition (the all-to-all of
    SURVEY.md 8e).  Items: (old block id, new block id, source slices, piece shape, nbytes)."""
    x = expr.operand("array")
    item = expr.dtype.itemsize
    send_items = {p: [] for p in range(W)}
    recv_items = {p: [] for p in range(W)}
    for nbid in expr.block_ids():
        r = owner_of(expr, nbid, W)
        for obid, sl, dsl in expr.pieces(nbid):
            o = owner_of(x, obid, W)
            if o == r:
                continue
            shape = tuple(s.stop - s.start for s in sl)
            nb = math.prod(shape) * item
            if o == me:
                send_items[r].append((obid, nbid, sl, shape, nb))
            if r == me:
                recv_items[o].append((obid, nbid, sl, shape, nb))
    return send_items, recv_items


def plan_rechunk_push(expr: TasksRechunk, W: int, me: int):
    """Pure (no device) plan of a rechunk across the partition as ONE gather per rank that stores
    straight into the owners' memory.  Every rank lays out the new blocks of every rank the same
    way (one slab per rank, blocks in block-id order, 512-byte aligned).  Returns
    ``(layout, totals, pushes)``: ``layout[r] = {new block id: byte offset in rank r's slab}``,
    ``totals[r]`` = slab bytes, ``pushes`` = [(old block id, source slices, owner rank of the new
    block, new block id, destination slices)] for every piece whose SOURCE block ``me`` owns --
    local pieces included: the same launch moves them."""
    x = expr.operand("array")
    item = expr.dtype.itemsize
    layout = [dict() for _ in range(W)]
    totals = [0] * W
    pushes = []
    per_dest = [[] for _ in range(W)]
    for nbid in expr.block_ids():
        r = owner_of(expr, nbid, W)
        layout[r][nbid] = totals[r]
        totals[r] += -(-math.prod(expr.block_shape(nbid)) * item // 512) * 512
        for obid, sl, dsl in expr.pieces(nbid):
            if owner_of(x, obid, W) == me:
                per_dest[r].append((obid, sl, r, nbid, dsl))
    # All-to-all schedule: consecutive pieces go to DIFFERENT owners, starting with the right-hand
    # neighbour -- at any moment rank r stores to r+1, r+2, ... and no owner is the target of every
    # rank at once (in block-id order all ranks would hammer rank 0's NVLink ingress first).
    rot = [per_dest[(me + 1 + k) % W] for k in range(W)]
    for i in range(max((len(q) for q in rot), default=0)):
        for q in rot:
            if i < len(q):
                pushes.append(q[i])
    return layout, totals, pushes


def _rechunk_push(ex: Executor, expr: TasksRechunk, src: BlockStore, st: BlockStore):
    """The all-to-all of a rechunk as the rechunk kernel itself: every rank's tiled
    gather reads its own old blocks and writes the pieces into the new blocks where they live --
    local HBM or a peer's HBM over NVLink (``_peer``) -- bracketed by two stream-ordered barriers.
    Bytes per element: one read + one write, wherever the destination is."""
    W, me = ex.world.size, ex.world.rank
    item = expr.dtype.itemsize
    layout, totals, pushes = plan_rechunk_push(expr, W, me)
    slab = alloc_bytes(totals[me], ex.device)
    for nbid, off in layout[me].items():
        st.blocks[nbid] = DeviceChunk(slab, expr.block_shape(nbid), expr.dtype, offset=off // item)
    bases = [p[0] for p in _peer.exchange_pointers(ex.device, [slab.data_ptr()], [1] * W, me)]
    windows = [slab if r == me else _peer.PeerBuffer(bases[r], ex.device, totals[r], r) for r in range(W)]
    copies = []
    for obid, sl, r, nbid, dsl in pushes:
        dst = DeviceChunk(windows[r], expr.block_shape(nbid), expr.dtype, offset=layout[r][nbid] // item)
        copies.extend(_copy_descs(src.blocks[obid][sl], dst[dsl], item))
    launch = rt.GatherLaunch(copies)
    bar = _peer.StreamBarrier(ex.device, me, W)
    ex._do(bar)                 # every owner is done with the previous contents of its slab
    ex._do(launch.run)
    ex._do(bar)                 # every piece has landed before anyone reads a new block
    st.keepalive.extend([launch, slab, bar, windows])
    return st


def _interleave_remote_reads(blocks, owners, me: int, W: int, in_items, out_item: int, band_rows: int = 256):
    """Element-wise blocks whose operands sit in peers' memory are cut into row bands and dealt so
    that consecutive bands read from DIFFERENT peers, starting with the right-hand neighbour: every
    NVLink port pair is busy all the time instead of all ranks pulling from rank 0 first."""
    if len(owners) != len(blocks) or not any(o != me for o in owners):
        return blocks
    per_owner = [[] for _ in range(W)]
    for b, o in zip(blocks, owners):
        if len(b.shape) != 2 or b.shape[0] <= band_rows or b.out1:
            per_owner[o].append(b)
            continue
        R, Ccols = b.shape
        for a in range(0, R, band_rows):
            n = min(band_rows, R - a)
            ins = [(ptr + a * st[0] * item, st) for (ptr, st), item in zip(b.inputs, in_items)]
            per_owner[o].append(rt.BlockArgs(shape=(n, Ccols), inputs=ins, out0=b.out0 + a * Ccols * out_item))
    rot = [per_owner[(me + 1 + k) % W] for k in range(W)]
    out = []
    for i in range(max(len(q) for q in rot)):
        for q in rot:
            if i < len(q):
                out.append(q[i])
    return out


def _push_views(ex: Executor, expr, st: BlockStore, src: BlockStore, moves):
    """Output blocks of a structural expression (slice, concatenate, expand_dims ...) whose source block
    lives on ANOTHER GPU: the source's owner stores the selected view straight into the block at its new
    owner -- one gather launch per rank over peer memory, bracketed by the stream barrier (the same
    mechanism as the rechunk all-to-all).  ``moves``: [(output block id, source owner, view(block), source
    block id)] in the same order on every rank."""
    if not _peer.enabled():
        raise NotImplementedError(f"{type(expr).__name__} that moves blocks between GPUs needs the peer-memory "
                                  "path (B2_COMM=peer); rechunk first")
    W, me = ex.world.size, ex.world.rank
    item = expr.dtype.itemsize
    layout, totals = {}, [0] * W
    for bid, _, _, _ in moves:
        r = ex.world.owner(expr, bid)
        layout[bid] = (r, totals[r])
        totals[r] += -(-math.prod(expr.block_shape(bid)) * item // 512) * 512
    slab = alloc_bytes(totals[me], ex.device)
    bases = [p[0] for p in _peer.exchange_pointers(ex.device, [slab.data_ptr()], [1] * W, me)]
    windows = [slab if r == me else _peer.PeerBuffer(bases[r], ex.device, totals[r], r) for r in range(W)]
    copies = []
    for bid, src_owner, view, ibid in moves:
        r, off = layout[bid]
        dst = DeviceChunk(windows[r], expr.block_shape(bid), expr.dtype, offset=off // item)
        if r == me:
            st.blocks[bid] = dst
        if src_owner == me:
            copies.extend(_copy_descs(view(src.blocks[ibid]), dst, item))
    launch = rt.GatherLaunch(copies)
    bar = _peer.StreamBarrier(ex.device, me, W)
    ex._do(bar)
    ex._do(launch.run)
    ex._do(bar)
    st.keepalive.extend([launch, slab, windows])


def plan_fused_peer_reads(plan: FusedPlan, replicated, W: int):
    """Pure plan of the remote block reads of a fused expression: ``exports[o]`` = the (leaf index,
    leaf block id) pairs rank o owns and some other rank reads, in one canonical order;
    ``readers[(o, i)]`` = the set of ranks reading export i of rank o."""
    expr = plan.fused
    wanted = {}
    for bid in expr.block_ids():
        r = owner_of(expr, bid, W)
        for k, (dep, _) in enumerate(plan.leaves):
            if replicated[k]:
                continue
            lbid = plan.leaf_block_id(k, bid)
            o = owner_of(dep, lbid, W)
            if o != r:
                wanted.setdefault((o, dep._name, lbid), [k, set()])[1].add(r)
    exports = [[] for _ in range(W)]
    readers = {}
    for (o, name, lbid) in sorted(wanted):
        k, rs = wanted[(o, name, lbid)]
        readers[(o, len(exports[o]))] = rs
        exports[o].append((k, lbid))
    return exports, readers


def _peer_reads_for_fused(ex: Executor, plan: FusedPlan, deps):
    """Remote operands of a fused expression (``x.T + x`` across the partition) are read IN PLACE:
    the owners export the blocks once, the fused kernel of the reading rank loads them over NVLink
    while it computes -- no pack, no send/recv, no staging copy.  Returns {(dep name, block id):
    DeviceChunk over peer memory}; ``__barrier__`` must bracket the launch on the tape."""
    import struct

    W, me = ex.world.size, ex.world.rank
    exports, readers = plan_fused_peer_reads(plan, [d.replicated for d in deps], W)
    if not any(exports):
        return {}
    rec = _peer.HANDLE_BYTES + 8 * 9
    mine = []
    for k, lbid in exports[me]:
        blk = deps[k].blocks[lbid]
        st = list(blk.strides) + [0] * (8 - blk.ndim)
        mine.append(_peer.export_handle(blk.ptr) + struct.pack("<9q", blk.ndim, *st))
    recs = _peer.exchange_records(ex.device, mine, [len(e) for e in exports], rec)
    out = {}
    for o in range(W):
        if o == me:
            continue
        for i, (k, lbid) in enumerate(exports[o]):
            if me not in readers[(o, i)]:
                continue
            raw = recs[o][i]
            meta = struct.unpack("<9q", raw[_peer.HANDLE_BYTES:])
            dep = plan.leaves[k][0]
            ptr = _peer.open_handle(raw[: _peer.HANDLE_BYTES])
            out[(dep._name, lbid)] = DeviceChunk(_peer.PeerBuffer(ptr, ex.device, owner=o), dep.block_shape(lbid),
                                                 dep.dtype, strides=meta[1: 1 + meta[0]])
    out["__barrier__"] = _peer.StreamBarrier(ex.device, me, W)
    return out


def _exchange_for_fused(ex: Executor, plan: FusedPlan, deps, out_ids):
    """Blocks of dependencies that some rank's output blocks read but another rank owns are
    packed per peer, exchanged over NCCL and exposed as DeviceChunks."""
    W, me = ex.world.size, ex.world.rank
    if _peer.enabled():
        return _peer_reads_for_fused(ex, plan, deps)
    send_items, recv_items = plan_fused_exchange(plan, [d.replicated for d in deps], W, me)
    if not any(send_items.values()) and not any(recv_items.values()):
        return {}
    pad = lambda n: -(-n // 256) * 256
    sends, recvs, keep, out = [], [], [], {}
    for p in range(W):
        if send_items[p]:
            buf = alloc_bytes(sum(pad(nb) for _, _, nb in send_items[p]), ex.device)
            off, copies = 0, []
            for k, lbid, nb in send_items[p]:
                blk = deps[k].blocks[lbid]
                flat = DeviceChunk(buf, blk.shape, blk.dtype, offset=off // blk.itemsize)
                copies.extend(_copy_descs(blk, flat, blk.itemsize))
                off += pad(nb)
            g = rt.GatherLaunch(copies)
            ex._do(g.run)
            keep.append(g)
            sends.append((p, buf))
        if recv_items[p]:
            buf = alloc_bytes(sum(pad(nb) for _, _, nb in recv_items[p]), ex.device)
            off = 0
            for k, lbid, nb in recv_items[p]:
                dep = plan.leaves[k][0]
                out[(dep._name, lbid)] = DeviceChunk(buf, dep.block_shape(lbid), dep.dtype, offset=off // dep.dtype.itemsize)
                off += pad(nb)
            recvs.append((p, buf))
    ex._do(lambda: _p2p_exchange(ex, sends, recvs))
    out["__keep__"] = (keep, sends, recvs)
    return out


def _exchange_for_rechunk(ex: Executor, expr: TasksRechunk, src: BlockStore, new_ids):
    """All-to-all of the rectangles a rechunk moves across the partition: pack (gather kernel)
    -> NCCL send/recv -> the local gather reads the received pieces in place."""
    W, me = ex.world.size, ex.world.rank
    item = expr.dtype.itemsize
    pad = lambda n: -(-n // 256) * 256
    send_items, recv_items = plan_rechunk_exchange(expr, W, me)
    sends, recvs, keep, out = [], [], [], {}
    for p in range(W):
        if send_items[p]:
            buf = alloc_bytes(sum(pad(it[-1]) for it in send_items[p]), ex.device)
            off, copies = 0, []
            for obid, nbid, sl, shape, nb in send_items[p]:
                piece = src.blocks[obid][sl]
                flat = DeviceChunk(buf, shape, expr.dtype, offset=off // item)
                copies.extend(_copy_descs(piece, flat, item))
                off += pad(nb)
            g = rt.GatherLaunch(copies)
            ex._do(g.run)
            keep.append(g)
            sends.append((p, buf))
        if recv_items[p]:
            buf = alloc_bytes(sum(pad(it[-1]) for it in recv_items[p]), ex.device)
            off = 0
            for obid, nbid, sl, shape, nb in recv_items[p]:
                out[(obid, nbid)] = DeviceChunk(buf, shape, expr.dtype, offset=off // item)
                off += pad(nb)
            recvs.append((p, buf))
    ex._do(lambda: _p2p_exchange(ex, sends, recvs))
    out["__keep__"] = (keep, sends, recvs)
    return out


# ----------------------------------------------------------------------------- results to host
def gather_to_host(ex: Executor, expr: ArrayExpr, store: BlockStore) -> np.ndarray:
    """finalize -> concatenate3 (``_core_utils.py:1426-1448``): assemble the blocks on the host.
    With several ranks every rank returns the full array (blocks travel as host objects)."""
    torch.cuda.synchronize()
    local = {bid: blk.to_numpy() for bid, blk in store.blocks.items()}
    if ex.world.size > 1 and not store.replicated:
        import torch.distributed as dist

        allb = [None] * ex.world.size
        dist.all_gather_object(allb, local)
        local = {k: v for d in allb for k, v in d.items()}
    if expr.ndim == 0:
        return local[()].reshape(())[()]
    out = np.empty(expr.shape, dtype=expr.dtype)
    for bid in expr.block_ids():
        start, shape = expr.block_start(bid), expr.block_shape(bid)
        if math.prod(shape) == 0:
            continue
        out[tuple(slice(s, s + n) for s, n in zip(start, shape))] = local[bid]
    return out
