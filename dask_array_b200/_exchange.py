"""Moving blocks between GPUs (one process per GPU): block-cyclic ownership, the pure planners every
rank evaluates identically (no negotiation traffic), the peer-memory data path (stores into / loads from
the peers' HBM over NVLink, ``_peer``) and the packed NCCL exchange kept for comparison
(``B2_COMM=nccl``).  Used by ``_executor.Executor``; SURVEY.md section 8e.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch

from . import _peer
from . import _runtime as rt
from ._blockwise import FusedPlan
from ._device import DeviceChunk, alloc_bytes
from ._expr import ArrayExpr
from ._rechunk import TasksRechunk

_RUN = 4096        # bytes per row a long contiguous run is folded into (== B2_GATHER_COL_BYTES)


def _fold_runs(descs):
    """A rectangle that is ONE contiguous run on both sides (whole blocks, halo bodies: rows == 1 or both
    pitches == the row length) would be tiled as a single row of 4 KiB column tiles -- a 4 KiB CTA each
    (measured on B200: 1.7 TB/s).  Folded into rows of 4 KiB the planner makes 64 KiB tiles of it."""
    out = []
    for s, d, rows, rb, sp, dp in descs:
        total = rows * rb
        if (rows == 1 or (sp == rb and dp == rb)) and total >= 32 * _RUN and s % 16 == 0 and d % 16 == 0:
            body = total // _RUN
            out.append((s, d, body, _RUN, _RUN, _RUN))
            tail = total - body * _RUN
            if tail:
                out.append((s + body * _RUN, d + body * _RUN, 1, tail, tail, tail))
        else:
            out.append((s, d, rows, rb, sp, dp))
    return out


def _copy_descs(src: DeviceChunk, dst: DeviceChunk, item: int):
    """2-D copy rectangles moving ``src`` into ``dst`` (same shape, arbitrary strides with a
    unit-stride innermost run)."""
    return _fold_runs(_copy_descs_raw(src, dst, item))


def _copy_descs_raw(src: DeviceChunk, dst: DeviceChunk, item: int):
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


# ----------------------------------------------------------------------------- tree partials
def _partial_fields(x, kind, bid):
    """[(field name, shape, dtype)] of the partial block ``bid`` of ``x`` -- derivable on every rank from
    the expression alone (``mean_chunk`` / ``moment_chunk`` / ``arg_chunk`` results, ``_common.py:270-281,
    368-404, 704-732``, in their device representation, see ``_reductions``)."""
    kshape = tuple(x.block_shape(bid))
    if kind == "moment":
        return [("", kshape + (3,), np.dtype(np.float64))]
    if kind == "mean":
        return [("total", kshape, x.dtype)]
    if kind == "arg":
        leaf = x
        while type(leaf).__name__ == "PartialReduce":
            leaf = leaf.operand("array")
        return [("arg", kshape, np.dtype(np.int64)), ("vals", kshape, leaf.operand("array").dtype)]
    return [("", kshape, x.dtype)]


def mean_count(x, bid):
    """Element count behind the ``total`` of partial block ``bid`` (the ``n`` of ``mean_chunk``
    ``_common.py:278-279``, summed by ``mean_combine`` :284-305): a host integer, from shapes alone."""
    if type(x).__name__ == "PartialReduce":
        for key, members in x.groups():
            if key == bid:
                return sum(mean_count(x.operand("array"), m) for m in members)
        raise KeyError(bid)
    red = x.root if type(x).__name__ == "FusedBlockwise" else x        # the ChunkReduce
    top = red.operand("array")
    return math.prod(top.block_shape(bid)[a] for a in red.operand("axis"))


def partials_layout(x, kind, W: int):
    """One slab per rank holding that rank's partial blocks of ``x`` in block-id order, every field
    16-byte aligned.  Returns ``(layout, sizes)``: ``layout[r][bid] = [(name, byte offset, shape, dtype)]``."""
    layout = [dict() for _ in range(W)]
    sizes = [0] * W
    for bid in x.block_ids():
        r = owner_of(x, bid, W)
        ent = []
        for name, shp, dt in _partial_fields(x, kind, bid):
            ent.append((name, sizes[r], shp, dt))
            sizes[r] += -(-math.prod(shp) * dt.itemsize // 16) * 16
        layout[r][bid] = ent
    return layout, sizes


def level_is_local(expr, W: int) -> bool:
    """Pure: does every group of the ``PartialReduce`` level ``expr`` live on the rank that owns the group's output
    block?  Then each owner folds its own groups and the level needs no exchange (``mean(axis=0)`` with the block
    columns dealt to the ranks, ``argmax(axis=1)`` of row panels); otherwise the partials are all-gathered."""
    x = expr.operand("array")
    return all(owner_of(x, m, W) == owner_of(expr, key, W) for key, members in expr.groups() for m in members)


def alloc_partials(ex, x, kind):
    """Allocate this rank's partial blocks of ``x`` inside ONE slab laid out by ``partials_layout`` (the
    slab is the send buffer of the all-gather: no pack step).  Returns (slab, {bid: {field: DeviceChunk}})."""
    layout, sizes = partials_layout(x, kind, ex.world.size)
    me = ex.world.rank
    slab = alloc_bytes(max(sizes[me], 16), ex.device)
    out = {}
    for bid, ent in layout[me].items():
        out[bid] = {name: DeviceChunk(slab, shp, dt, offset=off // dt.itemsize) for name, off, shp, dt in ent}
    return slab, out


_SMALL_ALLGATHER = 64 * 1024


def _allgather_blocks(ex, src: BlockStore, x):
    """All-gather the per-block partials of ``x`` so every rank can fold the tree (the exchange between
    the chunk step and ``PartialReduce``, ``reductions/_reduction.py:751-806``).  Peer path: every rank
    stores its slab into slot ``rank`` of every rank's window over NVLink between two stream-ordered
    barriers -- one launch for small payloads (``b2_peer_allgather``), barrier / tiled gather / barrier
    for big ones.  ``B2_COMM=nccl``: ``all_gather_into_tensor``."""
    import torch.distributed as dist

    W, me = ex.world.size, ex.world.rank
    kind = src.kind
    layout, sizes = partials_layout(x, kind, W)
    cap = max(max(sizes), 16)
    send = getattr(src, "slab", None)
    keep = []
    if send is None:
        send = alloc_bytes(cap, ex.device)
        copies = []
        for bid, ent in layout[me].items():
            blk = src.blocks[bid]
            for name, off, shp, dt in ent:
                chunk = blk[name] if isinstance(blk, dict) else blk
                nb = math.prod(shp) * dt.itemsize
                if nb:
                    if not chunk.is_contiguous:
                        raise NotImplementedError("all-gather of a non-contiguous partial block")
                    copies.append((chunk.ptr, send.data_ptr() + off, 1, nb, nb, nb))
        g = rt.GatherLaunch(copies)
        ex._do(g.run)
        keep.append(g)
    recv = alloc_bytes(cap * W, ex.device)
    bar = _peer.StreamBarrier(ex.device, me, W) if _peer.enabled() else None
    if bar is not None and getattr(bar, "allgather", None) is not None:
        bases = [p[0] for p in _peer.exchange_pointers(ex.device, [recv.data_ptr()], [1] * W, me)]
        table = torch.tensor(bases, dtype=torch.int64).to(ex.device)
        nb_me = sizes[me]
        if cap <= _SMALL_ALLGATHER:
            ex._do(lambda: bar.allgather(table.data_ptr(), send.data_ptr(), nb_me, cap), collective=True)
        else:
            push = rt.GatherLaunch([(send.data_ptr(), bases[(me + 1 + k) % W] + me * cap, 1, nb_me, nb_me, nb_me)
                                    for k in range(W)] if nb_me else [])
            ex._do(bar, collective=True)
            ex._do(push.run)
            ex._do(bar, collective=True)
            keep.append(push)
        keep.append(table)
    else:
        if send.numel() != cap:
            full = alloc_bytes(cap, ex.device)
            ex._do(lambda: full[: send.numel()].copy_(send))
            keep.append(send)
            send = full
        ex._do(lambda: dist.all_gather_into_tensor(recv, send), collective=True)
    out = {}
    for r in range(W):
        for bid, ent in layout[r].items():
            parts = {name: DeviceChunk(recv, shp, dt, offset=(r * cap + off) // dt.itemsize) for name, off, shp, dt in ent}
            if kind == "mean":
                out[bid] = {"total": parts["total"], "n": mean_count(x, bid)}
            elif kind == "arg":
                out[bid] = parts
            else:
                out[bid] = parts[""]
    src.keepalive.extend([send, recv, keep])
    return out


def _p2p_exchange(ex, sends, recvs):
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


def _fetch_blocks(ex, wanted, stores):
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
        ex._do(lambda: _p2p_exchange(ex, sends, recvs), collective=True)
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


def _rechunk_push(ex, expr: TasksRechunk, src: BlockStore, st: BlockStore):
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
    ex._do(bar, collective=True)                 # every owner is done with the previous contents of its slab
    ex._do(launch.run)
    ex._do(bar, collective=True)                 # every piece has landed before anyone reads a new block
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


def _push_views(ex, expr, st: BlockStore, src: BlockStore, moves):
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
    ex._do(bar, collective=True)
    ex._do(launch.run)
    ex._do(bar, collective=True)
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


def _peer_reads_for_fused(ex, plan: FusedPlan, deps):
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


def _exchange_for_fused(ex, plan: FusedPlan, deps, out_ids):
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
    ex._do(lambda: _p2p_exchange(ex, sends, recvs), collective=True)
    out["__keep__"] = (keep, sends, recvs)
    return out


def _exchange_for_rechunk(ex, expr: TasksRechunk, src: BlockStore, new_ids):
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
    ex._do(lambda: _p2p_exchange(ex, sends, recvs), collective=True)
    out["__keep__"] = (keep, sends, recvs)
    return out




# ----------------------------------------------------------------------------- results to host
def gather_to_host(ex, expr: ArrayExpr, store: BlockStore) -> np.ndarray:
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

