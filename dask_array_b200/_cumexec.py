"""Launch builder of the cumulative reductions (``reductions/_cumulative.py:100-265``): per-segment totals
with the reduction kernels, the scan of the totals with the scan kernel itself (recursive for long
vectors), one scan pass with carry.  Used by ``Executor._run_CumReduction``.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch

from . import _codegen as cg
from . import _lib
from . import _runtime as rt
from ._device import DeviceChunk, alloc_bytes

class _Cum:
    """Launch builder of one cumulative reduction (see ``Executor._run_CumReduction``)."""

    SEG = 4096            # elements per virtual row of a 1-D block (16-32 KiB: one warp's worth of work)

    def __init__(self, ex, st, acc, redop):
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

    # ---- N-d blocks, one GPU: single pass over chains of blocks (2 N bytes)
    def nd_chained(self, x, src, axis) -> bool:
        """Every chain of blocks along ``axis`` is walked by the threads that own its columns (scan along
        rows) or by the warps that own its rows (scan along the contiguous dim), running total in registers:
        each element is read once and written once.  Needs enough independent columns / rows to keep HBM
        busy; returns False (nothing done) when the array is too narrow -- the three-step plan then splits
        the axis itself."""
        import os

        ex, st, acc = self.ex, self.st, self.acc
        if ex.world.size != 1 or os.environ.get("B2_SCAN_CHAINED", "1") != "1":
            return False
        nax = x.numblocks[axis]
        others = [range(n) for d, n in enumerate(x.numblocks) if d != axis]
        chains, lanes, mode = [], 0, None
        for cid in itertools.product(*others):
            bids = [cid[:axis] + (i,) + cid[axis:] for i in range(nax)]
            ch = []
            for bid in bids:
                c = src.blocks[bid]
                if c.size == 0:
                    continue
                canon = rt.canonicalize_scan(c.shape, [c.strides], axis)
                if mode is None:
                    mode = canon.mode
                elif canon.mode != mode:
                    return False
                ch.append((bid, c, canon))
            if not ch:
                continue
            canon = ch[0][2]
            if mode == _lib.MODE_SR:
                if any(k[2].C != canon.C or k[2].B != canon.B for k in ch):
                    return False
                lanes += canon.B * canon.C
            else:
                if any(k[2].R != canon.R or k[2].B != canon.B for k in ch):
                    return False
                lanes += canon.B * canon.R * 32
            chains.append(ch)
        if not chains:
            return False
        # independent lanes needed to cover HBM latency: ~4 MB in flight (column strips hold 512 B each,
        # row warps 4 x 512 B)
        need = (16384 if mode == _lib.MODE_SR else 32 * 1024) * max(1, 4 // acc.itemsize if mode == _lib.MODE_SR else 1)
        if lanes < need:
            return False
        launch_chains = []
        for ch in chains:
            out_chain = []
            for bid, c, _ in ch:
                out = DeviceChunk.empty(c.shape, acc, ex.device)
                st.blocks[bid] = out
                out_chain.append(rt.BlockArgs(shape=c.shape, inputs=[(c.ptr, c.strides)], out0=out.ptr))
            launch_chains.append(out_chain)
        for bid in x.block_ids():                       # empty blocks along the axis
            if bid not in st.blocks:
                st.blocks[bid] = DeviceChunk.empty(x.block_shape(bid), acc, ex.device)
        self._run([rt.chained_scan_launch(self.prog, self.redop, axis, launch_chains, acc)])
        return True

    # ---- N-d blocks: the blocks along the axis are the segments, one totals table per chain
    def nd_blocks(self, x, src, axis):
        if self.nd_chained(x, src, axis):
            return
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
        if n <= SEG:                 # one warp walks it (longer vectors of totals go through one more level:
            #                           ncu showed ONE warp walking 65 536 totals for 336 us -- 23 % of the 2^28 scan)
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


def _sum_tables(ex, tables, acc):
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
        ex._do(lambda v=view: dist.all_reduce(v), collective=True)


