"""Launch runtime: JIT cache, block canonicalisation, persistent launch tables.

One ``FusedLaunch`` = one kernel launch covering every resident block of one fused
expression on this device -- the B200 counterpart of the per-block task dictionary the
reference emits in ``FusedBlockwise._layer`` (``dask_array/_blockwise.py:1690-1728``) and of
the Rust records emitter that expands a layer natively (``crates/dask-array-python``).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _codegen as cg
from . import _lib
from ._device import DeviceChunk, alloc_bytes, current_stream_ptr

_CACHE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_jit_cache")
_lock = threading.Lock()
_cubins: dict[str, bytes] = {}
_handles: dict[tuple, int] = {}


def compile_kernel(program: cg.Program, spec: cg.KernelSpec) -> bytes:
    """cubin of (program, spec); memoised in memory and under ``_jit_cache/`` (works on CPU)."""
    dig = spec.digest()
    with _lock:
        if dig in _cubins:
            return _cubins[dig]
    path = os.path.join(_CACHE_DIR, dig + ".cubin")
    if os.path.exists(path):
        with open(path, "rb") as f:
            cubin = f.read()
    else:
        cubin = _lib.jit_compile(cg.render(program, spec), f"b2_fused_{dig}.cu")
        try:
            os.makedirs(_CACHE_DIR, exist_ok=True)
            tmp = f"{path}.{os.getpid()}.tmp"
            with open(tmp, "wb") as f:
                f.write(cubin)
            os.replace(tmp, path)
        except OSError:
            pass
    with _lock:
        _cubins[dig] = cubin
    return cubin


def load_kernel(program: cg.Program, spec: cg.KernelSpec) -> int:
    """b2_kernel* for the current device."""
    key = (spec.digest(), torch.cuda.current_device())
    with _lock:
        if key in _handles:
            return _handles[key]
    cubin = compile_kernel(program, spec)
    geom = _lib.Geom(spec.mode, spec.redop, spec.vec, spec.tx, spec.ty, spec.rpt,
                     cg.packed_bytes(spec, program.out_dtype), 0)
    out = C.c_void_p()
    _lib.check(_lib.lib.b2_kernel_load(cubin, len(cubin), b"b2_fused", C.byref(geom), C.byref(out)))
    with _lock:
        _handles[key] = out.value
    return out.value


# ----------------------------------------------------------------------------- canonical views
@dataclass
class Canon:
    """One block canonicalised to (B, R, C)."""
    mode: int
    B: int
    R: int
    C: int
    in_strides: list          # per input (sb, sr, sc) in elements
    lead: list = field(default_factory=list)   # extra leading (EW) dims expanded on the host: [(n, [stride per input])]


def canonicalize(shape, in_strides, reduce_axes) -> Canon:
    """Collapse an N-d block (+ per-input element strides, 0 = broadcast) to (B, R, C).

    ``reduce_axes``: set of reduced dims (empty = element-wise).  Adjacent dims of the
    same kind are merged when every input is contiguous across them.
    """
    nd = len(shape)
    red = set(a % nd for a in reduce_axes) if nd else set()
    dims = []   # [n, kind, [stride per input]]
    for d in range(nd):
        if shape[d] == 1:
            continue
        dims.append([int(shape[d]), "X" if d in red else "K", [int(s[d]) for s in in_strides]])
    merged = []
    for n, kind, st in dims:
        if merged and merged[-1][1] == kind and all(p == q * n for p, q in zip(merged[-1][2], st)):
            merged[-1][0] *= n
            merged[-1][2] = st
        else:
            merged.append([n, kind, list(st)])
    nin = len(in_strides)
    zero = [0] * nin

    def pack(b, r, c, mode, lead=()):
        return Canon(mode, b[0], r[0], c[0], [(b[2][k], r[2][k], c[2][k]) for k in range(nin)], list(lead))

    one = [1, "K", zero]
    kinds = "".join(m[1] for m in merged)
    if not red or "X" not in kinds:
        if not red:
            if len(merged) == 1 and all(s in (0, 1) for s in merged[0][2]):
                # one flat contiguous run: give it rows so each thread keeps several loads in flight
                n, _, st = merged[0]
                for cc in (4096, 2048, 1024, 512, 256, 128, 64, 32, 16):
                    if n % cc == 0 and n > cc:
                        return pack(one, [n // cc, "K", [s * cc for s in st]], [cc, "K", st], _lib.MODE_EW)
            lead = [(m[0], m[2]) for m in merged[:-3]]
            tail = merged[-3:]
            while len(tail) < 3:
                tail.insert(0, one)
            return pack(tail[0], tail[1], tail[2], _lib.MODE_EW, lead)
        # reduction over size-1 axes only: a copy, expressed as mode C with one column
        if len(merged) > 1:
            raise NotImplementedError("degenerate reduction over non-contiguous kept dims")
        r = merged[0] if merged else one
        return pack(one, r, one, _lib.MODE_C)
    if kinds == "X":
        n, _, st = merged[0]
        if all(s in (0, 1) for s in st):
            for cc in (8192, 4096, 2048, 1024, 512, 256, 128, 64, 32, 16):
                if n % cc == 0 and n > cc:
                    return pack(one, [n // cc, "X", [s * cc for s in st]], [cc, "X", st], _lib.MODE_RC)
        return pack(one, one, merged[0], _lib.MODE_RC)
    if kinds == "XX":
        return pack(one, merged[0], merged[1], _lib.MODE_RC)
    if kinds == "KX":
        return pack(one, merged[0], merged[1], _lib.MODE_C)
    if kinds == "XK":
        return pack(one, merged[0], merged[1], _lib.MODE_R)
    if kinds == "KXK":
        return pack(merged[0], merged[1], merged[2], _lib.MODE_R)
    if kinds == "KXX":
        return pack(merged[0], merged[1], merged[2], _lib.MODE_RC)
    if kinds == "KKX":
        return pack(merged[0], merged[1], merged[2], _lib.MODE_C)
    # Interleaved patterns (e.g. sum(axis=(0, 2)) of a 3-D block = "XKX"): the kernels only need the
    # kept dims and the reduced dims grouped, not in memory order -- reorder the VIEW (kept dims
    # first, stable; no data moves), merge again.  The output is indexed by the kept dims in their
    # original order, which a stable reorder preserves.
    def merge(seq):
        out = []
        for n, kind, st in seq:
            if out and out[-1][1] == kind and all(p == q * n for p, q in zip(out[-1][2], st)):
                out[-1] = [out[-1][0] * n, kind, st]
            else:
                out.append([n, kind, list(st)])
        return out

    ks, xs = merge([d for d in dims if d[1] == "K"]), merge([d for d in dims if d[1] == "X"])
    # kept groups beyond what (B, R, C) can hold become host-side "lead" dims: one descriptor per
    # index, input AND output pointers offset (the output is contiguous over the kept dims)
    room = 2 if len(xs) == 1 else 1 if len(xs) == 2 else -1
    if room >= 0:
        lead = [(m[0], m[2]) for m in ks[:max(0, len(ks) - room)]]
        ks = ks[max(0, len(ks) - room):]
        if len(xs) == 1:
            while len(ks) < 2:
                ks.insert(0, one)
            return pack(ks[0], ks[1], xs[0], _lib.MODE_C, lead)
        return pack(ks[0] if ks else one, xs[0], xs[1], _lib.MODE_RC, lead)
    raise NotImplementedError(
        f"reduction pattern {kinds!r} (shape {tuple(shape)}, axes {sorted(red)}) needs more than three strided "
        "dimension groups: not supported by the B200 kernels yet")


def canonicalize_scan(shape, in_strides, axis) -> Canon:
    """(B, R, C) view of a block for a cumulative scan along ``axis`` (``_cumulative.py:100-265``):
    dims before the axis -> B (and R when the axis is innermost), the axis -> R (mode SR, dims after it
    -> C) or C (mode SC).  The output is written (B, R, C)-contiguous, i.e. in the block's own order."""
    nd = len(shape)
    axis %= nd
    groups = []      # [n, kind, strides]; kinds: "b" before, "x" the axis, "a" after
    for d in range(nd):
        if shape[d] == 1 and d != axis:
            continue
        kind = "x" if d == axis else ("b" if d < axis else "a")
        st = [int(s[d]) for s in in_strides]
        n = int(shape[d])
        if groups and groups[-1][1] == kind and kind != "x" and all(p == q * n for p, q in zip(groups[-1][2], st)):
            groups[-1][0] *= n
            groups[-1][2] = st
        else:
            groups.append([n, kind, st])
    nin = len(in_strides)
    one = [1, "k", [0] * nin]
    before = [g for g in groups if g[1] == "b"]
    after = [g for g in groups if g[1] == "a"]
    x = next(g for g in groups if g[1] == "x")

    def pack(b, r, c, mode):
        return Canon(mode, b[0], r[0], c[0], [(b[2][k], r[2][k], c[2][k]) for k in range(nin)], [])

    if after:
        if len(after) > 1 or len(before) > 1:
            raise NotImplementedError("cumulative scan over a block whose kept dims are not contiguous")
        return pack(before[0] if before else one, x, after[0], _lib.MODE_SR)
    if len(before) > 2:
        raise NotImplementedError("cumulative scan over a block whose kept dims are not contiguous")
    while len(before) < 2:
        before.insert(0, one)
    return pack(before[0], before[1], x, _lib.MODE_SC)


@dataclass
class BlockArgs:
    """Arguments of one block of a fused launch (all device pointers are ints)."""
    shape: tuple                      # logical N-d shape of the block the chain is evaluated on
    inputs: list                      # [(ptr, strides-in-elements per logical dim)]
    out0: int
    out1: int = 0
    arg_offset: int = 0
    arg_ravel: tuple | None = None    # (block_shape, block_start, total_shape) for axis=None arg reductions


class FusedLaunch:
    """Persistent launch table of one fused expression over its resident blocks."""

    def __init__(self, program: cg.Program, redop: int, reduce_axes, blocks: list[BlockArgs],
                 acc_dtype=None, out_is_contiguous: bool = True, keep_order: bool = False, scan_axis=None,
                 chain_links=None, n_heads=None):
        """``chain_links`` / ``n_heads`` (scan launches): ``blocks`` = the ``n_heads`` first blocks of the chains
        followed by their successors; ``chain_links[i]`` = table index of the block after block i (0 = none)."""
        if not blocks:
            raise ValueError("FusedLaunch needs at least one block")
        if len(program.inputs) >= 2 and len(blocks) > 2 and not keep_order:
            # blocks that read the same input blocks (x.T + x: output (i, j) and (j, i)) become
            # neighbours in the launch, so the second read of a tile can hit the 126 MB L2
            order = sorted(range(len(blocks)), key=lambda i: (tuple(sorted(p for p, _ in blocks[i].inputs)), i))
            blocks = [blocks[i] for i in order]
        self.program = program
        self.redop = redop
        nin = len(program.inputs)
        if scan_axis is not None:
            canons = [canonicalize_scan(b.shape, [st for _, st in b.inputs], scan_axis) for b in blocks]
        else:
            canons = [canonicalize(b.shape, [st for _, st in b.inputs], reduce_axes) for b in blocks]
        modes = {c.mode for c in canons}
        if len(modes) != 1:
            raise NotImplementedError(f"blocks of one launch canonicalise to different modes {modes}")
        self.mode = modes.pop()
        out_dt = program.out_dtype
        acc_dtype = np.dtype(acc_dtype) if acc_dtype is not None else out_dt
        # ---- layout class per input and the widest vector every block allows
        layouts = []
        for k in range(nin):
            cls = {("S" if c.in_strides[k][2] == 0 else "V" if c.in_strides[k][2] == 1 else "G")
                   for c in canons if c.C > 1} or {"V"}
            layouts.append(cls.pop() if len(cls) == 1 else "G")
        # transposed operands of an element-wise launch (contiguous along the OUTPUT's rows, e.g. the
        # x.T of x.T + x): staged through shared memory by b2_run_ewt instead of strided loads
        ewt, staged = False, 0
        if self.mode == _lib.MODE_EW:
            for k in range(nin):
                tile_bytes = -(-64 * 65 * program.inputs[k].itemsize // 16) * 16
                if layouts[k] == "G" and all(c.in_strides[k][1] == 1 for c in canons if c.R > 1) \
                        and all(c.R >= 32 and c.C >= 32 for c in canons) and staged + tile_bytes <= 40 * 1024:
                    layouts[k] = "T"
                    staged += tile_bytes
                    ewt = True
        scan = self.mode in (_lib.MODE_SR, _lib.MODE_SC)
        sizes = [d.itemsize for d in program.inputs] + ([out_dt.itemsize] if self.mode == _lib.MODE_EW else []) \
            + ([acc_dtype.itemsize] if scan else [])
        vmax = max(1, 16 // max(sizes or [out_dt.itemsize]))
        if self.mode == _lib.MODE_R:
            # the row-lane fold stages 256 x V partials in (static) shared memory: keep <= 32 KiB
            probe = cg.KernelSpec("", (), self.mode, redop, 1, 1, 1, 1, 1, acc_dtype.name)
            pb = max(1, cg.packed_bytes(probe, out_dt))
            while vmax > 1 and 256 * vmax * pb > 32768:
                vmax //= 2

        def vec_fits(v):
            for b, c in zip(blocks, canons):
                if c.C % v:
                    return False
                for k, (ptr, _) in enumerate(b.inputs):
                    sb, sr, sc = c.in_strides[k]
                    if layouts[k] == "T":
                        continue
                    if layouts[k] == "V":
                        it = program.inputs[k].itemsize
                        if ptr % (v * it) or (c.B > 1 and sb % v) or (c.R > 1 and sr % v):
                            return False
                if self.mode == _lib.MODE_EW and b.out0 % (v * out_dt.itemsize):
                    return False
                if scan and b.out0 % min(16, v * acc_dtype.itemsize):
                    return False
                for n, sts in c.lead:
                    if any(layouts[k] == "V" and s % v for k, s in enumerate(sts)):
                        return False
            return True

        v = vmax
        while v > 1 and not vec_fits(v):
            v //= 2
        if ewt:
            v = min(v, 4)
        if v == 1:
            layouts = [l if l in ("S", "T") else "G" for l in layouts]
        shapes = [(c.B, c.R, c.C) for c in canons]
        geo = cg.choose_geometry(program, self.mode, shapes, v)
        if ewt:      # 64 x 64 output tiles, 256 threads
            geo = dict(vec=v, tx=64 // v, ty=256 // (64 // v), rpt=64, unroll=1)
        if self.mode == _lib.MODE_SR:     # a thread per column strip walks every row of its block
            cmax = max(c.C for c in canons)
            geo = dict(vec=v, tx=min(128, max(32, cg._pow2_ceil(-(-cmax // v)))), ty=1,
                       rpt=1 << 30, unroll=8)          # one row tile per block (constant: one kernel for all R)
            if chain_links is not None:
                # chained single pass: the only parallelism is over columns, so every thread keeps U vector
                # loads in flight and a CTA is one warp (spreads the warps over all SMs).  Narrow vectors give
                # more warps for the same bytes in flight: pick the widest V that still leaves >= 24k threads.
                heads = canons[:n_heads]
                lanes = sum(c.B * c.C for c in heads)
                vv = v
                while vv > 1 and lanes // vv < 24576:
                    vv //= 2
                vv = int(os.environ.get("B2_SCAN_SR_VEC", vv))
                # loads in flight per thread, as deep as 255 registers allow (cuobjdump: 224 / 232-244 / 242 regs):
                # ncu showed the V=4, U=32 kernel latency-bound at 4 MB in flight (4.5 TB/s, 2.7 % warps active)
                per = max(sizes) * vv
                uu = 64 if per <= 4 else 48 if per <= 8 else 32 if max(sizes) <= 4 else 16
                uu = int(os.environ.get("B2_SCAN_SR_UNROLL", uu))
                if vv == 1:
                    layouts = [l if l in ("S", "T") else "G" for l in layouts]
                v = vv
                geo = dict(vec=v, tx=32, ty=1, rpt=1 << 30, unroll=uu)   # (8-byte x V=2 x 32 loads spills)
        elif self.mode == _lib.MODE_SC:   # a warp per row
            geo = dict(vec=v, tx=32, ty=8, rpt=8, unroll=4)
        variant, mirror, n_primary = "", None, len(blocks)
        if ewt:
            pairing = _mirror_pairs(program, layouts, blocks, canons, v)
            if pairing is not None:
                order, mirror, n_primary = pairing
                blocks = [blocks[i] for i in order]
                canons = [canons[i] for i in order]
                tt = 16 * v                       # a tile row is 16 chunks of 16 bytes (64 f4 / 32 f8)
                geo = dict(vec=v, tx=16, ty=16, rpt=tt, unroll=1)
                variant = "sym"
        self.spec = cg.KernelSpec(program.key(), tuple(layouts), self.mode, redop,
                                  acc_dtype=acc_dtype.name, variant=variant, **geo)
        self.kernel = load_kernel(program, self.spec)

        # ---- descriptor table (leading EW dims expanded into extra descriptors)
        descs = []
        for b, c in zip(blocks, canons):
            lead_ranges = [range(n) for n, _ in c.lead]
            lead_elems = math.prod(n for n, _ in c.lead) if c.lead else 1
            inner = c.B * c.R * c.C
            # output elements produced per descriptor and their size (reductions)
            out_per = {_lib.MODE_EW: inner, _lib.MODE_C: c.B * c.R, _lib.MODE_R: c.B * c.C, _lib.MODE_RC: c.B,
                       _lib.MODE_SR: inner, _lib.MODE_SC: inner}[self.mode]
            if self.mode == _lib.MODE_EW or redop in (_lib.RED_MIN, _lib.RED_MAX, _lib.RED_NANMIN, _lib.RED_NANMAX, _lib.RED_ARGMIN, _lib.RED_ARGMAX):
                out_item = out_dt.itemsize
            elif redop == _lib.RED_MOMENT:
                out_item = 24
            elif redop in (_lib.RED_ANY, _lib.RED_ALL):
                out_item = 1
            else:
                out_item = acc_dtype.itemsize
            for li, idx in enumerate(np.ndindex(*[len(r) for r in lead_ranges]) if c.lead else [()]):
                d = _lib.Block()
                for k, (ptr, _) in enumerate(b.inputs):
                    off = sum(i * c.lead[j][1][k] for j, i in enumerate(idx))
                    d.in_[k] = ptr + off * program.inputs[k].itemsize
                    d.in_sb[k], d.in_sr[k], d.in_sc[k] = c.in_strides[k]
                d.out0 = b.out0 + li * out_per * out_item
                d.out1 = b.out1 + (li * out_per * 8 if b.out1 else 0)
                d.B, d.R, d.C = c.B, c.R, c.C
                d.arg_offset = b.arg_offset
                if b.arg_ravel is not None:
                    bshape, bstart, total = b.arg_ravel
                    if len(bshape) > _lib.B2_MAX_ND:
                        raise NotImplementedError(f"ravelled arg reduction over {len(bshape)} dims")
                    d.arg_ndim = len(bshape)
                    for j in range(len(bshape)):
                        d.arg_shape[j], d.arg_start[j], d.arg_total[j] = bshape[j], bstart[j], total[j]
                descs.append(d)
            assert lead_elems >= 1
        if mirror is not None:
            assert len(descs) == len(mirror)
            for d, m in zip(descs, mirror):
                d.mirror = m
        if chain_links is not None:
            assert len(descs) == len(chain_links) and mirror is None
            for d, m in zip(descs, chain_links):
                d.mirror = m
            n_primary = n_heads
        # mirror-pair launches tile the primary block of each pair only; the partners sit behind
        # them in the table and are reached through `mirror`
        self.nblocks = n_primary if (mirror is not None or chain_links is not None) else len(descs)
        arr = (_lib.Block * len(descs))(*descs)
        need = C.c_size_t()
        tiles = C.c_int64()
        _lib.check(_lib.lib.b2_fused_plan(self.kernel, arr, self.nblocks, None, 0, C.byref(need), C.byref(tiles)))
        self.workspace = alloc_bytes(need.value, zero=True) if need.value else None
        if need.value:
            _lib.check(_lib.lib.b2_fused_plan(self.kernel, arr, self.nblocks, self.workspace.data_ptr(),
                                              need.value, C.byref(need), C.byref(tiles)))
        self.total_tiles = tiles.value
        raw = np.frombuffer(bytes(arr), dtype=np.uint8)
        self.table = torch.from_numpy(raw.copy()).to(torch.device("cuda", torch.cuda.current_device()))
        self.scalars = _lib.Scalars()

    profile = False       # set on an instance: record a CUDA-event pair around every launch

    def run(self, stream: int | None = None) -> None:
        st = current_stream_ptr() if stream is None else stream
        if self.profile:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(_lib.lib.b2_fused_launch(self.kernel, self.table.data_ptr(), self.nblocks,
                                            self.total_tiles, C.byref(self.scalars), st))
        if self.profile:
            e1.record()
            self.__dict__.setdefault("events", []).append((e0, e1))


def _mirror_pairs(program, layouts, blocks, canons, v):
    """Pair the blocks of an ``f(x, x.T)`` launch (``a + a.T``, tests/test_collection.py): block d reads
    ``(N=p, T=q)``, its mirror d' reads ``(N=q, T=p)`` with the transposed shape.  Returns
    ``(order, mirror index per table slot, number of primaries)`` or None when the launch is not of
    that form (the generic staged kernel b2_run_ewt then runs it)."""
    if sorted(layouts) != ["T", "V"]:
        return None
    kn, kt = layouts.index("V"), layouts.index("T")
    dn, dt_ = program.inputs[kn], program.inputs[kt]
    if dn != dt_ or dn.itemsize not in (4, 8) or v * dn.itemsize != 16:
        return None
    sig = {}
    for i, (b, c) in enumerate(zip(blocks, canons)):
        if c.B != 1 or c.lead or c.R % v or c.C % v:
            return None
        if c.in_strides[kn][2] != 1 or c.in_strides[kt][1] != 1:
            return None
        if b.inputs[kt][0] % 16 or b.out0 % 16:
            return None
        key = (b.inputs[kn][0], b.inputs[kt][0])
        if key in sig:
            return None
        sig[key] = i
    mirror_of = {}
    for (pn, pt), i in sig.items():
        j = sig.get((pt, pn))
        if j is None:
            return None
        ci, cj = canons[i], canons[j]
        if (cj.R, cj.C) != (ci.C, ci.R):
            return None
        # the partner reads the SAME memory: its row pitch is this block's transposed column pitch
        if cj.in_strides[kn][1] != ci.in_strides[kt][2] or cj.in_strides[kt][2] != ci.in_strides[kn][1]:
            return None
        if ci.in_strides[kt][2] % v:
            return None
        mirror_of[i] = j
    primaries = [i for i in range(len(blocks)) if i <= mirror_of[i]]
    secondaries = [i for i in range(len(blocks)) if i > mirror_of[i]]
    order = primaries + secondaries
    slot = {i: s for s, i in enumerate(order)}
    return order, [slot[mirror_of[i]] for i in order], len(primaries)


def scan_launches(program, redop, axis, blocks, acc_dtype):
    """Cumulative-scan launches (one per canonical mode among ``blocks``); BlockArgs.out1 = carry."""
    groups = {}
    for b in blocks:
        c = canonicalize_scan(b.shape, [st for _, st in b.inputs], axis)
        groups.setdefault(c.mode, []).append(b)
    return [FusedLaunch(program, redop, (), g, acc_dtype=acc_dtype, keep_order=True, scan_axis=axis)
            for g in groups.values()]


def chained_scan_launch(program, redop, axis, chains, acc_dtype):
    """ONE single-pass launch over chains of blocks along the scanned axis (``chains``: lists of BlockArgs in
    axis order, all of one canonical mode): heads first, successors behind them, linked through ``mirror``."""
    heads = [ch[0] for ch in chains]
    order, links = list(heads), [0] * len(heads)
    for h, ch in enumerate(chains):
        prev = h
        for b in ch[1:]:
            order.append(b)
            links.append(0)
            links[prev] = len(order) - 1
            prev = len(order) - 1
    return FusedLaunch(program, redop, (), order, acc_dtype=acc_dtype, keep_order=True, scan_axis=axis,
                       chain_links=links, n_heads=len(heads))


def fused_launches(program, redop, reduce_axes, blocks, acc_dtype=None, keep_order=False):
    """One FusedLaunch per canonical mode present among ``blocks`` (ragged edge blocks whose
    extent-1 dims collapse can need a different kernel shape than the interior blocks)."""
    groups = {}
    for b in blocks:
        c = canonicalize(b.shape, [st for _, st in b.inputs], reduce_axes)
        groups.setdefault(c.mode, []).append(b)
    return [FusedLaunch(program, redop, reduce_axes, g, acc_dtype=acc_dtype, keep_order=keep_order)
            for g in groups.values()]


# ----------------------------------------------------------------------------- AOT helpers
class CombineLaunch:
    """One PartialReduce output block (reductions/_reduction.py:968-983) as a replayable launch."""

    def __init__(self, redop: int, dtype, parts: list, parts1, nelem: int, out0: int, out1: int = 0,
                 post: int = _lib.POST_NONE, out_dtype=None, count: float = 0.0, ddof: float = 0.0):
        self.fanin = len(parts)
        ptrs = list(parts) + (list(parts1) if parts1 else [0] * self.fanin)
        self.table = torch.tensor(ptrs, dtype=torch.int64).to(torch.device("cuda", torch.cuda.current_device()))
        self.args = (redop, _lib.dtype_code(dtype), nelem, out0, out1, post,
                     _lib.dtype_code(out_dtype if out_dtype is not None else dtype), float(count), float(ddof))

    def run(self, stream: int | None = None) -> None:
        redop, dcode, nelem, out0, out1, post, ocode, count, ddof = self.args
        base = self.table.data_ptr()
        st = current_stream_ptr() if stream is None else stream
        _lib.check(_lib.lib.b2_combine(redop, dcode, base, base + 8 * self.fanin, self.fanin, nelem,
                                       out0, out1, post, ocode, count, ddof, st))


class CombineGroupsLaunch:
    """All output blocks of one PartialReduce level (reductions/_reduction.py:968-983) in ONE launch."""

    def __init__(self, redop: int, dtype, out_dtype, groups: list):
        """groups: dicts with parts, parts1 (or None), nelem, out0, out1, post, count, ddof."""
        dev = torch.device("cuda", torch.cuda.current_device())
        ptrs, offs = [], []
        for g in groups:
            offs.append(len(ptrs))
            ptrs.extend(g["parts"])
            ptrs.extend(g["parts1"] if g.get("parts1") else [0] * len(g["parts"]))
        self.ptr_table = torch.tensor(ptrs or [0], dtype=torch.int64).to(dev)
        base = self.ptr_table.data_ptr()
        arr = (_lib.Group * len(groups))()
        total = 0
        for a, g, off in zip(arr, groups, offs):
            fan = len(g["parts"])
            a.parts = base + 8 * off
            a.parts1 = base + 8 * (off + fan)
            a.out0, a.out1 = g["out0"], g.get("out1", 0)
            a.nelem, a.elem_begin, a.fanin = g["nelem"], total, fan
            a.post, a.count, a.ddof = g.get("post", _lib.POST_NONE), float(g.get("count", 0.0)), float(g.get("ddof", 0.0))
            total += g["nelem"]
        raw = np.frombuffer(bytes(arr), dtype=np.uint8)
        self.table = torch.from_numpy(raw.copy()).to(dev)
        self.n, self.total = len(groups), total
        self.codes = (redop, _lib.dtype_code(dtype), _lib.dtype_code(out_dtype if out_dtype is not None else dtype))

    def run(self, stream: int | None = None) -> None:
        if not self.total:
            return
        st = current_stream_ptr() if stream is None else stream
        _lib.check(_lib.lib.b2_combine_groups(*self.codes, self.table.data_ptr(), self.n, self.total, st))


def combine(redop, dtype, parts, parts1, nelem, out0, out1=0, post=_lib.POST_NONE, out_dtype=None,
            count=0.0, ddof=0.0):
    c = CombineLaunch(redop, dtype, parts, parts1, nelem, out0, out1, post, out_dtype, count, ddof)
    c.run()
    return c


class GatherLaunch:
    """Persistent tiled gather (rechunk / slicing / concatenation) -- b2_gather_*."""

    def __init__(self, copies: list[tuple]):
        """copies: (src_ptr, dst_ptr, rows, row_bytes, src_pitch, dst_pitch)"""
        copies = [c for c in copies if c[2] > 0 and c[3] > 0]
        self.n = len(copies)
        if not self.n:
            return
        arr = (_lib.Copy * self.n)()
        for d, c in zip(arr, copies):
            d.src, d.dst, d.rows, d.row_bytes, d.src_pitch, d.dst_pitch = c
        tiles = C.c_int64()
        _lib.check(_lib.lib.b2_gather_plan(arr, self.n, C.byref(tiles)))
        self.total_tiles = tiles.value
        # TMA bulk path when every rectangle is 16-byte aligned and big enough to matter
        self.bulk = (os.environ.get("B2_GATHER_BULK", "1") == "1" and all(c.vec_bytes == 16 for c in arr)
                     and sum(c.rows * c.row_bytes for c in arr) >= (1 << 20)
                     and all(c.tile_rows * min(c.row_bytes, 4096) <= 65536 for c in arr))
        raw = np.frombuffer(bytes(arr), dtype=np.uint8)
        self.table = torch.from_numpy(raw.copy()).to(torch.device("cuda", torch.cuda.current_device()))

    def run(self, stream: int | None = None) -> None:
        if not self.n:
            return
        st = current_stream_ptr() if stream is None else stream
        fn = _lib.lib.b2_gather_launch_bulk if self.bulk else _lib.lib.b2_gather_launch
        _lib.check(fn(self.table.data_ptr(), self.n, self.total_tiles, st))


def fill(chunk: DeviceChunk, value) -> None:
    v = np.asarray(value).astype(chunk.dtype)
    buf = (C.c_char * chunk.itemsize).from_buffer_copy(v.tobytes())
    _lib.check(_lib.lib.b2_fill(chunk.ptr, chunk.size, chunk.itemsize, C.addressof(buf), current_stream_ptr()))
