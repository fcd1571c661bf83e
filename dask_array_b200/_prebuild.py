"""Pre-compile (NVRTC -> sm_100a cubin) the fused kernels of the BASELINE configs into the
in-tree JIT cache.  Runs without a GPU (``__graft_entry__.build()`` calls it): it proves the
generated CUDA builds here and lets the GPU box start warm."""
from __future__ import annotations

import numpy as np

from . import _codegen as cg
from . import _lib
from . import _runtime as rt


def _chain(dtype="float32"):
    p = cg.Program()
    x = p.add_input(dtype)
    p.set_output(p.op("add", p.op("multiply", p.op("sin", x), p.const(2)), p.op("power", x, p.const(2))))
    return p


def _identity(dtype):
    p = cg.Program()
    p.set_output(p.op("positive", p.add_input(dtype)))
    return p


def config_kernels():
    """(name, program, layouts, mode, redop, canonical shape, vec, acc dtype)"""
    f4, f8 = np.dtype("float32"), np.dtype("float64")
    out = [
        ("c2 mean(axis=0) chunk", _chain(), ("V",), _lib.MODE_R, _lib.RED_SUM, (1, 4096, 4096), 4, f4),
        ("c2 std() chunk", _chain(), ("V",), _lib.MODE_RC, _lib.RED_MOMENT, (1, 2048, 8192), 4, f4),
        ("c2 chain materialised", _chain(), ("V",), _lib.MODE_EW, _lib.RED_NONE, (1, 4096, 4096), 4, f4),
    ]
    for nm, red in (("argmax", _lib.RED_ARGMAX), ("argmin", _lib.RED_ARGMIN), ("max", _lib.RED_MAX), ("min", _lib.RED_MIN)):
        out.append((f"c3 {nm}(axis=1) chunk", _identity(f8), ("V",), _lib.MODE_C, red, (1, 8192, 16384), 2, f8))
    return out


def extra_kernels():
    """(name, program, KernelSpec) of the kernels whose geometry is fixed by the launch builder: the
    mirror-pair kernel of c4's ``x.T + x`` and the cumulative scans."""
    out = []
    for dt, v in (("float32", 4), ("float64", 2)):
        p = cg.Program()
        a, b = p.add_input(dt), p.add_input(dt)
        p.set_output(p.op("add", a, b))
        out.append((f"c4 x.T + x mirror pair {dt}", p,
                    cg.KernelSpec(p.key(), ("T", "V"), _lib.MODE_EW, _lib.RED_NONE, vec=v, tx=16, ty=16, rpt=16 * v,
                                  unroll=1, acc_dtype=dt, variant="sym")))
        q = cg.Program()
        q.set_output(q.op("astype", q.add_input(dt), dtype=np.dtype(dt)))
        out.append((f"cumsum rows {dt}", q, cg.KernelSpec(q.key(), ("V",), _lib.MODE_SR, _lib.RED_SUM, vec=v, tx=128, ty=1,
                                                        rpt=1 << 30, unroll=8, acc_dtype=dt)))
        out.append((f"cumsum columns {dt}", q, cg.KernelSpec(q.key(), ("V",), _lib.MODE_SC, _lib.RED_SUM, vec=v, tx=32, ty=8,
                                                           rpt=8, unroll=4, acc_dtype=dt)))
        out.append((f"cumsum rows, chained single pass {dt}", q, cg.KernelSpec(
            q.key(), ("V",), _lib.MODE_SR, _lib.RED_SUM, vec=v, tx=32, ty=1, rpt=1 << 30, unroll=32 if v == 4 else 16, acc_dtype=dt)))
    return out


def prebuild(verbose: bool = False):
    n = 0
    for name, prog, layouts, mode, redop, shape, vec, acc in config_kernels():
        geo = cg.choose_geometry(prog, mode, [shape], vec)
        spec = cg.KernelSpec(prog.key(), layouts, mode, redop, acc_dtype=acc.name, **geo)
        cubin = rt.compile_kernel(prog, spec)
        n += 1
        if verbose:
            print(f"{name}: {len(cubin)} bytes ({spec.digest()})")
    for name, prog, spec in extra_kernels():
        cubin = rt.compile_kernel(prog, spec)
        n += 1
        if verbose:
            print(f"{name}: {len(cubin)} bytes ({spec.digest()})")
    return n
