"""dask_array_b200 -- B200-native execution backend for dask-array's data-parallel hot path.

``import dask_array_b200 as da`` gives the subset of ``import dask_array as da`` that
BASELINE.json's north star names: fused element-wise chains, tree reductions
(sum/mean/var/std/min/max/argmin/argmax), rechunk / transpose / basic slicing and the blocked
matmul.  Every result is computed by hand-written sm_100a CUDA kernels behind the C ABI of
``include/b200da.h``; there is no CPU fallback (importing without the built library fails).
"""
from . import _lib  # noqa: F401  (fails loudly when libb200da.so is missing)
from . import _eager  # noqa: F401  (NEP-13 / NEP-18 on DeviceChunk)
from ._device import DeviceChunk  # noqa: F401
from ._collection import (  # noqa: F401
    UFUNC_NAMES, Array, Compiled, _method, _ufunc, asarray, compile, compute, elemwise, from_array,
    dot, from_host_blocks, full, matmul, nanargmax, nanargmin, nanmax, nanmean, nanmin, nanprod, nanstd, nansum,
    nanvar, ones, random, tensordot, einsum, cumsum, cumprod, nancumsum, nancumprod,
    rechunk, transpose, where, zeros, divmod, modf, frexp, arange, linspace, clip, round, around, swapaxes, moveaxis,
    rollaxis, real, imag, conj, conjugate,
)

from ._views import broadcast_to, concatenate, expand_dims, ravel, squeeze, stack  # noqa: F401,E402
from ._reshape import reshape  # noqa: F401,E402
from . import _overlap as overlap  # noqa: F401,E402  (da.overlap.overlap / trim_internal / map_overlap ...)
from ._overlap import map_blocks, map_overlap, sliding_window_view  # noqa: F401,E402
from ._topk import argtopk, topk  # noqa: F401,E402
from ._routines import diff, flip, fliplr, flipud, roll  # noqa: F401,E402

from ._window import moving_window  # noqa: F401,E402


def move_sum(x, window, min_count=None, axis=-1):
    """bottleneck ``move_sum`` semantics (``MovingWindowReduction``, ``reductions/_sliding_window.py:249-400``)."""
    return moving_window(x, window, "move_sum", min_count, axis)


def move_mean(x, window, min_count=None, axis=-1):
    return moving_window(x, window, "move_mean", min_count, axis)


def move_min(x, window, min_count=None, axis=-1):
    return moving_window(x, window, "move_min", min_count, axis)


def move_max(x, window, min_count=None, axis=-1):
    return moving_window(x, window, "move_max", min_count, axis)


from . import plugin  # noqa: F401,E402  (register / get / lower_reference: the reference-facing boundary)
from .plugin import get, register  # noqa: F401,E402

for _n in UFUNC_NAMES:
    globals()[_n] = _ufunc(_n)
for _n in ("sum", "prod", "mean", "var", "std", "min", "max", "any", "all", "argmin", "argmax"):
    globals()[_n] = _method(_n)
del _n

__all__ = ["Array", "from_array", "asarray", "ones", "zeros", "full", "arange", "linspace", "random", "elemwise", "where",
           "transpose", "rechunk", "matmul", "clip", "round", "around", "swapaxes", "moveaxis", "rollaxis"] + UFUNC_NAMES
