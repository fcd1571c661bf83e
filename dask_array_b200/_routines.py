"""Routines the reference composes from basic slices, concatenation and element-wise operators:
``roll`` (``manipulation/_roll.py:8-77``), ``diff`` (``routines/_diff.py:6-80``), ``flip`` / ``flipud`` /
``fliplr`` (``manipulation/_flip.py:10-55``).  Composed the same way here, they inherit the kernels, the
chunk unification and the multi-GPU paths of their parts: the two shifted operands of ``diff`` are
zero-copy views fused into one kernel per device, a ``roll`` is two views under new block ids.
"""
from __future__ import annotations

from collections.abc import Iterable
from numbers import Integral


def _along(ndim, axis, sl):
    """Index tuple selecting ``sl`` along ``axis`` and everything elsewhere."""
    return tuple(sl if d == axis % ndim else slice(None) for d in range(ndim))


def _as_tuple(v):
    return tuple(v) if isinstance(v, Iterable) else (v,)


def flip(m, axis=None):
    """Reverse the element order along ``axis`` (all axes by default): negative-step views."""
    from ._collection import asarray

    m = asarray(m)
    axes = range(m.ndim) if axis is None else _as_tuple(axis)
    if any(not -m.ndim <= ax < m.ndim for ax in axes):
        raise ValueError(f"`axis` of {axis} invalid for {m.ndim}-D array")
    rev = {ax % m.ndim for ax in axes}
    return m[tuple(slice(None, None, -1) if d in rev else slice(None) for d in range(m.ndim))]


def flipud(m):
    return flip(m, 0)


def fliplr(m):
    return flip(m, 1)


def _rotate(x, shift, axis):
    """Elements move ``shift`` places towards higher indices along ``axis``, wrapping around."""
    from ._views import concatenate

    n = x.shape[axis]
    cut = (-shift) % n if n else 0
    if cut == 0:
        return x
    return concatenate([x[_along(x.ndim, axis, slice(cut, None))], x[_along(x.ndim, axis, slice(None, cut))]], axis=axis)


def roll(array, shift, axis=None):
    """``np.roll``.  ``axis=None`` rolls the FLATTENED array, which needs ``ravel`` / ``reshape`` (outside
    the hot path): supported for 1-D arrays only."""
    from ._collection import asarray

    out = asarray(array)
    if axis is None:
        if not isinstance(shift, Integral):
            raise TypeError("Expect `shift` to be an instance of Integral when `axis` is None.")
        if out.ndim != 1:
            raise NotImplementedError("roll with axis=None flattens the array (reshape is outside the B200 hot path)")
        return _rotate(out, shift, 0)
    shifts, axes = _as_tuple(shift), _as_tuple(axis)
    if len(shifts) != len(axes):
        raise ValueError("Must have the same number of shifts as axes.")
    for s, ax in zip(shifts, axes):
        out = _rotate(out, s, ax % out.ndim)
    return out


def diff(a, n=1, axis=-1, prepend=None, append=None):
    """``np.diff``: ``n`` times ``x[1:] - x[:-1]`` along ``axis``, after optional ``prepend`` / ``append``."""
    from ._collection import asarray
    from ._views import broadcast_to, concatenate

    a = asarray(a)
    n, axis = int(n), int(axis)
    if n < 0:
        raise ValueError(f"order must be non-negative but got {n}")
    if n == 0:
        return a

    def edge(v):
        v = asarray(v)
        if v.ndim:
            return v
        shape = tuple(1 if d == axis % a.ndim else s for d, s in enumerate(a.shape))
        return broadcast_to(v, shape)

    parts = ([edge(prepend)] if prepend is not None else []) + [a] + ([edge(append)] if append is not None else [])
    out = concatenate(parts, axis) if len(parts) > 1 else a
    upper, lower = _along(out.ndim, axis, slice(1, None)), _along(out.ndim, axis, slice(None, -1))
    for _ in range(n):
        out = out[upper] - out[lower]
    return out
