"""Routines that are pure compositions of basic slices, concatenation and element-wise operators in the
reference -- ``roll`` (``manipulation/_roll.py:8-77``), ``diff`` (``routines/_diff.py:6-80``), ``flip`` /
``flipud`` / ``fliplr`` (``manipulation/_flip.py:10-55``).  They are composed the same way here, so they
inherit the kernels, the chunk unification and the multi-GPU paths of their parts; the shifted operands of
``diff`` are read through zero-copy views and fused into one kernel per device.
"""
from __future__ import annotations

from collections.abc import Iterable
from numbers import Integral


def flip(m, axis=None):
    """``flip`` (``manipulation/_flip.py:10-33``)."""
    from ._collection import asarray

    m = asarray(m)
    sl = m.ndim * [slice(None)]
    if axis is None:
        axis = range(m.ndim)
    if not isinstance(axis, Iterable):
        axis = (axis,)
    try:
        for ax in axis:
            sl[ax] = slice(None, None, -1)
    except IndexError as e:
        raise ValueError(f"`axis` of {axis} invalid for {m.ndim}-D array") from e
    return m[tuple(sl)]


def flipud(m):
    return flip(m, 0)


def fliplr(m):
    return flip(m, 1)


def roll(array, shift, axis=None):
    """``roll`` (``manipulation/_roll.py:8-77``).  ``axis=None`` (roll of the flattened array) needs
    ``ravel`` / ``reshape`` and is supported for 1-D arrays only."""
    from ._collection import asarray
    from ._views import concatenate

    result = asarray(array)
    if axis is None:
        if result.ndim != 1:
            raise NotImplementedError("roll with axis=None flattens the array (reshape is outside the B200 hot path)")
        if not isinstance(shift, Integral):
            raise TypeError("Expect `shift` to be an instance of Integral when `axis` is None.")
        shift, axis = (shift,), (0,)
    else:
        shift = tuple(shift) if isinstance(shift, Iterable) else (shift,)
        axis = tuple(axis) if isinstance(axis, Iterable) else (axis,)
    if len(shift) != len(axis):
        raise ValueError("Must have the same number of shifts as axes.")
    for i, s in zip(axis, shift):
        n = result.shape[i]
        s = 0 if n == 0 else -s % n
        if s == 0:
            continue
        sl1, sl2 = result.ndim * [slice(None)], result.ndim * [slice(None)]
        sl1[i], sl2[i] = slice(s, None), slice(None, s)
        result = concatenate([result[tuple(sl1)], result[tuple(sl2)]], axis=i)
    return result


def diff(a, n=1, axis=-1, prepend=None, append=None):
    """``diff`` (``routines/_diff.py:6-80``)."""
    from ._collection import asarray
    from ._views import broadcast_to, concatenate

    a = asarray(a)
    n, axis = int(n), int(axis)
    if n == 0:
        return a
    if n < 0:
        raise ValueError(f"order must be non-negative but got {n}")
    parts = []
    if prepend is not None:
        p = asarray(prepend)
        if p.ndim == 0:
            shape = list(a.shape)
            shape[axis] = 1
            p = broadcast_to(p, tuple(shape))
        parts.append(p)
    parts.append(a)
    if append is not None:
        q = asarray(append)
        if q.ndim == 0:
            shape = list(a.shape)
            shape[axis] = 1
            q = broadcast_to(q, tuple(shape))
        parts.append(q)
    if len(parts) > 1:
        a = concatenate(parts, axis)
    sl_1, sl_2 = a.ndim * [slice(None)], a.ndim * [slice(None)]
    sl_1[axis], sl_2[axis] = slice(1, None), slice(None, -1)
    r = a
    for _ in range(n):
        r = r[tuple(sl_1)] - r[tuple(sl_2)]
    return r
