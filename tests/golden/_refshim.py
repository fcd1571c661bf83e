"""Import shim that lets the *unmodified* reference sources under /root/reference
run in this container, where the third-party packages ``dask`` and ``toolz`` are
absent.

TEST INFRASTRUCTURE ONLY.  It is used by ``tests/golden/generate.py`` (run by hand
in the build container, never on the GPU box) to execute the reference's own
chunk / combine / aggregate / planner functions and record their outputs as golden
fixtures.  Nothing in the product imports it.

How it works
------------
* ``dask_array`` is registered as a namespace-like package whose ``__path__`` is the
  reference directory, so ``import dask_array.reductions._common`` executes the real
  file but skips ``dask_array/__init__.py`` (which needs the full ``dask``).
* ``dask``, ``toolz`` and ``tlz`` are stub packages.  The handful of helpers the
  hot-path functions call at run time (``Dispatch``, ``deepmap``, ``partition_all``,
  ``lol_tuples``, ``cached_cumsum`` ...) are small restatements of the published
  behaviour of dask 2025.12 / toolz 1.1 (versions pinned by the reference's uv.lock).
  Every other attribute resolves to an inert ``_Stub`` class that can be subclassed,
  called or used as a decorator, which is enough for class bodies to be defined.
"""
from __future__ import annotations

import functools
import importlib.abc
import importlib.machinery
import inspect
import itertools
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub


class _Stub(metaclass=_StubMeta):
    """Inert placeholder: subclassable, callable, usable as a decorator."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Stub()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub()

    def __iter__(self):
        return iter(())

    def __class_getitem__(cls, item):
        return cls


# --------------------------------------------------------------------------- toolz
def partition_all(n, seq):
    it = iter(seq)
    while True:
        chunk = tuple(itertools.islice(it, n))
        if not chunk:
            return
        yield chunk


def compose(*funcs):
    if not funcs:
        return lambda x: x
    if len(funcs) == 1:
        return funcs[0]

    def composed(*a, **k):
        out = funcs[-1](*a, **k)
        for f in reversed(funcs[:-1]):
            out = f(out)
        return out

    return composed


def get(ind, seq, default="__no__default__"):
    if isinstance(ind, list):
        return tuple(seq[i] for i in ind)
    return seq[ind]


def accumulate(binop, seq, initial="__no__default__"):
    seq = iter(seq)
    if initial == "__no__default__":
        try:
            result = next(seq)
        except StopIteration:
            return
    else:
        result = initial
    yield result
    for elem in seq:
        result = binop(result, elem)
        yield result


def pluck(ind, seqs, default="__no__default__"):
    if isinstance(ind, list):
        return (tuple(s[i] for i in ind) for s in seqs)
    return (s[ind] for s in seqs)


def first(seq):
    return next(iter(seq))


def concat(seqs):
    return itertools.chain.from_iterable(seqs)


def frequencies(seq):
    d = {}
    for x in seq:
        d[x] = d.get(x, 0) + 1
    return d


def curry(f, *a, **k):
    return functools.partial(f, *a, **k) if (a or k) else f


def unique(seq, key=None):
    seen = set()
    for x in seq:
        v = x if key is None else key(x)
        if v not in seen:
            seen.add(v)
            yield x


def merge(*dicts, **kw):
    if len(dicts) == 1 and not isinstance(dicts[0], dict):
        dicts = dicts[0]
    out = {}
    for d in dicts:
        out.update(d)
    return out


def partition(n, seq):
    it = iter(seq)
    return zip(*([it] * n))


def groupby(key, seq):
    fn = key if callable(key) else (lambda x, _k=key: x[_k])
    out = {}
    for x in seq:
        out.setdefault(fn(x), []).append(x)
    return out


def valmap(func, d):
    return {k: func(v) for k, v in d.items()}


def broadcast_dimensions(argpairs, numblocks, sentinels=(1, (1,)), consolidate=None):
    """dask.blockwise.broadcast_dimensions (dask 2025.12): per index label, the set of block
    dimensions of every operand carrying it, minus the broadcast sentinels, then consolidated."""
    pairs = [(a, ind) for a, ind in argpairs if ind is not None]
    g = {}
    for name, inds in pairs:
        if name in numblocks:
            for i, d in zip(inds, numblocks[name]):
                g.setdefault(i, set()).add(d)
    g2 = {k: (v - set(sentinels) if len(v) > 1 else v) for k, v in g.items()}
    if consolidate:
        return valmap(consolidate, g2)
    if g2 and not set(map(len, g2.values())) == {1}:
        raise ValueError(f"Shapes do not align {g}")
    return valmap(first, g2)


_TOOLZ = dict(
    partition=partition, groupby=groupby, valmap=valmap,
    partition_all=partition_all, compose=compose, get=get, accumulate=accumulate,
    pluck=pluck, first=first, concat=concat, frequencies=frequencies, curry=curry,
    unique=unique, merge=merge, identity=lambda x: x,
    partial=functools.partial, reduce=functools.reduce,
    memoize=lambda f=None, **k: (f if f is not None else (lambda g: g)),
)


# --------------------------------------------------------------------------- dask.utils
class Dispatch:
    """Type-keyed single dispatch with MRO lookup (dask.utils.Dispatch)."""

    def __init__(self, name=None):
        self._lookup = {}
        self._lazy = {}
        if name:
            self.__name__ = name

    def register(self, type, func=None):
        def wrapper(func):
            if isinstance(type, tuple):
                for t in type:
                    self.register(t, func)
            else:
                self._lookup[type] = func
            return func

        return wrapper(func) if func is not None else wrapper

    def register_lazy(self, toplevel, func=None):
        def wrapper(func):
            self._lazy[toplevel] = func
            return func

        return wrapper(func) if func is not None else wrapper

    def dispatch(self, cls):
        lk = self._lookup
        for cls2 in cls.__mro__:
            if cls2 in lk:
                return lk[cls2]
        raise TypeError(f"No dispatch for {cls}")

    def __call__(self, arg, *args, **kwargs):
        return self.dispatch(type(arg))(arg, *args, **kwargs)


def deepmap(func, *seqs):
    if isinstance(seqs[0], (list, Iterator_)):
        return [deepmap(func, *items) for items in zip(*seqs)]
    return func(*seqs)


from collections.abc import Iterator as Iterator_  # noqa: E402


def derived_from(original_klass, version=None, ua_args=None, skipblocks=0, inconsistencies=None):
    def wrapper(method):
        return method

    return wrapper


def funcname(func):
    while isinstance(func, functools.partial):
        func = func.func
    name = getattr(func, "__name__", None) or type(func).__name__
    return name[:50]


def getargspec(func):
    if isinstance(func, functools.partial):
        return getargspec(func.func)
    return inspect.getfullargspec(func)


def has_keyword(func, keyword):
    try:
        return keyword in inspect.signature(func).parameters
    except Exception:
        return False


def is_arraylike(x):
    return bool(
        hasattr(x, "shape") and isinstance(x.shape, tuple) and hasattr(x, "dtype")
        and "Dask" not in type(x).__name__
    ) and not isinstance(x, type)


def cached_cumsum(seq, initial_zero=False):
    out = list(itertools.accumulate(seq))
    if initial_zero:
        out = [0] + out
    return tuple(out)


def parse_bytes(s):
    if isinstance(s, (int, float)):
        return int(s)
    s = s.replace(" ", "")
    units = {"kib": 2**10, "mib": 2**20, "gib": 2**30, "tib": 2**40, "kb": 10**3, "mb": 10**6,
             "gb": 10**9, "tb": 10**12, "b": 1, "": 1}
    i = len(s)
    while i and not s[i - 1].isdigit():
        i -= 1
    return int(float(s[:i] or 1) * units[s[i:].lower()])


def ndimlist(seq):
    if not isinstance(seq, (list, tuple)):
        return 0
    if len(seq) == 0:
        return 1
    return 1 + ndimlist(seq[0])


def concrete(seq):
    if isinstance(seq, Iterator_):
        seq = list(seq)
    if isinstance(seq, (tuple, list)):
        seq = list(map(concrete, seq))
    return seq


def _typename(typ):
    """dask.utils.typename: ``module.qualname`` of a type (of the instance's type when given an instance)."""
    if not isinstance(typ, type):
        return _typename(type(typ))
    mod = getattr(typ, "__module__", None)
    if not mod or mod == "builtins":
        return typ.__name__
    return f"{mod}.{typ.__qualname__}"


_UTILS = dict(
    ndimlist=ndimlist, concrete=concrete,
    Dispatch=Dispatch, deepmap=deepmap, derived_from=derived_from, funcname=funcname,
    getargspec=getargspec, has_keyword=has_keyword, is_arraylike=is_arraylike,
    cached_cumsum=cached_cumsum, parse_bytes=parse_bytes, cached_property=functools.cached_property,
    is_cupy_type=lambda x: False, is_series_like=lambda x: False, is_dataframe_like=lambda x: False,
    is_index_like=lambda x: False, format_bytes=lambda n: f"{n} B",
    typename=lambda t, short=False: _typename(t),
)


# --------------------------------------------------------------------------- dask.config
class _Config:
    """dask.config with the reference's defaults (dask_array/__init__.py:21-29 and the
    dask 2025.12 array schema)."""

    values = {
        "array.chunk-size": "128MiB",
        "array.chunk-size-tolerance": 1.25,       # dask 2025.12 default (quoted in dask_array/_shuffle.py:97)
        "array.rechunk.threshold": 32,   # raised by dask_array/__init__.py:21-29
        "array.rechunk.method": "tasks",
        "array.slicing.split-large-chunks": None,
        "array.optimize-graph": True,
        "array.unify-chunks-policy": "auto",      # dask_array/__init__.py:14-29
        "array.unify-chunks-limit": "512 MiB",
    }

    def get(self, key, default="__no__default__"):
        if key in self.values:
            return self.values[key]
        if default == "__no__default__":
            raise KeyError(key)
        return default

    class set:  # context manager / direct setter
        def __init__(self, arg=None, **kw):
            self.old = dict(_Config.values)
            _Config.values.update(arg or {})
            _Config.values.update({k.replace("__", "."): v for k, v in kw.items()})

        def __enter__(self):
            return self

        def __exit__(self, *a):
            _Config.values.clear()
            _Config.values.update(self.old)


# --------------------------------------------------------------------------- dask.core / blockwise
def flatten(seq, container=list):
    if isinstance(seq, str):
        yield seq
    else:
        for item in seq:
            if isinstance(item, container):
                yield from flatten(item, container=container)
            else:
                yield item


def lol_tuples(head, ind, values, dummies):
    """dask.blockwise.lol_tuples: nested lists of keys, one nesting level per dummy index."""
    if not ind:
        return head
    if ind[0] not in dummies:
        return lol_tuples(head + (values[ind[0]],), ind[1:], values, dummies)
    return [lol_tuples(head + (v,), ind[1:], values, dummies) for v in dummies[ind[0]]]


def _make_module(name, attrs=None, is_pkg=True):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs or {})
    if is_pkg:
        mod.__path__ = []

    def _getattr(attr, _name=name, _cache={}):
        if attr.startswith("__"):
            raise AttributeError(attr)
        # a distinct placeholder class per symbol keeps multiple-inheritance MROs consistent
        if attr not in _cache:
            _cache[attr] = _StubMeta(attr, (_Stub,), {"__module__": _name})
        return _cache[attr]

    mod.__getattr__ = _getattr
    return mod


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("dask", "toolz", "tlz")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        name = spec.name
        attrs = {}
        root = name.split(".")[0]
        if root in ("toolz", "tlz"):
            attrs = dict(_TOOLZ)
        elif name == "dask.utils":
            attrs = {k: v for k, v in _UTILS.items() if v is not None}
        elif name == "dask":
            attrs = {"config": _Config(), "is_dask_collection": lambda x: False}
        elif name == "dask.config":
            c = _Config()
            attrs = {"get": c.get, "set": _Config.set}
        elif name == "dask.core":
            attrs = {"flatten": flatten}
        elif name == "dask.blockwise":
            attrs = {"lol_tuples": lol_tuples, "broadcast_dimensions": broadcast_dimensions}
        elif name == "dask.base":
            attrs = {"is_dask_collection": lambda x: False, "tokenize": lambda *a, **k: "token"}
        elif name == "dask.tokenize":
            attrs = {"_tokenize_deterministic": lambda *a, **k: "token", "tokenize": lambda *a, **k: "token"}
        return _make_module(name, attrs)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Install the stubs and the ``dask_array`` path package.  Idempotent."""
    global _installed
    if _installed:
        return
    sys.meta_path.insert(0, _StubFinder())
    pkg = types.ModuleType("dask_array")
    pkg.__path__ = [REFERENCE_ROOT + "/dask_array"]
    pkg.__version__ = "reference"
    sys.modules["dask_array"] = pkg
    # sub-packages whose __init__ pulls in far more than the hot path: register them
    # as bare path packages too so that only the requested leaf modules execute.
    for sub in ("reductions", "linalg", "core", "creation", "slicing", "manipulation", "io",
                "stacking", "random", "routines"):
        m = types.ModuleType(f"dask_array.{sub}")
        m.__path__ = [f"{REFERENCE_ROOT}/dask_array/{sub}"]
        sys.modules[f"dask_array.{sub}"] = m
        setattr(pkg, sub, m)
    _installed = True
