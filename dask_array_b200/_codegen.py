"""CUDA source generator for fused element-wise chains (+ optional reduction epilogue).

The reference runs a ``FusedBlockwise`` group as a Python loop over NumPy calls, one
full-size temporary per operator (``dask_array/_blockwise.py:1697-1728`` ->
``dask._task_spec._execute_subgraph``).  Here the same group becomes ONE kernel: a
``Program`` (SSA list of NumPy-named operators over typed inputs) is rendered to a
``Chain`` functor and handed to the templates in ``csrc/b2_device.cuh``; temporaries live
in registers.

dtype rules follow ``Elemwise._info`` (``_blockwise.py:928-966``): the result dtype of every
operator is whatever NumPy returns for 1-element dummies plus the raw Python scalars
(NEP 50 weak promotion), so ``f4 * 2 -> f4``, ``i4 / i4 -> f8``, ``i4 ** 2 -> i4``,
comparisons -> bool.  Arithmetic is emitted in that result type with operands cast to it,
which is what NumPy's ufunc loops do.  Float contraction is disabled at compile time
(``--fmad=false`` in b2_jit_compile) so ``+ - * /`` and ``sqrt`` are bit-identical to NumPy.
"""
from __future__ import annotations

import hashlib
import math
import operator
from dataclasses import dataclass

import numpy as np

from . import _lib

MAX_INPUTS = _lib.B2_MAX_IN   # array inputs of one fused kernel (descriptor table slots)
CODEGEN_VERSION = "22"     # part of every kernel's cache key: bump when generated code changes

CTYPE = {
    "bool": "bool", "int8": "signed char", "uint8": "unsigned char", "int16": "short",
    "uint16": "unsigned short", "int32": "int", "uint32": "unsigned int", "int64": "long long",
    "uint64": "unsigned long long", "float32": "float", "float64": "double",
}


def ctype(dt) -> str:
    try:
        return CTYPE[np.dtype(dt).name]
    except KeyError:
        raise NotImplementedError(f"dtype {np.dtype(dt)} has no B200 kernel type") from None


def literal(value, dt) -> str:
    """Exact C++ literal of ``value`` cast to dtype ``dt``."""
    dt = np.dtype(dt)
    with np.errstate(all="ignore"):
        v = np.asarray(value).astype(dt)[()]
    if dt == np.bool_:
        return "true" if bool(v) else "false"
    if dt.kind in "iu":
        iv = int(v)
        if dt == np.int64:
            return "(-9223372036854775807LL-1)" if iv == -(2**63) else f"{iv}LL"
        if dt == np.uint64:
            return f"{iv}ULL"
        return f"(({ctype(dt)}){iv})"
    if dt == np.float32:
        if math.isnan(v):
            return "__int_as_float(0x7fc00000)"
        if math.isinf(v):
            return "__int_as_float(0x7f800000)" if v > 0 else "__int_as_float(0xff800000)"
        return float(v).hex() + "f"
    if dt == np.float64:
        if math.isnan(v):
            return "__longlong_as_double(0x7ff8000000000000LL)"
        if math.isinf(v):
            return "__longlong_as_double(0x7ff0000000000000LL)" if v > 0 else "__longlong_as_double(0xfff0000000000000LL)"
        return float(v).hex()
    raise NotImplementedError(f"literal of dtype {dt}")


@dataclass(frozen=True)
class Ref:
    kind: str          # "in" | "tmp" | "const"
    index: int         # input / tmp number (const: -1)
    dtype: object      # np.dtype, or None for a weak Python scalar
    value: object = None

    @property
    def weak(self) -> bool:
        return self.kind == "const" and self.dtype is None


# ----------------------------------------------------------------------------- op table
_UNARY_MATH = {  # NumPy name -> (fp32 function, fp64 function)
    "sin": ("sinf", "sin"), "cos": ("cosf", "cos"), "tan": ("tanf", "tan"),
    "arcsin": ("asinf", "asin"), "arccos": ("acosf", "acos"), "arctan": ("atanf", "atan"),
    "sinh": ("sinhf", "sinh"), "cosh": ("coshf", "cosh"), "tanh": ("tanhf", "tanh"),
    "arcsinh": ("asinhf", "asinh"), "arccosh": ("acoshf", "acosh"), "arctanh": ("atanhf", "atanh"),
    "exp": ("expf", "exp"), "exp2": ("exp2f", "exp2"), "expm1": ("expm1f", "expm1"),
    "log": ("logf", "log"), "log2": ("log2f", "log2"), "log10": ("log10f", "log10"),
    "log1p": ("log1pf", "log1p"), "sqrt": ("sqrtf", "sqrt"), "cbrt": ("cbrtf", "cbrt"),
    "floor": ("floorf", "floor"), "ceil": ("ceilf", "ceil"), "trunc": ("truncf", "trunc"),
    "rint": ("rintf", "rint"), "fabs": ("fabsf", "fabs"),
}
_BINARY_MATH = {
    "arctan2": ("atan2f", "atan2"), "hypot": ("hypotf", "hypot"), "copysign": ("copysignf", "copysign"),
    "nextafter": ("nextafterf", "nextafter"), "fmod": ("fmodf", "fmod"),
    "fmax": ("fmaxf", "fmax"), "fmin": ("fminf", "fmin"),
}
_CMP = {"less": "<", "less_equal": "<=", "greater": ">", "greater_equal": ">=", "equal": "==", "not_equal": "!="}
_ARITH = {"add": "+", "subtract": "-", "multiply": "*"}
_BITWISE = {"bitwise_and": "&", "bitwise_or": "|", "bitwise_xor": "^"}
_LOGICAL = {"logical_and": "&&", "logical_or": "||"}

# Python operator / alias -> canonical NumPy ufunc name
ALIASES = {
    operator.add: "add", operator.sub: "subtract", operator.mul: "multiply",
    operator.truediv: "true_divide", operator.floordiv: "floor_divide", operator.pow: "power",
    operator.mod: "remainder", operator.neg: "negative", operator.pos: "positive",
    operator.abs: "absolute", operator.invert: "invert", operator.and_: "bitwise_and",
    operator.or_: "bitwise_or", operator.xor: "bitwise_xor", operator.lshift: "left_shift",
    operator.rshift: "right_shift", operator.lt: "less", operator.le: "less_equal",
    operator.gt: "greater", operator.ge: "greater_equal", operator.eq: "equal", operator.ne: "not_equal",
    "divide": "true_divide", "mod": "remainder", "abs": "absolute", "bitwise_not": "invert",
    "conjugate": "positive", "conj": "positive",
    # the reference's own element-wise helpers (reductions/_common.py:625-636: std = safe_sqrt(var))
    "_sqrt": "sqrt", "safe_sqrt": "sqrt",
}


def canonical_name(op) -> str:
    """Name of an Elemwise ``op`` (operator.* function, np.ufunc or string)."""
    try:
        if op in ALIASES:
            return ALIASES[op]
    except TypeError:           # unhashable callable object
        pass
    name = op if isinstance(op, str) else getattr(op, "__name__", None)
    if name in ALIASES:
        return ALIASES[name]
    if name is None:
        raise NotImplementedError(f"cannot fuse {op!r} into a B200 kernel")
    return name


def _np_func(name):
    if name == "where":
        return np.where
    f = getattr(np, name, None)
    if f is None:
        raise NotImplementedError(f"element-wise operator {name!r} has no B200 code generator")
    return f


def infer_dtype(name, args, **kw):
    """Result dtype the way Elemwise._info does it (``_blockwise.py:941-946``)."""
    if name == "astype":
        return np.dtype(kw["dtype"])
    if name in ("frexp_mantissa", "frexp_exponent"):
        dummy = np.empty((1,), dtype=args[0].dtype) if not args[0].weak else args[0].value
        with np.errstate(all="ignore"):
            m, e = np.frexp(dummy)
        return np.asarray(m if name == "frexp_mantissa" else e).dtype
    dummies = [np.empty((1,), dtype=a.dtype) if not a.weak else a.value for a in args]
    with np.errstate(all="ignore"):
        return np.asarray(_np_func(name)(*dummies)).dtype


class Program:
    """SSA program of one fused chain."""

    def __init__(self):
        self.inputs: list[np.dtype] = []
        self.ops: list[tuple] = []       # (name, args, out_dtype, kwargs)
        self.output: Ref | None = None

    # ---- construction
    def add_input(self, dtype) -> Ref:
        dt = np.dtype(dtype)
        ctype(dt)
        if len(self.inputs) >= _lib.B2_MAX_IN:
            raise NotImplementedError(f"a fused kernel takes at most {_lib.B2_MAX_IN} array inputs")
        self.inputs.append(dt)
        return Ref("in", len(self.inputs) - 1, dt)

    def const(self, value) -> Ref:
        if isinstance(value, (np.generic, np.ndarray)):
            arr = np.asarray(value)
            if arr.ndim != 0:
                raise NotImplementedError("only 0-d constants can be baked into a kernel")
            return Ref("const", -1, arr.dtype, arr[()])
        if isinstance(value, (bool, int, float)):
            return Ref("const", -1, None, value)
        raise NotImplementedError(f"constant of type {type(value).__name__}")

    def typed_const(self, value, dtype) -> Ref:
        return Ref("const", -1, np.dtype(dtype), np.asarray(value).astype(dtype)[()])

    def op(self, name, *args: Ref, **kw) -> Ref:
        name = canonical_name(name)
        out = infer_dtype(name, args, **kw)
        ctype(out)
        _emit(name, list(args), out, kw)   # validate now, fail loudly early
        self.ops.append((name, tuple(args), out, tuple(sorted(kw.items()))))
        return Ref("tmp", len(self.ops) - 1, out)

    def set_output(self, ref: Ref) -> None:
        if ref.weak:
            raise NotImplementedError("chain output cannot be a bare Python scalar")
        self.output = ref

    @property
    def out_dtype(self) -> np.dtype:
        return np.dtype(self.output.dtype)

    # ---- identity / rendering
    def key(self) -> str:
        parts = ["in:" + ",".join(d.name for d in self.inputs)]
        for name, args, out, kw in self.ops:
            a = ";".join(_ref_key(r) for r in args)
            parts.append(f"{name}({a})->{out.name}{kw if kw else ''}")
        parts.append("out:" + _ref_key(self.output))
        return "|".join(parts)

    _PACKED_OPS = {"sin", "cos", "add", "subtract", "multiply", "negative", "positive", "absolute", "square"}

    def packable(self) -> bool:
        """fp32-only chain whose operators all have a packed f32x2 form (FFMA2 family)."""
        f4 = np.dtype(np.float32)
        if not self.ops or any(d != f4 for d in self.inputs) or self.out_dtype != f4:
            return False
        for name, args, out, kw in self.ops:
            if out != f4:
                return False
            if name == "power":
                if not (args[1].kind == "const" and args[1].value == 2):
                    return False
            elif name not in self._PACKED_OPS:
                return False
            for r in args:
                if r.kind != "const" and np.dtype(r.dtype) != f4:
                    return False
        return True

    def body_packed(self, contract: bool) -> list[str]:
        """Statements computing the pair (o[v], o[v+1]) with packed f32x2 arithmetic.  ``contract``:
        fold a single-use multiply / square into the add that consumes it (FFMA2) -- only for kernels
        whose chain is observed through a reduction; element-wise kernels keep NumPy's two roundings."""
        f4 = np.float32
        uses = {}
        for name, args, out, kw in self.ops:
            for r in args:
                if r.kind == "tmp":
                    uses[r.index] = uses.get(r.index, 0) + 1
        if self.output.kind == "tmp":
            uses[self.output.index] = uses.get(self.output.index, 0) + 1

        def ex(r):
            if r.kind == "in":
                return f"x{r.index}"
            if r.kind == "tmp":
                return f"t{r.index}"
            return f"b2_bc({literal(r.value, f4)})"

        def as_product(r):
            """(a, b) if r is a single-use multiply / square / power-2 that may be contracted."""
            if not contract or r.kind != "tmp" or uses.get(r.index, 0) != 1:
                return None
            name, args, out, kw = self.ops[r.index]
            if name == "multiply":
                return ex(args[0]), ex(args[1])
            if name == "square" or name == "power":
                return ex(args[0]), ex(args[0])
            return None

        lines, skipped = [], set()
        for k in range(len(self.inputs)):
            lines.append(f"const b2f2 x{k} = b2_pk(g.a{k}[v], g.a{k}[v + 1]);")
        # decide contractions first so that the folded products are not emitted
        folded = {}
        for j, (name, args, out, kw) in enumerate(self.ops):
            if name in ("add", "subtract"):
                for pos in ((1, 0) if name == "add" else (1,)):     # a + (b*c), (b*c) + a, a - ... only a*b - c handled below
                    pr = as_product(args[pos]) if args[pos].kind == "tmp" and args[pos].index not in skipped else None
                    if pr and name == "add":
                        folded[j] = (pr, args[1 - pos])
                        skipped.add(args[pos].index)
                        break
        for j, (name, args, out, kw) in enumerate(self.ops):
            if j in skipped:
                continue
            if j in folded:
                (pa, pb), other = folded[j]
                expr = f"b2_fma2({pa}, {pb}, {ex(other)})"
            elif name in ("sin", "cos"):
                expr = f"b2_{name}f_fast2({ex(args[0])}, big)"
            elif name == "add":
                expr = f"b2_add2({ex(args[0])}, {ex(args[1])})"
            elif name == "subtract":
                expr = f"b2_sub2({ex(args[0])}, {ex(args[1])})"
            elif name == "multiply":
                expr = f"b2_mul2({ex(args[0])}, {ex(args[1])})"
            elif name in ("square", "power"):
                expr = f"b2_mul2({ex(args[0])}, {ex(args[0])})"
            elif name == "negative":
                expr = f"b2_neg2({ex(args[0])})"
            elif name == "absolute":
                expr = f"b2_abs2({ex(args[0])})"
            else:   # positive
                expr = ex(args[0])
            lines.append(f"const b2f2 t{j} = {expr};")
        lines.append(f"b2_upk({ex(self.output)}, o[v], o[v + 1]);")
        return lines

    def uses_fast_sincos(self) -> bool:
        return any(name in ("sin", "cos") and out == np.float32 for name, _, out, _ in self.ops)

    def body(self, fast: bool = False, slow_calls: bool = False) -> list[str]:
        """C++ statements computing ``o[v]`` from ``g.a<k>[v]``.  ``fast``: fp32 sin/cos use
        b2_sinf_fast/b2_cosf_fast and track max|arg| in ``big`` (the caller redoes the vector
        with libdevice when an argument leaves the Cody-Waite range)."""
        lines = []
        for k, dt in enumerate(self.inputs):
            lines.append(f"const {ctype(dt)} x{k} = g.a{k}[v];")
        for j, (name, args, out, kw) in enumerate(self.ops):
            if fast and name in ("sin", "cos") and out == np.float32:
                expr = f"b2_{name}f_fast({_typed_expr(args[0], out)}, big)"
            elif slow_calls and name in ("sin", "cos") and out == np.float32:
                expr = f"b2_{name}f_slow({_typed_expr(args[0], out)})"
            else:
                expr = _emit(name, list(args), out, dict(kw))
            lines.append(f"const {ctype(out)} t{j} = {expr};")
        lines.append(f"o[v] = {_typed_expr(self.output, self.out_dtype)};")
        return lines


def _ref_key(r: Ref) -> str:
    if r.kind == "const":
        dt = "weak" if r.dtype is None else np.dtype(r.dtype).name
        return f"c[{r.value!r}:{dt}]"
    return f"{r.kind}{r.index}"


def _ref_expr(r: Ref) -> str:
    if r.kind == "in":
        return f"x{r.index}"
    if r.kind == "tmp":
        return f"t{r.index}"
    return None  # constants are rendered where their target type is known


def _typed_expr(r: Ref, dt) -> str:
    """Expression of ``r`` converted to dtype ``dt`` (NumPy casting inside a ufunc loop)."""
    dt = np.dtype(dt)
    if r.kind == "const":
        return literal(r.value, dt)
    e = _ref_expr(r)
    if np.dtype(r.dtype) == dt:
        return e
    return f"(({ctype(dt)})({e}))"


def _common_type(args) -> np.dtype:
    vals = [a.value if a.weak else np.dtype(a.dtype) for a in args]
    return np.result_type(*vals)


def _emit(name, args, out, kw) -> str:
    """C++ expression for one operator applied to ``args`` (operands cast like a ufunc loop)."""
    out = np.dtype(out)
    T = ctype(out)

    def a(i, dt=out):
        return _typed_expr(args[i], dt)

    isf = out.kind == "f"
    fidx = 0 if out == np.float32 else 1

    if name == "astype":
        src = args[0]
        if out == np.int64 and not src.weak and src.kind != "const" and np.dtype(src.dtype).kind == "f":
            return f"b2_cast<long long>({_ref_expr(src)})"      # x86 NumPy semantics for non-finite values
        return a(0)
    if name in _ARITH:
        if out == np.bool_ and name == "subtract":
            raise NotImplementedError("numpy boolean subtract is not defined")
        return f"(({T})({a(0)} {_ARITH[name]} {a(1)}))"
    if name == "true_divide":
        return f"({a(0)} / {a(1)})"
    if name == "floor_divide":
        if isf:
            return f"b2_floordiv_f({a(0)}, {a(1)})"
        fn = "b2_floordiv_uint" if out.kind == "u" else "b2_floordiv_int"
        return f"{fn}<{T}>({a(0)}, {a(1)})"
    if name == "remainder":
        if isf:
            return f"b2_mod_f({a(0)}, {a(1)})"
        fn = "b2_mod_uint" if out.kind == "u" else "b2_mod_int"
        return f"{fn}<{T}>({a(0)}, {a(1)})"
    if name in ("power", "float_power"):
        e = args[1]
        if e.kind == "const" and name == "power":
            ev = e.value
            # NumPy's scalar-exponent fast paths (square / sqrt / reciprocal / identity)
            if ev == 2:
                return f"(({T})({a(0)} * {a(0)}))"
            if ev == 1:
                return a(0)
            if isf and ev == 0.5:
                return f"{_UNARY_MATH['sqrt'][fidx]}({a(0)})"
            if isf and ev == -1:
                return f"({literal(1, out)} / {a(0)})"
        if isf:
            return f"{'powf' if out == np.float32 else 'pow'}({a(0)}, {a(1)})"
        return f"b2_ipow<{T}>({a(0)}, (i64)({a(1, np.int64) if out.kind == 'i' else a(1)}))"
    if name == "negative":
        return f"(({T})(-{a(0)}))"
    if name == "positive":
        return a(0)
    if name == "absolute":
        if isf:
            return f"{_UNARY_MATH['fabs'][fidx]}({a(0)})"
        if out.kind == "u" or out == np.bool_:
            return a(0)
        return f"(({T})({a(0)} < 0 ? -{a(0)} : {a(0)}))"
    if name == "invert":
        return f"(!{a(0)})" if out == np.bool_ else f"(({T})(~{a(0)}))"
    if name in _BITWISE:
        return f"(({T})({a(0)} {_BITWISE[name]} {a(1)}))"
    if name in ("left_shift", "right_shift"):
        bits = out.itemsize * 8
        sh = a(1)
        if name == "left_shift":
            return f"(({sh}) >= {bits} || ({sh}) < 0 ? ({T})0 : ({T})({a(0)} << {sh}))"
        fill = f"(({a(0)}) < 0 ? ({T})-1 : ({T})0)" if out.kind == "i" else f"({T})0"
        return f"(({sh}) >= {bits} || ({sh}) < 0 ? {fill} : ({T})({a(0)} >> {sh}))"
    if name in _CMP:
        ct = _common_type(args)
        return f"({a(0, ct)} {_CMP[name]} {a(1, ct)})"
    if name in _LOGICAL:
        return f"(({a(0, np.bool_)}) {_LOGICAL[name]} ({a(1, np.bool_)}))"
    if name == "logical_xor":
        return f"(({a(0, np.bool_)}) != ({a(1, np.bool_)}))"
    if name == "logical_not":
        return f"(!({a(0, np.bool_)}))"
    if name in ("maximum", "minimum"):
        return f"b2_np_{'max' if name == 'maximum' else 'min'}<{T}>({a(0)}, {a(1)})"
    if name in _UNARY_MATH:
        if not isf:
            if name in ("floor", "ceil", "trunc", "rint"):
                return a(0)
            raise NotImplementedError(f"{name} on dtype {out}")
        return f"{_UNARY_MATH[name][fidx]}({a(0)})"
    if name == "fmod" and not isf:
        # C's % truncates toward zero like np.fmod; NumPy gives 0 for a zero divisor, and x % -1 is 0 (INT_MIN too)
        if out.kind == "u":
            return f"((({a(1)}) == 0) ? ({T})0 : ({T})(({a(0)}) % ({a(1)})))"
        return f"((({a(1)}) == 0 || ({a(1)}) == ({T})-1) ? ({T})0 : ({T})(({a(0)}) % ({a(1)})))"
    if name in _BINARY_MATH:
        if not isf:
            raise NotImplementedError(f"{name} on dtype {out}")
        return f"{_BINARY_MATH[name][fidx]}({a(0)}, {a(1)})"
    if name == "square":
        return f"(({T})({a(0)} * {a(0)}))"
    if name == "reciprocal":
        return f"(({T})({literal(1, out)} / {a(0)}))"
    if name == "sign":
        return f"b2_sign<{T}>({a(0)})"
    if name in ("isnan", "isinf", "isfinite", "signbit"):
        src = np.dtype(args[0].dtype) if not args[0].weak else np.dtype(np.float64)
        if src.kind != "f":
            return {"isnan": "false", "isinf": "false", "isfinite": "true"}.get(name) or f"({a(0, src)} < 0)"
        return f"{name}({a(0, src)})"
    if name in ("deg2rad", "radians"):
        return f"(({T})({a(0)} * {literal(np.pi / 180.0, out)}))"
    if name in ("rad2deg", "degrees"):
        return f"(({T})({a(0)} * {literal(180.0 / np.pi, out)}))"
    if name in ("frexp_mantissa", "frexp_exponent"):
        # np.frexp (``_ufunc.py:429-436``, the two DoubleOutputs): x = m * 2**e with 0.5 <= |m| < 1; zero, inf and
        # NaN keep x as the mantissa and report exponent 0, like NumPy / glibc
        src = np.dtype(args[0].dtype) if not args[0].weak else np.dtype(np.float64)
        if src.kind != "f":
            src = np.dtype(np.float64)
        fx = "frexpf" if src == np.float32 else "frexp"
        xs = a(0, src)
        if name == "frexp_mantissa":
            return f"([&]{{ int e_ = 0; const {ctype(src)} m_ = {fx}({xs}, &e_); return isfinite({xs}) ? m_ : {xs}; }}())"
        return f"([&]{{ int e_ = 0; (void){fx}({xs}, &e_); return (isfinite({xs}) && {xs} != 0) ? e_ : 0; }}())"
    if name == "where":
        return f"(({a(0, np.bool_)}) ? {a(1)} : {a(2)})"
    if name == "clip":
        return f"b2_np_min<{T}>(b2_np_max<{T}>({a(0)}, {a(1)}), {a(2)})"
    if name == "logaddexp":
        f_log1p, f_exp, f_abs = (("log1pf", "expf", "fabsf") if out == np.float32 else ("log1p", "exp", "fabs"))
        return (f"(({a(0)}) == ({a(1)}) ? ({a(0)}) + {literal(math.log(2.0), out)} : "
                f"b2_np_max<{T}>({a(0)}, {a(1)}) + {f_log1p}({f_exp}(-{f_abs}(({a(0)}) - ({a(1)})))))")
    raise NotImplementedError(f"element-wise operator {name!r} has no B200 code generator")


# ----------------------------------------------------------------------------- kernels
_MODE_NAME = {_lib.MODE_EW: "B2M_EW", _lib.MODE_R: "B2M_R", _lib.MODE_C: "B2M_C", _lib.MODE_RC: "B2M_RC",
              _lib.MODE_SR: "B2M_SR", _lib.MODE_SC: "B2M_SC"}
_RED_NAME = {
    _lib.RED_NONE: "B2R_NONE", _lib.RED_SUM: "B2R_SUM", _lib.RED_MIN: "B2R_MIN", _lib.RED_MAX: "B2R_MAX",
    _lib.RED_ARGMIN: "B2R_ARGMIN", _lib.RED_ARGMAX: "B2R_ARGMAX", _lib.RED_MOMENT: "B2R_MOMENT",
    _lib.RED_PROD: "B2R_PROD", _lib.RED_ANY: "B2R_ANY", _lib.RED_ALL: "B2R_ALL",
    _lib.RED_NANMIN: "B2R_NANMIN", _lib.RED_NANMAX: "B2R_NANMAX",
}


@dataclass(frozen=True)
class KernelSpec:
    """Everything that is baked into one compiled kernel (also its cache key)."""
    program_key: str
    layouts: tuple       # per input: "V" contiguous vector, "S" column-stride 0, "G" general stride
    mode: int
    redop: int
    vec: int
    tx: int
    ty: int
    rpt: int
    unroll: int
    acc_dtype: str       # accumulator / output dtype name of SUM/PROD; working type of MOMENT
    variant: str = ""    # "sym": mirror-pair kernel for f(x, x.T) chains (b2_run_ewt_sym)

    def digest(self) -> str:
        import os

        tune = os.environ.get("B2_MINB", "")
        return hashlib.sha1((CODEGEN_VERSION + _header_hash() + tune + repr(self)).encode()).hexdigest()[:20]


_HEADER_HASH = None


def _header_hash() -> str:
    """Digest of the device header every generated kernel includes (``b2_device.cuh``, embedded in the
    library): a cached cubin can never outlive a change of the templates it was built from."""
    global _HEADER_HASH
    if _HEADER_HASH is None:
        _HEADER_HASH = hashlib.sha1(_lib.lib.b2_device_header()).hexdigest()[:12]
    return _HEADER_HASH


def packed_bytes(spec: KernelSpec, out_dtype) -> int:
    """sizeof(A::Packed) in b2_device.cuh for this reduction."""
    it = np.dtype(out_dtype).itemsize
    r = spec.redop
    if r in (_lib.RED_SUM, _lib.RED_PROD):
        return np.dtype(spec.acc_dtype).itemsize
    if r in (_lib.RED_MIN, _lib.RED_MAX, _lib.RED_NANMIN, _lib.RED_NANMAX):
        return max(it, 4) * 2 if it <= 4 else 16
    if r in (_lib.RED_ARGMIN, _lib.RED_ARGMAX):
        return 16
    if r == _lib.RED_MOMENT:
        return 24
    if r in (_lib.RED_ANY, _lib.RED_ALL):
        return 1
    return 0


def render(program: Program, spec: KernelSpec) -> str:
    """Full translation unit of one fused kernel (entry point ``b2_fused``)."""
    T = ctype(program.out_dtype)
    regs = [f"{ctype(dt)} a{k}[B2_V];" for k, dt in enumerate(program.inputs)] or ["char _unused;"]
    ptrs, setup_r, setup_c, loads, adv = [], [], [], [], []
    stage, loads_n, loads_s, toff = [], [], [], 0
    ewt = "T" in spec.layouts
    for k, (dt, lay) in enumerate(zip(program.inputs, spec.layouts)):
        ct = ctype(dt)
        if lay == "T":
            # staged through shared memory by b2_run_ewt (no pointer walking)
            ptrs.append(f"const {ct}* p{k}; i64 s{k};")
            setup_r.append(f"P.p{k} = nullptr; P.s{k} = 0;")
            setup_c.append(f"P.p{k} = nullptr; P.s{k} = 0;")
            loads.append(f"/* input {k} is staged */")
            adv.append("")
            stage.append(
                f"{{ const {ct}* src = (const {ct}*)blk.in[{k}] + b * blk.in_sb[{k}]; {ct}* sm = ({ct}*)(smem + {toff});\n"
                f"              const i64 sr = blk.in_sr[{k}], scs = blk.in_sc[{k}]; {ct} tmp[B2_TT * B2_TT / NTHR];\n"
                f"              _Pragma(\"unroll\") for (int m = 0; m < B2_TT * B2_TT / NTHR; ++m) {{ const int idx = tid + m * NTHR, ii = idx / B2_TT, jj = idx % B2_TT;\n"
                f"                  const i64 r = r0 + jj, c = c0 + ii; if (r < blk.R && c < blk.C) tmp[m] = b2_ld(src + r * sr + c * scs); }}\n"
                f"              _Pragma(\"unroll\") for (int m = 0; m < B2_TT * B2_TT / NTHR; ++m) {{ const int idx = tid + m * NTHR, ii = idx / B2_TT, jj = idx % B2_TT;\n"
                f"                  sm[jj * B2_TP + ii] = tmp[m]; }} }}")
            loads_s.append(f"{{ const {ct}* sm = (const {ct}*)(smem + {toff}) + lr * B2_TP + lc;\n"
                           f"              _Pragma(\"unroll\") for (int v = 0; v < B2_V; ++v) g.a{k}[v] = sm[v]; }}")
            toff += -(-64 * 65 * dt.itemsize // 16) * 16
            continue
        ptrs.append(f"const {ct}* p{k}; i64 s{k};" + (f" i64 c{k};" if lay == "G" else ""))
        colmul = {"V": "c", "S": "0", "G": f"c * blk.in_sc[{k}]"}[lay]
        base = f"(const {ct}*)blk.in[{k}] + b * blk.in_sb[{k}] + r * blk.in_sr[{k}] + {colmul}"
        extra = f" P.c{k} = blk.in_sc[{k}];" if lay == "G" else ""
        setup_r.append(f"P.p{k} = {base}; P.s{k} = rstep * blk.in_sr[{k}];{extra}")
        cstep = {"V": "cstep", "S": "0", "G": f"cstep * blk.in_sc[{k}]"}[lay]
        setup_c.append(f"P.p{k} = {base}; P.s{k} = {cstep};{extra}")
        if lay == "V":
            loads.append(f"b2_load_vec<{ct}, B2_V>(P.p{k} + k * P.s{k}, g.a{k});")
        elif lay == "S":
            loads.append(f"b2_load_bcast<{ct}, B2_V>(P.p{k} + k * P.s{k}, g.a{k});")
        else:
            loads.append(f"b2_load_strided<{ct}, B2_V>(P.p{k} + k * P.s{k}, P.c{k}, g.a{k});")
        adv.append(f"P.p{k} += n * P.s{k};")
        loads_n.append(loads[-1].replace("P.p%d + k * P.s%d" % (k, k), "P.p%d + (i64)lr * P.s%d" % (k, k)))
    if not ptrs:
        ptrs = ["char _unused;"]
    if ewt and toff > 40 * 1024:
        raise NotImplementedError("too many transposed operands for the shared-memory staged kernel")
    acc = ctype(spec.acc_dtype)
    nl = "\n            "
    fast = program.uses_fast_sincos()
    header = "// b2-options: fmad\n" if spec.mode != _lib.MODE_EW else ""
    packed = program.packable() and spec.vec % 2 == 0
    if packed:
        fast_body = nl.join(program.body_packed(contract=spec.mode != _lib.MODE_EW))
        fast_loop = f"for (int v = 0; v < B2_V; v += 2) {{{nl}{fast_body}\n        }}"
    else:
        fast_loop = f"for (int v = 0; v < B2_V; ++v) {{{nl}{nl.join(program.body(fast=fast))}\n        }}"
    compute = f"""    static constexpr bool HAS_SLOW = {'true' if fast else 'false'};
    // exact chain (libdevice transcendental functions, full argument range)
    __device__ __forceinline__ static void compute_slow(const Regs& g, const B2Scalars& sc, out_t (&o)[B2_V]) {{
#pragma unroll
        for (int v = 0; v < B2_V; ++v) {{
            {nl.join(program.body(slow_calls=fast))}
        }}
    }}
    __device__ __forceinline__ static void compute(const Regs& g, const B2Scalars& sc, out_t (&o)[B2_V], float& big) {{
#pragma unroll
        {fast_loop}
    }}"""
    sym = ""
    if spec.variant == "sym":
        kn, kt = spec.layouts.index("V"), spec.layouts.index("T")
        ct = ctype(program.inputs[kn])
        sym = f"""    typedef {ct} sym_t;
    __device__ __forceinline__ static void sym_put(sym_t* tile, int off, const Regs& g) {{ b2_store_vec<{ct}, B2_V>(tile + off, g.a{kn}); }}
    __device__ __forceinline__ static void sym_get(const sym_t* tile, int base, Regs& g) {{
#pragma unroll
        for (int v = 0; v < B2_V; ++v) g.a{kt}[v] = tile[base + v * {spec.rpt}];
    }}"""
        run = f"b2_run_ewt_sym<Chain, B2_V, {spec.rpt}>(blocks, nblocks, sc);"
    elif ewt:
        run = f"b2_run_ewt<Chain, B2_V, {spec.tx}, {spec.ty}>(blocks, nblocks, sc);"
    elif spec.mode in (_lib.MODE_SR, _lib.MODE_SC):
        if fast:
            raise NotImplementedError("cumulative scans take exact element-wise chains only")
        run = (f"b2_run_scan<Chain, {_MODE_NAME[spec.mode]}, {_RED_NAME[spec.redop]}, B2_V, {spec.tx}, {spec.ty}, "
               f"{spec.rpt}, {spec.unroll}, {acc}>(blocks, nblocks, sc);")
    else:
        run = (f"b2_run<Chain, {_MODE_NAME[spec.mode]}, {_RED_NAME[spec.redop]}, B2_V, {spec.tx}, {spec.ty}, "
               f"{spec.rpt}, {spec.unroll}, {acc}>(blocks, nblocks, sc);")
    return f"""{header}// generated by dask_array_b200/_codegen.py -- one FusedBlockwise expression
// program: {program.key()}
#include "b2_device.cuh"
#define B2_V {spec.vec}
struct Chain {{
    typedef {T} out_t;
    struct Regs {{ {' '.join(regs)} }};
    struct Ptrs {{ {' '.join(ptrs)} }};
    __device__ __forceinline__ static void setup_rows(const B2Block& blk, i64 b, i64 r, i64 c, i64 rstep, Ptrs& P) {{
        {(nl[:-4]).join(setup_r)}
    }}
    __device__ __forceinline__ static void setup_cols(const B2Block& blk, i64 b, i64 r, i64 c, i64 cstep, Ptrs& P) {{
        {(nl[:-4]).join(setup_c)}
    }}
    __device__ __forceinline__ static void load(const Ptrs& P, int k, Regs& g) {{
        {(nl[:-4]).join(loads)}
    }}
    __device__ __forceinline__ static void advance(Ptrs& P, int n) {{
        {(nl[:-4]).join(adv)}
    }}
    static constexpr int TBYTES = {max(toff, 16)};
    template <int NTHR>
    __device__ __forceinline__ static void stage(const B2Block& blk, i64 b, i64 r0, i64 c0, unsigned char* smem, int tid) {{
        {(nl[:-4]).join(stage)}
    }}
    __device__ __forceinline__ static void load_n(const Ptrs& P, int lr, Regs& g) {{
        {(nl[:-4]).join(loads_n)}
    }}
    __device__ __forceinline__ static void load_s(const unsigned char* smem, int lr, int lc, Regs& g) {{
        {(nl[:-4]).join(loads_s)}
    }}
{compute}
{sym}
}};
extern "C" __global__ void __launch_bounds__({spec.tx * spec.ty}, {_min_blocks(spec, packed)})
b2_fused(const B2Block* __restrict__ blocks, int nblocks, const B2Scalars sc) {{
    {run}
}}
"""


def _min_blocks(spec, packed: bool = False) -> int:
    import os

    if os.environ.get("B2_MINB"):
        return int(os.environ["B2_MINB"])
    if spec.tx * spec.ty > 256:
        return 1
    if spec.variant == "sym":
        return 5 if spec.vec == 4 else 4      # B200 sweep (c4): 5 CTAs/SM (51 regs) f4 5943 GB/s; 6 spills
    if packed:
        return 3
    if spec.mode == _lib.MODE_C and spec.redop != _lib.RED_MOMENT:
        return 4
    light = spec.mode in (_lib.MODE_R, _lib.MODE_RC) and spec.redop in (
        _lib.RED_SUM, _lib.RED_PROD, _lib.RED_MIN, _lib.RED_MAX, _lib.RED_ANY, _lib.RED_ALL, _lib.RED_NANMIN, _lib.RED_NANMAX)
    return 4 if light else 3


def _pow2_ceil(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def choose_geometry(program: Program, mode: int, shapes, vec: int) -> dict:
    """Pick (TX, TY, RPT, U) for a launch over blocks of canonical shapes ``shapes``.

    B200 sizing: 256-thread CTAs, 16-byte loads per thread, 8 (1 input) .. 2 loads in flight
    per thread, and ~256-512 KiB of input per CTA so that a 4 GiB array is >= 8k CTAs
    (>> 148 SMs x resident CTAs) while the two-stage partials stay <= 2 % of the traffic.
    """
    sizes = [d.itemsize for d in program.inputs] or [program.out_dtype.itemsize]
    if mode == _lib.MODE_EW:
        sizes = sizes + [program.out_dtype.itemsize]
    V = vec
    Cmax = max(s[2] for s in shapes)
    Rmax = max(s[1] for s in shapes)
    tx = min(256, _pow2_ceil(max(1, -(-Cmax // V))))
    ty = 256 // tx
    nin = max(1, len(program.inputs))
    U = 8 if nin == 1 else (4 if nin <= 3 else 2)
    row_bytes = sum(sizes) * (tx * V if mode != _lib.MODE_C else Cmax)
    target = 512 * 1024 if mode != _lib.MODE_EW else 256 * 1024
    if mode in (_lib.MODE_R, _lib.MODE_RC):
        # same-box sweep on B200 (c2 chain, 4 GiB): U=4 / 1 MiB tiles beat U=8 / 512 KiB by 6-16 %:
        # 4 loads in flight per thread x 4 (3 for the moment accumulator) CTAs/SM keeps HBM as busy with
        # fewer live registers (profiles/README.md)
        # With the packed f32x2 path the chain kernels are no longer issue-bound and deeper prefetch wins
        # again: U=8 at 3 CTAs/SM -> mean 6587, moment 6181 GB/s (second sweep, same box).
        U = 8 if (program.packable() and nin == 1 and V % 2 == 0) else min(U, 4)
        target = 1024 * 1024
        # Small launches (BASELINE config 2 dealt over 8 GPUs: 8 blocks = 512 MiB per rank): with 1 MiB tiles
        # that is 512 CTAs for 444 resident slots -- 1.15 waves, the kernel ran at 71 % of its large-launch
        # bandwidth (B200 x 8, round 2).  Sweep of that launch on one B200 (profiles/r2_small_launch_sweep.py,
        # mean / std kernel in us): rows per tile 16: 174 / 155, 32: 133 / 129, 64: 115 / 116, 128: 109 / 113,
        # 256: 120 / 127 -- ~2.3 waves is the sweet spot between wave quantisation and per-tile fixed cost
        # (descriptor fetch, two-stage partial, fold); never below 64 KiB.
        total = sum(s[0] * s[1] * s[2] for s in shapes) * max(sizes)
        target = int(min(target, max(64 * 1024, total // 1024)))
    if mode == _lib.MODE_C:
        # B200 sweep (c3, fp64 rows of 128 KiB): 2 loads in flight per thread at 4 CTAs/SM beat 8 at 3
        # (argmax 5.6-5.9 -> 6.8 TB/s, max 6.7 -> 7.2 TB/s): a CTA then walks its row nearly in order and
        # keeps fewer DRAM pages open; latency is covered by the extra resident warps instead
        U = 2
    rpt = max(1, min(Rmax, target // max(1, row_bytes)))
    step = ty * U if mode != _lib.MODE_C else ty
    rpt = -(-rpt // step) * step
    import os                                   # tuning overrides (sweeps only)
    if os.environ.get("B2_U"):
        U = int(os.environ["B2_U"])
    if os.environ.get("B2_RPT") and mode != _lib.MODE_C:
        rpt = int(os.environ["B2_RPT"])
    return dict(vec=V, tx=tx, ty=ty, rpt=int(rpt), unroll=U)
