"""Appendix A of SURVEY.md -- the element-wise vocabulary of the fused-kernel code generator.  CPU part: EVERY
exported ufunc name types like NumPy (NEP 50 included) for float32 / float64 / int32 / mixed operands and its
kernel compiles for sm_100a (NVRTC needs no GPU).  The values are compared with NumPy on the GPU in
tests/test_zz_late_gpu.py."""
import numpy as np
import pytest

import dask_array_b200 as da
from dask_array_b200 import _codegen as cg, _lib

UNARY_FLOAT = ["negative", "positive", "exp", "exp2", "log", "log2", "log10", "log1p", "expm1", "sqrt", "square", "cbrt",
               "reciprocal", "sin", "cos", "tan", "arcsin", "arccos", "arctan", "sinh", "cosh", "tanh", "arcsinh",
               "arccosh", "arctanh", "deg2rad", "rad2deg", "degrees", "radians", "isfinite", "isinf", "isnan", "signbit",
               "floor", "ceil", "trunc", "rint", "fabs", "sign", "absolute", "abs", "logical_not"]
BINARY_FLOAT = ["add", "subtract", "multiply", "divide", "true_divide", "floor_divide", "power", "float_power", "remainder",
                "mod", "fmod", "logaddexp", "arctan2", "hypot", "greater", "greater_equal", "less", "less_equal",
                "not_equal", "equal", "logical_and", "logical_or", "logical_xor", "maximum", "minimum", "fmax", "fmin",
                "copysign", "nextafter"]
UNARY_INT = ["bitwise_not", "invert", "negative", "positive", "square", "absolute", "abs", "sign", "logical_not"]
BINARY_INT = ["bitwise_and", "bitwise_or", "bitwise_xor", "left_shift", "right_shift", "add", "subtract", "multiply",
              "floor_divide", "remainder", "mod", "fmod", "power", "maximum", "minimum", "true_divide", "greater", "equal"]


def test_the_lists_cover_every_exported_name():
    assert set(da.UFUNC_NAMES) <= set(UNARY_FLOAT + BINARY_FLOAT + UNARY_INT + BINARY_INT)
    for name in da.UFUNC_NAMES:
        assert callable(getattr(da, name))


def _check(op, dtypes):
    p = cg.Program()
    refs = [p.add_input(d) for d in dtypes]
    out = p.op(op, *refs)
    p.set_output(out)
    with np.errstate(all="ignore"):
        want = getattr(np, op)(*[np.ones((1,), d) for d in dtypes]).dtype
    assert out.dtype == want, (op, dtypes, out.dtype, want)
    spec = cg.KernelSpec(p.key(), tuple("V" for _ in p.inputs), _lib.MODE_EW, _lib.RED_NONE, acc_dtype=p.out_dtype.name,
                         **cg.choose_geometry(p, _lib.MODE_EW, [(1, 64, 256)], 2))
    assert len(_lib.jit_compile(cg.render(p, spec))) > 1000


@pytest.mark.parametrize("op", UNARY_FLOAT)
def test_unary_float(op):
    _check(op, ("float32",))
    _check(op, ("float64",))


@pytest.mark.parametrize("op", BINARY_FLOAT)
def test_binary_float(op):
    _check(op, ("float32", "float32"))
    _check(op, ("float64", "float32"))


@pytest.mark.parametrize("op", UNARY_INT)
def test_unary_int(op):
    _check(op, ("int32",))


@pytest.mark.parametrize("op", BINARY_INT)
def test_binary_int(op):
    _check(op, ("int32", "int32"))
    _check(op, ("int64", "int16"))


@pytest.mark.parametrize("op", ["add", "multiply", "true_divide", "power", "maximum", "less", "floor_divide", "remainder"])
def test_weak_python_scalars_follow_nep50(op):
    for dt, scalar in (("float32", 2), ("float32", 2.5), ("int32", 3), ("int32", 2.5), ("uint8", 3), ("int64", 2.0)):
        p = cg.Program()
        out = p.op(op, p.add_input(dt), p.const(scalar))
        with np.errstate(all="ignore"):
            want = getattr(np, op)(np.ones((1,), dt), scalar).dtype
        assert out.dtype == want, (op, dt, scalar, out.dtype, want)
