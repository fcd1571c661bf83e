"""CPU checks (no GPU needed): the C-ABI library loads and exports every symbol the header
declares, generated kernels compile for sm_100a, compute entry points fail loudly without
a device, and the product never imports the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from dask_array_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b200da.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b2_status", "b2_dtype"}
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in b200da.h but not exported"
    assert set(_lib.SYMBOLS) == declared
    assert lib.b2_abi_version() == 1
    header_text = ctypes.c_char_p.in_dll  # noqa: B018  (keep ctypes import used)
    _lib.lib.b2_device_header.restype = ctypes.c_char_p
    assert b"b2_run" in _lib.lib.b2_device_header()


def test_struct_layouts_match_header():
    from dask_array_b200 import _lib
    assert ctypes.sizeof(_lib.Block) == 384
    assert ctypes.sizeof(_lib.Scalars) == 128
    assert ctypes.sizeof(_lib.Geom) == 32
    assert ctypes.sizeof(_lib.Copy) == 72


def test_config_kernels_compile_for_sm100a(tmp_path):
    from dask_array_b200 import _prebuild
    assert _prebuild.prebuild() >= 7
    # a cubin really is sm_100a
    from dask_array_b200 import _codegen as cg, _lib
    name, prog, layouts, mode, redop, shape, vec, acc = _prebuild.config_kernels()[0]
    spec = cg.KernelSpec(prog.key(), layouts, mode, redop, acc_dtype=acc.name, **cg.choose_geometry(prog, mode, [shape], vec))
    cubin = _lib.jit_compile(cg.render(prog, spec))
    path = tmp_path / "k.cubin"
    path.write_bytes(cubin)
    out = subprocess.run(["cuobjdump", "-lelf", str(path)], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.parametrize("op,args", [("floor_divide", ("int32", "int32")), ("remainder", ("float64", "float32")),
                                     ("where", ("bool", "int64", "float32")), ("left_shift", ("int16", "int16")),
                                     ("arctan2", ("float32", "float32")), ("logaddexp", ("float64", "float64")),
                                     ("isnan", ("float32",)), ("cos", ("float32",)), ("power", ("int64", 3)),
                                     ("maximum", ("uint8", "uint8")), ("less", ("int32", 2.5))])
def test_operator_vocabulary_compiles_and_types_like_numpy(op, args):
    from dask_array_b200 import _codegen as cg, _lib
    p = cg.Program()
    refs, dummies = [], []
    for a in args:
        if isinstance(a, str):
            refs.append(p.add_input(a)); dummies.append(np.ones((1,), a))
        else:
            refs.append(p.const(a)); dummies.append(a)
    out = p.op(op, *refs)
    p.set_output(out)
    fn = np.where if op == "where" else getattr(np, op)
    assert out.dtype == np.asarray(fn(*dummies)).dtype
    lay = tuple("V" for _ in p.inputs)
    spec = cg.KernelSpec(p.key(), lay, _lib.MODE_EW, _lib.RED_NONE, acc_dtype=p.out_dtype.name,
                         **cg.choose_geometry(p, _lib.MODE_EW, [(1, 64, 256)], 2))
    assert len(_lib.jit_compile(cg.render(p, spec))) > 1000


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    import dask_array_b200 as da
    x = da.from_array(np.ones((8, 8)), chunks=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        (x + 1).sum().compute()
    from dask_array_b200 import _lib
    n = ctypes.c_int()
    assert _lib.lib.b2_device_sm_count(ctypes.byref(n)) == -2          # B2_ERR_CUDA
    assert b"failed" in _lib.lib.b2_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dask_array_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "from oracle" not in src and "import oracle" not in src, f
    code = "import sys; sys.path.insert(0, %r); import dask_array_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" % ROOT
    subprocess.check_call([sys.executable, "-c", code])


def test_nvrtc_error_is_reported():
    from dask_array_b200 import _lib
    with pytest.raises(_lib.B2Error, match="NVRTC"):
        _lib.jit_compile("#include \"b2_device.cuh\"\nthis is not CUDA;")
