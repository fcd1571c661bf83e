"""ctypes binding of ``libb200da.so`` (C ABI in ``include/b200da.h``).

There is no fallback of any kind: if the shared library is missing the import fails
with instructions to build it, and every compute entry point raises ``B2Error`` when no
CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200da.so")

B2_MAX_IN = 6
B2_MAX_ND = 4

# enums (include/b200da.h)
MODE_EW, MODE_R, MODE_C, MODE_RC, MODE_SR, MODE_SC = 0, 1, 2, 3, 4, 5
RED_NONE, RED_SUM, RED_MIN, RED_MAX, RED_ARGMIN, RED_ARGMAX, RED_MOMENT, RED_PROD, RED_ANY, RED_ALL = range(10)
RED_NANMIN, RED_NANMAX = 10, 11
POST_NONE, POST_MEAN, POST_VAR, POST_STD = range(4)

_DTYPE_CODES = {
    "bool": 0, "int8": 1, "uint8": 2, "int16": 3, "uint16": 4, "int32": 5, "uint32": 6,
    "int64": 7, "uint64": 8, "float32": 9, "float64": 10, "bfloat16": 11, "float16": 12,
}


def dtype_code(dt) -> int:
    name = dt if isinstance(dt, str) and dt in _DTYPE_CODES else np.dtype(dt).name
    try:
        return _DTYPE_CODES[name]
    except KeyError:
        raise B2Error(f"dtype {dt!r} is not supported by the B200 backend") from None


class B2Error(RuntimeError):
    """Raised for every non-zero status returned by libb200da."""


class Block(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p * B2_MAX_IN),
        ("in_sb", C.c_int64 * B2_MAX_IN),
        ("in_sr", C.c_int64 * B2_MAX_IN),
        ("in_sc", C.c_int64 * B2_MAX_IN),
        ("out0", C.c_void_p),
        ("out1", C.c_void_p),
        ("B", C.c_int64),
        ("R", C.c_int64),
        ("C", C.c_int64),
        ("tile_begin", C.c_int64),
        ("tiles_r", C.c_int64),
        ("tiles_c", C.c_int64),
        ("work", C.c_void_p),
        ("counter", C.c_void_p),
        ("arg_offset", C.c_int64),
        ("arg_ndim", C.c_int32),
        ("mirror", C.c_int32),
        ("arg_shape", C.c_int64 * B2_MAX_ND),
        ("arg_start", C.c_int64 * B2_MAX_ND),
        ("arg_total", C.c_int64 * B2_MAX_ND),
    ]


class Scalars(C.Structure):
    _fields_ = [("f", C.c_double * 8), ("i", C.c_int64 * 8)]


class Geom(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("redop", C.c_int32), ("vec", C.c_int32), ("tx", C.c_int32),
        ("ty", C.c_int32), ("rpt", C.c_int32), ("packed_bytes", C.c_int32), ("_pad", C.c_int32),
    ]


class Group(C.Structure):
    _fields_ = [
        ("parts", C.c_void_p), ("parts1", C.c_void_p), ("out0", C.c_void_p), ("out1", C.c_void_p),
        ("nelem", C.c_int64), ("elem_begin", C.c_int64), ("fanin", C.c_int32), ("post", C.c_int32),
        ("count", C.c_double), ("ddof", C.c_double),
    ]


class GemmProblem(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("npairs", C.c_int32), ("accumulate", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("C", C.c_void_p), ("ldc", C.c_int64),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("Kpair", C.c_void_p),
    ]


class WindowJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("B", C.c_int64), ("R", C.c_int64), ("C", C.c_int64),
                ("src_pitch", C.c_int64), ("tile_begin", C.c_int64), ("col_tiles", C.c_int64), ("row_tiles", C.c_int64)]


class Copy(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int64), ("row_bytes", C.c_int64),
        ("src_pitch", C.c_int64), ("dst_pitch", C.c_int64), ("tile_begin", C.c_int64),
        ("tile_rows", C.c_int32), ("tiles_c", C.c_int32), ("vec_bytes", C.c_int32), ("_pad", C.c_int32),
    ]


# every symbol declared in include/b200da.h (tests check the library exports all of them)
SYMBOLS = [
    "b2_abi_version", "b2_last_error", "b2_launch_count", "b2_device_sm_count",
    "b2_jit_compile", "b2_free", "b2_device_header", "b2_kernel_load", "b2_kernel_free",
    "b2_fused_plan", "b2_fused_launch", "b2_combine", "b2_gather_plan", "b2_gather_launch",
    "b2_fill", "b2_gemm_tn", "b2_memcpy2d", "b2_gemm_tn_pairs", "b2_split3_bf16", "b2_combine_groups", "b2_gemm_tn_batched", "b2_gather_launch_bulk",
    "b2_ipc_export", "b2_ipc_open", "b2_peer_barrier", "b2_topk_rows", "b2_peer_barrier_dev", "b2_peer_allgather", "b2_take", "b2_gemm_tn_simt", "b2_window_reduce_batched",
]


class IpcHandle(C.Structure):
    _fields_ = [("reserved", C.c_ubyte * 64), ("offset", C.c_int64), ("size", C.c_int64)]



def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dask_array_b200/csrc`. The B200 backend has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    lib.b2_abi_version.restype = i32
    lib.b2_last_error.restype = C.c_char_p
    lib.b2_launch_count.restype = i64
    lib.b2_device_sm_count.argtypes = [C.POINTER(i32)]
    lib.b2_jit_compile.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(sz)]
    lib.b2_free.argtypes = [vp]
    lib.b2_free.restype = None
    lib.b2_device_header.restype = C.c_char_p
    lib.b2_kernel_load.argtypes = [vp, sz, C.c_char_p, C.POINTER(Geom), C.POINTER(vp)]
    lib.b2_kernel_free.argtypes = [vp]
    lib.b2_fused_plan.argtypes = [vp, C.POINTER(Block), i32, vp, sz, C.POINTER(sz), C.POINTER(i64)]
    lib.b2_fused_launch.argtypes = [vp, vp, i32, i64, C.POINTER(Scalars), vp]
    lib.b2_combine.argtypes = [i32, i32, vp, vp, i32, i64, vp, vp, i32, i32, C.c_double, C.c_double, vp]
    lib.b2_combine_groups.argtypes = [i32, i32, i32, vp, i32, i64, vp]
    lib.b2_gather_plan.argtypes = [C.POINTER(Copy), i32, C.POINTER(i64)]
    lib.b2_gather_launch.argtypes = [vp, i32, i64, vp]
    lib.b2_gather_launch_bulk.argtypes = [vp, i32, i64, vp]
    lib.b2_fill.argtypes = [vp, i64, i32, vp, vp]
    lib.b2_memcpy2d.argtypes = [vp, i64, vp, i64, i64, i64, i32, vp]
    lib.b2_gemm_tn.argtypes = [i32, vp, i64, vp, i64, vp, i64, i64, i64, i64, i32, vp]
    lib.b2_gemm_tn_pairs.argtypes = [i32, vp, vp, i32, i64, i64, vp, i64, i64, i64, i64, i32, vp]
    lib.b2_split3_bf16.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.b2_gemm_tn_batched.argtypes = [i32, C.POINTER(GemmProblem), i32, vp, sz, C.POINTER(sz), vp]
    lib.b2_ipc_export.argtypes = [vp, C.POINTER(IpcHandle)]
    lib.b2_ipc_open.argtypes = [C.POINTER(IpcHandle), C.POINTER(vp)]
    lib.b2_peer_barrier.argtypes = [vp, i32, i32, C.c_uint64, vp]
    lib.b2_peer_barrier_dev.argtypes = [vp, vp, i32, i32, vp]
    lib.b2_gemm_tn_simt.argtypes = [i32, vp, i64, vp, i64, vp, i64, i64, i64, i64, i32, vp]
    lib.b2_window_reduce_batched.argtypes = [i32, i32, C.POINTER(WindowJob), i32, vp, i64, i32, i32, vp]
    lib.b2_take.argtypes = [i32, vp, vp, vp, i64, i64, i64, vp]
    lib.b2_peer_allgather.argtypes = [vp, vp, vp, vp, i64, i64, i32, i32, vp]
    lib.b2_topk_rows.argtypes = [i32, vp, i64, i64, i64, i32, i32, i32, vp, vp, i64, vp, i64, vp]
    for name in SYMBOLS:
        getattr(lib, name)  # AttributeError here = header and library out of sync
    return lib


lib = _load()
assert C.sizeof(Block) == 384, C.sizeof(Block)


def check(rc: int) -> None:
    if rc != 0:
        raise B2Error(f"libb200da error {rc}: {lib.b2_last_error().decode(errors='replace')}")


def jit_compile(source: str, name: str = "b2_fused.cu") -> bytes:
    """NVRTC-compile ``source`` for sm_100a; works without a GPU."""
    out = C.c_void_p()
    n = C.c_size_t()
    check(lib.b2_jit_compile(source.encode(), name.encode(), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out.value, n.value)
    finally:
        lib.b2_free(out)


def launch_count() -> int:
    return int(lib.b2_launch_count())
