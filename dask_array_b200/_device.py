"""``DeviceChunk`` -- the GPU-resident chunk type of the B200 backend.

This is the object that takes the place of ``numpy.ndarray`` as a dask-array block
(the "chunk type" registered through ``dask_array/_chunk_types.py:31-54``).  It owns (or
views) device memory allocated by PyTorch -- PyTorch is used for allocation and streams
only -- and exposes it through ``__cuda_array_interface__`` and DLPack.  Views (basic
slices, transposes, broadcasts) are stride tricks, exactly like the NumPy views the
reference's chunk functions hand around (``np.transpose`` in
``manipulation/_transpose.py:17``, ``np.broadcast_to`` in ``creation/_utils.py:72``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

_TORCH_DTYPES = {
    "bool": torch.bool, "int8": torch.int8, "uint8": torch.uint8, "int16": torch.int16,
    "int32": torch.int32, "int64": torch.int64, "float32": torch.float32, "float64": torch.float64,
    "float16": torch.float16, "uint16": torch.uint16, "uint32": torch.uint32, "uint64": torch.uint64,
    "bfloat16": torch.bfloat16,
}


def current_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("dask_array_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def current_stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def alloc_bytes(nbytes: int, device=None, zero: bool = False) -> torch.Tensor:
    device = device or current_device()
    n = max(int(nbytes), 1)
    return (torch.zeros if zero else torch.empty)(n, dtype=torch.uint8, device=device)


def contiguous_strides(shape) -> tuple:
    st, acc = [], 1
    for n in reversed(shape):
        st.append(acc)
        acc *= max(int(n), 1)
    return tuple(reversed(st))


class DeviceChunk:
    """N-d strided view of device memory.  ``strides`` are in ELEMENTS."""

    __array_priority__ = 100.0   # wins ``max(key=__array_priority__)`` lookups (_core_utils.py:241)

    def __init__(self, buf: torch.Tensor, shape, dtype, strides=None, offset: int = 0):
        self.buf = buf                      # keeps the allocation alive
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.strides = tuple(int(s) for s in (strides if strides is not None else contiguous_strides(self.shape)))
        self.offset = int(offset)           # elements from the start of ``buf``

    # ---- construction
    @classmethod
    def empty(cls, shape, dtype, device=None) -> "DeviceChunk":
        dtype = np.dtype(dtype)
        shape = tuple(int(s) for s in shape)
        return cls(alloc_bytes(math.prod(shape) * dtype.itemsize, device), shape, dtype)

    @classmethod
    def from_numpy(cls, arr, device=None, out: "DeviceChunk | None" = None) -> "DeviceChunk":
        """Upload a host block.  A block that is a strided slice of a bigger C-ordered host
        array (from_array of a whole NumPy array) goes up with one pitched copy per 2-D plane,
        without a host-side gather; ``out`` re-uses an existing device block."""
        from . import _lib

        arr = np.asarray(arr)
        device = device or current_device()
        dst = out if out is not None else cls.empty(arr.shape, arr.dtype, device)
        if arr.size == 0:
            return dst
        item = arr.dtype.itemsize
        if arr.ndim >= 1 and arr.strides[-1] != item:
            arr = np.ascontiguousarray(arr)
        a2 = arr if arr.ndim >= 2 else arr.reshape(1, -1)
        # merge trailing dims that are contiguous, treat the next one as the pitched dim
        inner = a2.shape[-1] * item
        lead_shape = a2.shape[:-2]
        st = current_stream_ptr()
        rows, pitch = a2.shape[-2], a2.strides[-2]
        if pitch < inner and rows > 1:
            a2 = np.ascontiguousarray(a2)
            pitch = a2.strides[-2]
        base = a2.__array_interface__["data"][0]
        plane = rows * inner
        for n, idx in enumerate(np.ndindex(*lead_shape) if lead_shape else [()]):
            off = sum(i * s for i, s in zip(idx, a2.strides[:-2]))
            _lib.check(_lib.lib.b2_memcpy2d(dst.ptr + n * plane, inner, base + off, pitch, inner, rows, 0, st))
        dst._host_keepalive = a2
        return dst

    # ---- basic properties
    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return math.prod(self.shape)

    @property
    def itemsize(self) -> int:
        return self.dtype.itemsize

    @property
    def nbytes(self) -> int:
        return self.size * self.dtype.itemsize

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.offset * self.dtype.itemsize

    @property
    def device(self) -> torch.device:
        return self.buf.device

    @property
    def is_contiguous(self) -> bool:
        return self.size <= 1 or all(
            n == 1 or s == c for n, s, c in zip(self.shape, self.strides, contiguous_strides(self.shape))
        )

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of a 0-d chunk")
        return self.shape[0]

    def __repr__(self):
        return f"DeviceChunk(shape={self.shape}, dtype={self.dtype}, strides={self.strides}, device={self.device})"

    # ---- views
    def _view(self, shape, strides, offset=None) -> "DeviceChunk":
        return DeviceChunk(self.buf, shape, self.dtype, strides, self.offset if offset is None else offset)

    def transpose(self, *axes) -> "DeviceChunk":
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if not axes or axes == (None,):
            axes = tuple(reversed(range(self.ndim)))
        axes = tuple(a % self.ndim for a in axes)
        if sorted(axes) != list(range(self.ndim)):
            raise ValueError(f"axes {axes} do not match a {self.ndim}-d chunk")
        return self._view([self.shape[a] for a in axes], [self.strides[a] for a in axes])

    @property
    def T(self) -> "DeviceChunk":
        return self.transpose()

    def broadcast_to(self, shape) -> "DeviceChunk":
        shape = tuple(int(s) for s in shape)
        nd = len(shape)
        if nd < self.ndim:
            raise ValueError("cannot broadcast to fewer dimensions")
        sh = (1,) * (nd - self.ndim) + self.shape
        st = (0,) * (nd - self.ndim) + self.strides
        out = []
        for have, want, s in zip(sh, shape, st):
            if have == want:
                out.append(s)
            elif have == 1:
                out.append(0)
            else:
                raise ValueError(f"cannot broadcast {self.shape} to {shape}")
        return self._view(shape, out)

    def reshape(self, *shape) -> "DeviceChunk":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            known = -math.prod(shape)
            shape[shape.index(-1)] = self.size // known if known else 0
        if math.prod(shape) != self.size:
            raise ValueError(f"cannot reshape {self.shape} into {tuple(shape)}")
        if not self.is_contiguous:
            raise NotImplementedError("reshape of a non-contiguous DeviceChunk (copy it first)")
        return self._view(shape, contiguous_strides(shape))

    def __getitem__(self, index) -> "DeviceChunk":
        """Basic indexing only (slices with step 1.., ints, None, Ellipsis): always a view."""
        if not isinstance(index, tuple):
            index = (index,)
        if any(isinstance(i, (np.ndarray, DeviceChunk, list)) for i in index):
            from ._eager import fancy_getitem          # the integer-array gathers of _arg_combine

            return fancy_getitem(self, index)
        n_real = sum(1 for i in index if i is not None and i is not Ellipsis)
        if any(i is Ellipsis for i in index):
            k = next(p for p, i in enumerate(index) if i is Ellipsis)
            index = index[:k] + (slice(None),) * (self.ndim - n_real) + index[k + 1:]
        else:
            index = index + (slice(None),) * (self.ndim - n_real)
        shape, strides, offset, dim = [], [], self.offset, 0
        for ix in index:
            if ix is None:
                shape.append(1)
                strides.append(0)
                continue
            n, s = self.shape[dim], self.strides[dim]
            if isinstance(ix, slice):
                start, stop, step = ix.indices(n)
                cnt = len(range(start, stop, step))
                shape.append(cnt)
                strides.append(s * step)
                offset += start * s
            elif isinstance(ix, (int, np.integer)):
                i = int(ix)
                if i < 0:
                    i += n
                if not 0 <= i < n:
                    raise IndexError(f"index {ix} out of bounds for axis {dim} with size {n}")
                offset += i * s
            else:
                raise NotImplementedError(f"DeviceChunk supports basic indexing only, got {type(ix).__name__}")
            dim += 1
        return self._view(shape, strides, offset)

    # ---- interop
    @property
    def __cuda_array_interface__(self) -> dict:
        return {
            "shape": self.shape,
            "typestr": self.dtype.str,
            "data": (self.ptr, False),
            "strides": None if self.is_contiguous else tuple(s * self.itemsize for s in self.strides),
            "version": 3,
        }

    def as_torch(self) -> torch.Tensor:
        """Typed strided ``torch.Tensor`` view (zero copy)."""
        tdt = _TORCH_DTYPES.get(self.dtype.name)
        if tdt is None:
            raise NotImplementedError(f"no torch dtype for {self.dtype}")
        flat = self.buf.view(tdt)
        return torch.as_strided(flat, self.shape, self.strides, self.offset)

    def __dlpack__(self, stream=None):
        return self.as_torch().__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self.as_torch().__dlpack_device__()

    def to_numpy(self) -> np.ndarray:
        """Synchronous device -> host copy (strided views are gathered by torch's copy)."""
        if self.size == 0:
            return np.empty(self.shape, self.dtype)
        if any(s < 0 for s in self.strides):       # reversed views (x[::-1]): torch has no negative strides
            from ._eager import copy

            return copy(self).to_numpy()
        t = self.as_torch()
        if self.dtype.name in ("uint16", "uint32", "uint64", "bfloat16"):
            host = t.contiguous().view(torch.uint8).cpu().numpy().view(self.dtype)
            return host.reshape(self.shape)
        return t.cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.to_numpy()
        return a.astype(dtype) if dtype is not None else a
