"""Peer memory: blocks of OTHER ranks mapped into this process (one process per GPU).

The reference moves blocks between workers as pickled task results; the cross-partition steps of
the hot path are the rechunk all-to-all (``_rechunk.py:1171-1323``) and transposed / shifted block
reads of a fused expression (``manipulation/_transpose.py:66-75``).  On an NVSwitch box every GPU
can load from and store to every peer's HBM, so instead of pack -> NCCL send/recv -> unpack the
SAME gather and fused kernels run on pointers into the peer's memory (``include/b200da.h``:
``b2_ipc_export`` / ``b2_ipc_open``).  ``torch.distributed`` (NCCL) carries only the 80-byte
handles (once, at plan time) and the stream-ordered barriers around each kernel.

``B2_COMM=nccl`` selects the packed NCCL send/recv exchange instead (kept for A/B measurement).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


def enabled() -> bool:
    return os.environ.get("B2_COMM", "peer") != "nccl"


class PeerBuffer:
    """Stands where a ``torch.Tensor`` stands inside a ``DeviceChunk``: device memory owned by another
    rank, mapped here through CUDA IPC (the owner keeps the allocation alive)."""

    def __init__(self, ptr: int, device: torch.device, nbytes: int = 0, owner: int = -1):
        self._ptr, self.device, self.nbytes, self.owner = int(ptr), device, int(nbytes), owner

    def data_ptr(self) -> int:
        return self._ptr


def export_handle(ptr: int) -> bytes:
    h = _lib.IpcHandle()
    _lib.check(_lib.lib.b2_ipc_export(ptr, C.byref(h)))
    return bytes(h)


def open_handle(raw: bytes) -> int:
    h = _lib.IpcHandle.from_buffer_copy(raw)
    out = C.c_void_p()
    _lib.check(_lib.lib.b2_ipc_open(C.byref(h), C.byref(out)))
    return int(out.value)


HANDLE_BYTES = C.sizeof(_lib.IpcHandle)


def exchange_records(device, mine: list[bytes], counts: list[int], rec_bytes: int) -> list[list[bytes]]:
    """All-gather fixed-size records: rank r contributes ``counts[r]`` records (every rank knows all
    counts from the replicated expression metadata).  Returns the records per rank."""
    import torch.distributed as dist

    W = len(counts)
    cap = max(max(counts), 1) * rec_bytes
    send = torch.zeros(cap, dtype=torch.uint8)
    raw = b"".join(mine)
    assert len(raw) == len(mine) * rec_bytes
    if raw:
        send[: len(raw)] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    send = send.to(device)
    recv = torch.empty(cap * W, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    host = recv.cpu().numpy().tobytes()
    return [[host[r * cap + i * rec_bytes: r * cap + (i + 1) * rec_bytes] for i in range(counts[r])]
            for r in range(W)]


def exchange_pointers(device, my_ptrs: list[int], counts: list[int], rank: int, need=None) -> list[list[int]]:
    """Every rank exports ``my_ptrs``; returns, per rank, the same pointers as usable from THIS
    process (0 where ``need`` -- a set of (rank, index) -- says this rank never touches it)."""
    recs = exchange_records(device, [export_handle(p) for p in my_ptrs], counts, HANDLE_BYTES)
    out = []
    for r, rr in enumerate(recs):
        if r == rank:
            out.append(list(my_ptrs))
        else:
            out.append([open_handle(h) if (need is None or (r, i) in need) else 0 for i, h in enumerate(rr)])
    return out


class NcclBarrier:
    """Stream-ordered barrier as a 4-byte NCCL all-reduce (``B2_BARRIER=nccl``; ~20 us)."""

    def __init__(self, device):
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)

    def __call__(self):
        import torch.distributed as dist

        dist.all_reduce(self.flag)

    allgather = None          # the NCCL barrier has no peer windows: callers fall back to all_gather_into_tensor


class PeerBarrier:
    """Stream-ordered barrier through peer memory (``b2_peer_barrier_dev``): every rank stores an epoch
    into every peer's signal array over NVLink and waits for its own array to fill.  Everything the
    ranks enqueued before it has completed -- and is visible system-wide -- before anything enqueued
    after it runs.  The epoch is a device counter the kernel advances itself, so the launch is
    argument-stable and can be replayed from a CUDA graph.  One instance per process; all ranks call it
    in the same order."""

    def __init__(self, device, rank: int, world: int):
        self.rank, self.world, self.device = rank, world, device
        self.sig = torch.zeros(max(world, 2), dtype=torch.int64, device=device)
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        # the all-gather inside is ordered after the zero-fill on every rank: no signal can precede it
        ptrs = exchange_pointers(device, [self.sig.data_ptr()], [1] * world, rank)
        self.table = torch.tensor([p[0] for p in ptrs], dtype=torch.int64).to(device)

    def __call__(self):
        from ._device import current_stream_ptr

        _lib.check(_lib.lib.b2_peer_barrier_dev(self.table.data_ptr(), self.epoch.data_ptr(), self.rank, self.world,
                                                current_stream_ptr()))

    def allgather(self, windows_table: int, send_ptr: int, nbytes: int, slot_bytes: int):
        """``b2_peer_allgather`` on this barrier's signal table: barrier -> my record into slot ``rank`` of
        every rank's window -> barrier, one launch."""
        from ._device import current_stream_ptr

        _lib.check(_lib.lib.b2_peer_allgather(self.table.data_ptr(), self.epoch.data_ptr(), windows_table, send_ptr,
                                              nbytes, slot_bytes, self.rank, self.world, current_stream_ptr()))


_BARRIER = None


def StreamBarrier(device, rank: int = 0, world: int = 1):
    """The process-wide stream-ordered barrier (created collectively on first use)."""
    global _BARRIER
    if _BARRIER is None:
        _BARRIER = NcclBarrier(device) if os.environ.get("B2_BARRIER", "peer") == "nccl" else PeerBarrier(device, rank, world)
    return _BARRIER
