"""Round-2 host logic that needs no GPU: long contiguous copies folded into 4 KiB rows for the tiled gather, the
small-launch tile height, and the cross-GPU push plan of a sliding-window halo (every piece stored exactly once,
by the owner of its source block, into the slab of the owner of the new block)."""
import numpy as np
import pytest

import dask_array_b200 as da
from dask_array_b200 import _codegen as cg, _lib
from dask_array_b200._exchange import _fold_runs, owner_of, plan_rechunk_push
from dask_array_b200._prebuild import _chain


def test_fold_runs_keeps_bytes_and_alignment():
    descs = [(4096, 8192, 1, 64 << 20, 64 << 20, 64 << 20),          # one 64 MiB run -> rows of 4 KiB
             (4096, 8192, 16, 8192, 8192, 8192),                     # rows that ARE contiguous (pitch == row): folded
             (4096, 8192, 16, 8192, 16384, 8192),                    # pitched source: untouched
             (4100, 8192, 1, 1 << 20, 1 << 20, 1 << 20),             # unaligned source: untouched
             (4096, 8192, 1, (1 << 20) + 100, 0, 0)]                 # body + tail
    out = _fold_runs(descs)
    assert out[0] == (4096, 8192, (64 << 20) // 4096, 4096, 4096, 4096)
    assert out[1] == (4096, 8192, 32, 4096, 4096, 4096)
    assert out[2] == descs[2] and out[3] == descs[3]
    body, tail = out[4], out[5]
    assert body[2] * body[3] + tail[2] * tail[3] == (1 << 20) + 100 and tail[0] == 4096 + body[2] * 4096
    moved = lambda ds: sum(r * rb for _, _, r, rb, _, _ in ds)       # noqa: E731
    assert moved(out) == moved(descs)


def test_small_launch_tile_height():
    """choose_geometry: 1 MiB tiles for launches of >= ~2.3 waves, shorter tiles below (B200 sweep)."""
    rows = {n: cg.choose_geometry(_chain(), _lib.MODE_R, [(1, 4096, 4096)] * n, 4)["rpt"] for n in (64, 32, 16, 8, 1)}
    assert rows[64] == rows[32] == rows[16] == 256 and rows[8] == 128 and 8 <= rows[1] <= 32
    assert cg.choose_geometry(_chain(), _lib.MODE_RC, [(1, 2048, 8192)] * 8, 4)["rpt"] < 256


@pytest.mark.parametrize("W", [2, 3, 8])
def test_window_halo_push_plan_across_ranks(W):
    x = da.from_array(np.zeros((40, 12)), chunks=(6, 12))
    halo = da.sliding_window_view(x, 15, axis=0).sum(axis=-1).expr.optimize().operand("array")
    assert type(halo).__name__ == "WindowHalo"
    src = halo.operand("array")
    all_pieces = {(nbid, obid) for nbid in halo.block_ids() for obid, _, _ in halo.pieces(nbid)}
    pushed = set()
    layouts = None
    for me in range(W):
        layout, totals, pushes = plan_rechunk_push(halo, W, me)
        layouts = layouts or layout
        assert layout == layouts                                   # every rank derives the same slab layout
        for obid, sl, r, nbid, dsl in pushes:
            assert owner_of(src, obid, W) == me and owner_of(halo, nbid, W) == r
            assert (nbid, obid) not in pushed
            pushed.add((nbid, obid))
    assert pushed == all_pieces
