"""World-size-2 tests of the multi-GPU HOST logic on CPU (gloo): block-cyclic ownership and the
exchange schedules every rank derives independently must pair up (no deadlock, right bytes).
Blocks are NumPy arrays here; on the GPU the same schedules drive pack kernels + NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exchange(sends, recvs):
    ops = [dist.P2POp(dist.isend, t, p) for p, t in sends] + [dist.P2POp(dist.irecv, t, p) for p, t in recvs]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dask_array_b200 as da
        from dask_array_b200._blockwise import FusedPlan
        from dask_array_b200._executor import owner_of, plan_fused_exchange, plan_rechunk_exchange

        n = 64
        xh = np.arange(n * n, dtype=np.float64).reshape(n, n)
        if case == "records":
            # the handle exchange of the peer-memory path: fixed-size records, per-rank counts known to all
            from dask_array_b200 import _peer
            counts = [3, 1]
            mine = [bytes([rank * 16 + i]) * 80 for i in range(counts[rank])]
            recs = _peer.exchange_records(torch.device("cpu"), mine, counts, 80)
            ok = [len(r) for r in recs] == counts and all(
                recs[r][i] == bytes([r * 16 + i]) * 80 for r in range(world) for i in range(counts[r]))
            q.put((rank, bool(ok), 0))
        elif case == "rechunk":
            x = da.from_host_blocks(lambda bid: None, xh.shape, (n, 8), xh.dtype, token="gloo-x")   # opaque leaf: the rechunk stays
            expr = x.rechunk((8, n)).optimize().expr
            src = expr.operand("array")
            mine = {bid: xh[tuple(slice(s, s + k) for s, k in zip(src.block_start(bid), src.block_shape(bid)))].copy()
                    for bid in src.block_ids() if owner_of(src, bid, world) == rank}
            send_items, recv_items = plan_rechunk_exchange(expr, world, rank)
            sends, recvs, landing = [], [], {}
            for p in range(world):
                if send_items[p]:
                    flat = np.concatenate([mine[obid][sl].ravel() for obid, nbid, sl, shape, nb in send_items[p]])
                    sends.append((p, torch.from_numpy(flat)))
                if recv_items[p]:
                    total = sum(int(np.prod(shape)) for *_, shape, nb in recv_items[p])
                    buf = torch.empty(total, dtype=torch.float64)
                    recvs.append((p, buf))
                    off = 0
                    for obid, nbid, sl, shape, nb in recv_items[p]:
                        landing[(obid, nbid)] = (buf, off, shape)
                        off += int(np.prod(shape))
            _exchange(sends, recvs)
            ok = True
            for nbid in expr.block_ids():
                if owner_of(expr, nbid, world) != rank:
                    continue
                out = np.empty(expr.block_shape(nbid))
                for obid, sl, dsl in expr.pieces(nbid):
                    if obid in mine:
                        out[dsl] = mine[obid][sl]
                    else:
                        buf, off, shape = landing[(obid, nbid)]
                        out[dsl] = buf.numpy()[off:off + int(np.prod(shape))].reshape(shape)
                start = expr.block_start(nbid)
                ok &= np.array_equal(out, xh[start[0]:start[0] + out.shape[0], start[1]:start[1] + out.shape[1]])
            # 7/8... of the array crosses at G ranks: (G-1)/G of the bytes
            sent = sum(nb for p in send_items for *_, nb in send_items[p])
            q.put((rank, bool(ok), sent))
        else:   # x.T + x over a 4 x 4 block grid: block (i, j) needs x[j, i] from its owner
            x = da.from_array(xh, chunks=(16, 16))
            fused = (x.T + x).optimize().expr
            plan = FusedPlan(fused)
            send_items, recv_items = plan_fused_exchange(plan, [False] * len(plan.leaves), world, rank)
            dep = plan.leaves[0][0]
            blk = lambda bid: xh[bid[0] * 16:(bid[0] + 1) * 16, bid[1] * 16:(bid[1] + 1) * 16].copy()
            sends, recvs, got = [], [], {}
            for p in range(world):
                if send_items[p]:
                    for k, lbid, nb in send_items[p]:
                        assert owner_of(dep, lbid, world) == rank
                    sends.append((p, torch.from_numpy(np.concatenate([blk(lbid).ravel() for k, lbid, nb in send_items[p]]))))
                if recv_items[p]:
                    buf = torch.empty(sum(nb for *_, nb in recv_items[p]) // 8, dtype=torch.float64)
                    recvs.append((p, buf))
                    off = 0
                    for k, lbid, nb in recv_items[p]:
                        got[(k, lbid)] = (buf, off)
                        off += nb // 8
            _exchange(sends, recvs)
            ok = True
            for bid in fused.block_ids():
                if owner_of(fused, bid, world) != rank:
                    continue
                vals = []
                for k in range(len(plan.leaves)):
                    lbid = plan.leaf_block_id(k, bid)
                    if owner_of(dep, lbid, world) == rank:
                        vals.append(blk(lbid))
                    else:
                        buf, off = got[(k, lbid)]
                        vals.append(buf.numpy()[off:off + 256].reshape(16, 16))
                # which leaf is the transposed one is recorded in its dimension map
                tot = sum(v.T if plan.leaves[k][1] == (1, 0) else v for k, v in enumerate(vals))
                want = (xh.T + xh)[bid[0] * 16:(bid[0] + 1) * 16, bid[1] * 16:(bid[1] + 1) * 16]
                ok &= np.array_equal(tot, want)
            q.put((rank, bool(ok), sum(nb for p in send_items for *_, nb in send_items[p])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["rechunk", "fused_transpose", "records"])
def test_exchange_schedules_pair_up(case):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + ["rechunk", "fused_transpose", "records"].index(case)
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    sent = sum(s for _, _, s in results)
    if case == "rechunk":
        assert sent == 64 * 64 * 8 // 2          # (G-1)/G of the array crosses the partition


def test_owner_is_block_cyclic():
    sys.path.insert(0, ROOT)
    import dask_array_b200 as da
    from dask_array_b200._executor import owner_of
    x = da.from_array(np.zeros((64, 64)), chunks=(8, 8)).expr
    assert [owner_of(x, (0, j), 8) for j in range(8)] == list(range(8))
    assert owner_of(x, (3, 5), 8) == 5 and owner_of(x, (3, 5), 1) == 0
    assert owner_of(x, (1, 0), 3) == 8 % 3


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("old,new", [((64, 8), (8, 64)), ((16, 16), (64, 4)), ((20, 7), (9, 64))])
def test_rechunk_push_plan_covers_every_new_block_once(world, old, new):
    """Peer-store rechunk: the pushes of all ranks, applied to per-rank slabs laid out as every rank
    derives them, must write every element of every new block exactly once with the right value."""
    sys.path.insert(0, ROOT)
    import dask_array_b200 as da
    from dask_array_b200._executor import owner_of, plan_rechunk_push
    n = 64
    xh = np.arange(n * n, dtype=np.float64).reshape(n, n)
    expr = da.from_host_blocks(lambda bid: None, xh.shape, old, xh.dtype, token=f"push-{old}").rechunk(new).optimize().expr
    src = expr.operand("array")
    plans = [plan_rechunk_push(expr, world, me) for me in range(world)]
    layout, totals, _ = plans[0]
    assert all(p[0] == layout and p[1] == totals for p in plans)            # same layout on every rank
    slabs = [np.full(t // 8, np.nan) for t in totals]
    hits = [np.zeros(t // 8, dtype=np.int64) for t in totals]
    moved = 0
    for me, (_, _, pushes) in enumerate(plans):
        for obid, sl, r, nbid, dsl in pushes:
            assert owner_of(src, obid, world) == me and owner_of(expr, nbid, world) == r
            start = src.block_start(obid)
            piece = xh[tuple(slice(s0 + s.start, s0 + s.stop) for s0, s in zip(start, sl))]
            shape = expr.block_shape(nbid)
            off = layout[r][nbid] // 8
            view = slabs[r][off: off + shape[0] * shape[1]].reshape(shape)
            view[dsl] = piece
            hits[r][off: off + shape[0] * shape[1]].reshape(shape)[dsl] += 1
            moved += piece.size * (me != r)
    for nbid in expr.block_ids():
        r = owner_of(expr, nbid, world)
        shape, start = expr.block_shape(nbid), expr.block_start(nbid)
        off = layout[r][nbid] // 8
        got = slabs[r][off: off + shape[0] * shape[1]].reshape(shape)
        assert np.array_equal(got, xh[start[0]:start[0] + shape[0], start[1]:start[1] + shape[1]])
        assert (hits[r][off: off + shape[0] * shape[1]] == 1).all()
    if old == (64, 8) and new == (8, 64) and 8 % world == 0:
        assert moved == n * n * (world - 1) // world                          # (G-1)/G crosses NVLink


@pytest.mark.parametrize("world", [2, 4])
def test_fused_peer_read_plan(world):
    sys.path.insert(0, ROOT)
    import dask_array_b200 as da
    from dask_array_b200._blockwise import FusedPlan
    from dask_array_b200._executor import owner_of, plan_fused_peer_reads
    x = da.from_array(np.zeros((64, 64)), chunks=(16, 16))
    fused = (x.T + x).optimize().expr
    plan = FusedPlan(fused)
    exports, readers = plan_fused_peer_reads(plan, [False] * len(plan.leaves), world)
    dep = plan.leaves[0][0]
    for o, ex in enumerate(exports):
        assert len(set(ex)) == len(ex)
        for i, (k, lbid) in enumerate(ex):
            assert owner_of(dep, lbid, world) == o and o not in readers[(o, i)]
    # every remote read of every output block is covered by exactly one export of the owner
    for bid in fused.block_ids():
        r = owner_of(fused, bid, world)
        for k in range(len(plan.leaves)):
            lbid = plan.leaf_block_id(k, bid)
            o = owner_of(dep, lbid, world)
            if o != r:
                i = [j for j, (kk, lb) in enumerate(exports[o]) if lb == lbid]
                assert len(i) == 1 and r in readers[(o, i[0])]
