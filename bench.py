#!/usr/bin/env python
"""bench.py -- fused blockwise+reduction throughput of the B200 backend (BASELINE.json).

Workload (``configs[1]``): fp32 ``x`` of shape (32768, 32768) in 4096^2 chunks (8 x 8 blocks,
4 GiB) per GPU; one step = ``(sin(x)*2 + x**2).mean(axis=0)`` AND ``(sin(x)*2 + x**2).std()``,
each a single pass over ``x`` (2 x 4.295 GB algorithmic bytes per GPU per step).  With N GPUs
the array is (32768, 32768*N): block columns are dealt block-cyclically, every rank holds 64
blocks (weak scaling); the per-block partials are all-gathered over NCCL and every rank folds
the tree.

    python bench.py --gpus N --steps K --warmup W            # this backend
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle)

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  Inputs (4 GiB per GPU) are far larger than the
126 MB L2, so no explicit L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BLOCK = 4096
GRID = 8                       # 8 x 8 blocks per GPU
ITEM = 4
BYTES_PER_PASS = GRID * GRID * BLOCK * BLOCK * ITEM      # 4.295 GB per reduction per GPU
METRIC = "fused blockwise+reduction GB/s"


def chain(x):
    import dask_array_b200 as da

    return da.sin(x) * 2 + x**2


# ----------------------------------------------------------------------------- host data
def host_blocks(rank: int, world: int, pinned: bool, nblock_cols: int = GRID):
    """This rank's blocks, generated on the host like ``random/_expr.py:29-32,97-126``:
    block (i, j) draws from ``SeedSequence(0).spawn(nblocks)[ravel(i, j)]``."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    ncols_total = nblock_cols * world
    cols = [j for j in range(ncols_total) if j % world == rank]
    n = GRID * len(cols)
    if pinned:
        store = torch.empty((n, BLOCK, BLOCK), dtype=torch.float32, pin_memory=True).numpy()
    else:
        store = np.empty((n, BLOCK, BLOCK), dtype=np.float32)
    seeds = np.random.SeedSequence(0).spawn(GRID * ncols_total)
    index = {}
    jobs = []
    for k, (i, j) in enumerate((i, j) for i in range(GRID) for j in cols):
        index[(i, j)] = k
        jobs.append((k, seeds[i * ncols_total + j]))

    def fill(job):
        k, seed = job
        np.random.Generator(np.random.PCG64(seed)).random(out=store[k], dtype=np.float32)

    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(fill, jobs))
    return store, index, ncols_total


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML while the benchmark runs."""

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples = []
        self.timed = False
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:      # NVML missing: report that instead of inventing numbers
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self.timed, sm, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        nv = self.nv
        under = [s for s in self.samples if s[0]] or self.samples
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        seen = set()
        for _, _, r in under:
            for n, bit in names.items():
                if r & bit:
                    seen.add(n)
        return {"sm_mhz": float(np.median([s[1] for s in under])) if under else None,
                "sm_max_mhz": float(self.max_sm), "reasons": sorted(seen), "samples": len(under)}


# ----------------------------------------------------------------------------- CPU legs
def cpu_sample_run(store, index, ncols_total, cols, workers):
    """The oracle (NumPy restatement of the reference's threaded compute) on a bounded sample:
    block columns ``cols`` (each 8 blocks of 4096^2 fp32), mean(axis=0) and std()."""
    from oracle import reference as ref

    blocks = {(i, jj): store[index[(i, j)]] for i in range(GRID) for jj, j in enumerate(cols)}
    x = ref.Blocked(blocks, ((BLOCK,) * GRID, (BLOCK,) * len(cols)))
    t0 = time.perf_counter()
    y = ref.elemwise(ref.fused_chain, x, workers=workers)      # 4 temporaries per block, as the reference
    m = ref.da_mean(y, axis=0, workers=workers)
    s = ref.da_std(y, workers=workers)
    dt = time.perf_counter() - t0
    nbytes = 2 * len(blocks) * BLOCK * BLOCK * ITEM
    return nbytes / dt / 1e9, dt, (m, s)


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    """``--impl reference``: the reference's own CPU implementation of the path.  The reference
    cannot be imported here (dask/toolz absent, no network: SURVEY.md 8c), so this is the
    oracle port, with all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    ncols = GRID                                         # the whole N=1 workload: all 64 blocks (4 GiB)
    store, index, ncols_total = host_blocks(0, 1, pinned=False, nblock_cols=ncols)
    cols = list(range(ncols))
    for _ in range(args.warmup):
        cpu_sample_run(store, index, ncols_total, cols, workers)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample_run(store, index, ncols_total, cols, workers)
    dt = time.perf_counter() - t0
    nbytes = args.steps * 2 * GRID * ncols * BLOCK * BLOCK * ITEM
    value = nbytes / dt / 1e9
    sample = f"{GRID * ncols} of 64 blocks (4096^2 fp32) per step, both reductions (the full single-GPU workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: (sin(x)*2+x**2).mean(axis=0) and .std(), fp32 (32768,32768) chunks 4096^2",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": workers, "kind": "port", "sample": sample,
                         "cpu": cpu_model(), "numpy": np.__version__},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- configs 3 and 4
def run_other(args):
    """Secondary configs (single GPU, resident data): c3 = fp64 argmax/argmin/min/max(axis=1) on
    (65536,16384) chunks (8192,16384); c4 = rechunk (16384,16384) (16384,256)->(256,16384) and x.T + x."""
    import torch
    import dask_array_b200 as da
    from dask_array_b200 import _lib

    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        if args.config != "c4":
            if rank == 0:
                print(json.dumps({"error": f"--config {args.config} is a single-GPU line; only c2 and c4 shard"}))
            return
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, nbytes, label, extra=None):
        for _ in range(max(args.warmup, 3)):
            step.run()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(args.steps):
            step.run()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out = {"metric": METRIC, "config": {"workload": label}, "value": nbytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
               "ms_per_step": ms, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "frac_of_measured_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak / world, "algorithmic_bytes_per_step": nbytes,
               "gpu_launches": _lib.launch_count() - n0, "data": "synthetic", "higher_is_better": True}
        out.update(extra or {})
        if world > 1:
            # (G-1)/G of the array crosses the partition; per-GPU egress over NVLink
            cross = nbytes / 2 * (world - 1) / world
            out.update({"scaling": "strong", "comm": os.environ.get("B2_COMM", "peer"),
                        "nvlink_bytes_per_step": cross, "nvlink_GBps_per_gpu_egress": cross / world / (ms * 1e-3) / 1e9})
        if rank == 0:
            print(json.dumps(out))

    if args.config == "c1":
        # README example: latency only (80 kB blocks are not a roofline config)
        x = da.ones((1000, 1000), chunks=(100, 100))
        for label, arr in (("(x + x.T)[:100, :100]", (x + x.T)[:100, :100]), ("(x + x.T).sum()", (x + x.T).sum())):
            t0 = time.perf_counter(); first = arr.compute(); cold = time.perf_counter() - t0
            step = da.compile(arr)
            for _ in range(5):
                step.run()
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step.run()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.steps
            t0 = time.perf_counter(); val = arr.compute(); full = time.perf_counter() - t0
            step.capture()                                  # the same tape as ONE CUDA graph
            for _ in range(5):
                step.run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step.run()
            torch.cuda.synchronize()
            gdt = (time.perf_counter() - t0) / args.steps
            print(json.dumps({"metric": "latency", "config": {"workload": "c1: README example " + label},
                              "replay_us": dt * 1e6, "graph_replay_us": gdt * 1e6,
                              "compute_call_ms": full * 1e3, "first_call_ms": cold * 1e3,
                              "launches_per_replay": (_lib.launch_count() - n0) / args.steps,
                              "result": float(np.asarray(val).ravel()[0]), "n_gpus": 1}))
        return
    if args.config == "cum":
        # cumulative scans.  Algorithmic bytes: read N + write N (the kernels move 3 N:
        # the per-segment totals are a separate read pass).
        n, cb = 32768, 4096
        base = np.random.default_rng(0).random((cb, cb), dtype=np.float32)
        x = da.from_host_blocks(lambda bid: base, (n, n), (cb, cb), np.float32, token="cum-f4").persist()
        for axis in (0, 1):
            step = da.compile(x.cumsum(axis=axis))
            timed(step, 2 * n * n * 4, f"cum: fp32 (32768,32768) chunks 4096^2 cumsum(axis={axis}) (2N bytes; 3N moved)",
                  {"dtype": "f32"})
            del step
        del x
        m = 1 << 28
        vb = np.random.default_rng(1).random(1 << 24)
        v = da.from_host_blocks(lambda bid: vb, (m,), (1 << 24,), np.float64, token="cum-f8").persist()
        step = da.compile(v.cumsum())
        timed(step, 2 * m * 8, "cum: fp64 vector 2^28 chunks 2^24 cumsum() (2N bytes; 3N moved)", {"dtype": "f64"})
        return
    if args.config == "c5":
        import ml_dtypes
        n, cb = 32768, 4096
        rng = np.random.default_rng(0)
        base32 = (rng.random((cb, cb), dtype=np.float32) - 0.5)
        tflops_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        for dt, label in ((ml_dtypes.bfloat16, "bf16"), (np.float32, "fp32 (bf16x3 split, 6 products)")):
            base = base32.astype(dt)
            x = da.from_host_blocks(lambda bid: base, (n, n), (cb, cb), dt, token=f"c5x{label}").persist()
            y = da.from_host_blocks(lambda bid: base, (n, n), (cb, cb), dt, token=f"c5y{label}").persist()
            step = da.compile((x @ y.T).sum())
            for _ in range(2):
                step.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(1, min(args.steps, 3))
            e0.record()
            for _ in range(reps):
                step.run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            flop = 2.0 * n * n * n
            mma_flop = flop * (6 if dt == np.float32 else 1)
            print(json.dumps({"metric": "blocked matmul TFLOP/s", "config": {"workload": f"c5: (x @ y.T).sum() 32768^2 {label} chunks 4096^2"},
                              "value": flop / (ms * 1e-3) / 1e12, "unit": "TFLOP/s (algorithmic 2N^3)", "ms_per_step": ms,
                              "tensor_pipe_TFLOPs": mma_flop / (ms * 1e-3) / 1e12,
                              "frac_of_measured_bf16_sustained": mma_flop / (ms * 1e-3) / 1e12 / tflops_peak,
                              "n_gpus": 1, "steps": reps, "data": "synthetic (one random 4096^2 block tiled)"}))
            del x, y, step
        return
    if args.config == "c3":
        R, Cc, RB = 65536, 16384, 8192
        rng = np.random.default_rng(0)

        def blk(bid):
            return np.random.default_rng(bid[0]).random((RB, Cc))
        x = da.from_host_blocks(blk, (R, Cc), (RB, Cc), np.float64, token="c3").persist()
        nbytes = R * Cc * 8
        for name in ("argmax", "argmin", "max", "min"):
            step = da.compile(getattr(x, name)(axis=1))
            timed(step, nbytes, f"c3: fp64 (65536,16384) chunks (8192,16384) {name}(axis=1)", {"dtype": "f64"})
        step = da.compile(x.argmax(axis=1), x.argmin(axis=1), x.max(axis=1), x.min(axis=1))
        timed(step, 4 * nbytes, "c3: all four reductions, four passes", {"dtype": "f64"})
    else:
        n = 16384

        def gen(dt, r0, nr, c0, nc):
            """value(i, j) = i * n + j (int32 bit patterns viewed as f4: distinct values)"""
            v = np.add.outer(np.arange(r0, r0 + nr, dtype=np.int64) * n, np.arange(c0, c0 + nc, dtype=np.int64))
            return v.astype(np.float64) if dt == np.float64 else v.astype(np.int32).view(np.float32)

        def source(dt, chunks):
            blk = lambda bid: gen(dt, bid[0] * chunks[0], chunks[0], bid[1] * chunks[1], chunks[1])
            return da.from_host_blocks(blk, (n, n), chunks, dt, token=f"c4-{np.dtype(dt).name}-{chunks}").persist()

        for dt in (np.float32, np.float64):
            item = np.dtype(dt).itemsize
            x = source(dt, (n, 256))
            y = x.rechunk((256, n))
            step = da.compile(y)
            timed(step, 2 * n * n * item, f"c4: rechunk (16384,16384) {np.dtype(dt).name} (16384,256)->(256,16384)",
                  {"dtype": np.dtype(dt).name})
            # bit-exact check of one new row panel this rank owns against the generator
            torch.cuda.synchronize()
            mine = [b for b in sorted(step.stores[0].blocks)][:1]
            for b in mine:
                got = step.stores[0].blocks[b].to_numpy()
                want = gen(dt, b[0] * 256, 256, 0, n)
                if not np.array_equal(got.view(np.uint8), want.view(np.uint8)):
                    raise SystemExit(f"bench c4: rechunked block {b} differs from the source on rank {rank}")
            del x, y, step
            sq = source(dt, (2048, 2048))
            step = da.compile(sq.T + sq)
            timed(step, 2 * n * n * item, f"c4: x.T + x (16384,16384) {np.dtype(dt).name} chunks 2048^2 "
                  "(2N bytes: mirror-pair kernel, every tile read once)" if world == 1 else
                  f"c4: x.T + x (16384,16384) {np.dtype(dt).name} chunks 2048^2 " + (
                      "(remote operand read in place over NVLink)" if os.environ.get("B2_COMM", "peer") != "nccl"
                      else "(remote blocks fetched with packed NCCL send/recv)"),
                  {"dtype": np.dtype(dt).name})
            del sq, step
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "cum"],
                    help="c2 = headline (default). c3 / c4 = the other BASELINE configs (extra lines for DESIGN.md)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config != "c2":
        return run_other(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dask_array_b200 as da
    from dask_array_b200 import _lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: generated on the host once, uploaded untimed for the resident measurement
    store, index, ncols_total = host_blocks(rank, world, pinned=True)
    shape = (GRID * BLOCK, ncols_total * BLOCK)
    xh = da.from_host_blocks(lambda bid: store[index[bid]], shape, (BLOCK, BLOCK), np.float32, token=f"c2-r{rank}")
    x = xh.persist()
    y = chain(x)
    step = da.compile(y.mean(axis=0), y.std())           # runs once: JIT + plan (untimed)
    fused = step.fused_launches()
    for k in fused:
        k.profile = True
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step.run()
    for k in fused:
        k.__dict__.pop("events", None)
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.timed = True
    e0.record()
    for _ in range(args.steps):
        step.run()
    e1.record()
    barrier()
    sampler.timed = False
    launches = _lib.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    nl = torch.tensor([launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    value = args.steps * 2 * BYTES_PER_PASS * world / (ms * 1e-3) / 1e9

    # ---- per-kernel roofline (CUDA events recorded around every launch in the timed region)
    kern = []
    for k in fused:
        ev = k.__dict__.get("events", [])
        if not ev or k.total_tiles < 1000:
            continue
        dur = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        kern.append({"redop": int(k.redop), "mode": int(k.mode), "ms": dur, "GBps": BYTES_PER_PASS / (dur * 1e-3) / 1e9,
                     "launches": len(ev)})
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("dominant_kernel_dram_bytes_per_launch")
    except OSError:
        pass
    dom = max(kern, key=lambda d: d["ms"]) if kern else None
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "achieved": dom["GBps"], "peak": peak, "unit": "GB/s", "frac": dom["GBps"] / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_PASS,
                    "kernel": "b2_fused<moment,RC>" if dom["redop"] == 6 else "b2_fused<sum,R>", "kernels": kern}

    # ---- end to end: host blocks in pinned memory -> H2D -> both reductions -> D2H results
    ye = chain(xh)
    e2e_step = da.compile(ye.mean(axis=0), ye.std())
    e2e_step.run(); e2e_step.results()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step.run()
        res = e2e_step.results()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": 2 * BYTES_PER_PASS * world / e2e_s / 1e9, "unit": "GB/s",
           "h2d_bytes_per_step": int(BYTES_PER_PASS * world), "d2h_bytes_per_step": int(res[0].nbytes + 4),
           "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps}
    sampler.stop()
    sampler.join(timeout=2)

    # ---- CPU baseline beside it (rank 0, N=1): oracle on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        cols = list(range(GRID))                                                    # all 64 blocks
        cpu_sample_run(store, index, ncols_total, cols[:1], workers)               # warm-up
        runs = [cpu_sample_run(store, index, ncols_total, cols, workers) for _ in range(5)]
        v, dt, (m, s) = max(runs, key=lambda r: r[0])
        cpu = {"value": v, "unit": "GB/s", "cores": workers, "kind": "port",
               "sample": f"all 64 blocks (the full N=1 workload), both reductions, best of 5 ({dt:.2f} s each, "
                         f"{sum(r[1] for r in runs):.1f} s of CPU work)",
               "cpu": cpu_model(), "numpy": np.__version__}
        # the GPU results agree with the CPU port (the oracle as checker)
        if not np.allclose(res[0], m, rtol=1e-5) or not np.allclose(res[1], s, rtol=1e-5):
            raise SystemExit("bench: GPU mean(axis=0) / std() disagree with the CPU oracle")

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c2: (sin(x)*2+x**2).mean(axis=0) and .std(), fp32 (32768, 32768*N) "
                                   "chunks 4096^2, 64 blocks (4 GiB) per GPU",
                       "global_shape": list(shape), "chunks": [BLOCK, BLOCK], "placement": "block-cyclic",
                       "l2": "inputs (4 GiB/GPU) >> 126 MB L2; no flush needed"},
            "pct_of_measured_hbm_peak": 100.0 * value / world / peak,
            "pct_of_nominal_8TBps": 100.0 * value / world / 8000.0,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(nl.item()), "clocks": sampler.summary(),
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
