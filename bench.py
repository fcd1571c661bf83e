#!/usr/bin/env python
"""bench.py -- fused blockwise+reduction throughput of the B200 backend (BASELINE.json).

Headline workload (``configs[1]``, "c2"): fp32 ``x`` in 4096^2 chunks; one step =
``(sin(x)*2 + x**2).mean(axis=0)`` AND ``(sin(x)*2 + x**2).std()``, each a single pass over ``x``
(2 x 4 bytes per element of algorithmic traffic per step).

* ``value`` (``"scaling": "weak"``): every GPU holds 8 x 8 blocks (4 GiB); with N GPUs the array is
  (32768, 32768*N), block columns dealt block-cyclically.  ``mean(axis=0)`` is device-local, the
  ``std()`` tree needs ONE exchange of (n, mean, M2) triples: a single peer-memory all-gather launch
  (no NCCL in the step).  The step is replayed from one CUDA graph per rank.
* ``strong``: the SAME step on BASELINE's fixed (32768, 32768) array dealt over the N GPUs.
* ``first_call_ms`` / ``compute_call_ms``: what the replayed tape leaves out -- the first call (optimise, kernel
  cubins from the JIT cache, planning, launch) and a plain ``da.compute()`` of the same graph (host planning +
  launches + the gather of the results), both untimed for ``value``.
* ``parity``: every rank checks its share of the results against an fp64 ground truth of its own
  blocks; rank 0 also checks against the CPU oracle (the restatement of the reference).
* ``configs``: the other BASELINE configs, measured in the same run with the same clock sampler:
  c3 (fp64 arg/min/max), c4 (rechunk + ``x.T + x``; NVLink all-to-all at N > 1), c5 (blocked
  matmul, tcgen05), the cumulative scans and the sliding-window reductions.

    python bench.py --gpus N --steps K --warmup W            # this backend
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)
    python bench.py --config c3|c4|c5|cum|c1                 # one secondary config alone (profiling)

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  Inputs are far larger than the 126 MB L2 (>= 512 MiB per GPU in every
timed configuration), so no explicit L2 flush is needed between iterations.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BLOCK = 4096
GRID = 8                       # 8 x 8 blocks per GPU (weak) / in total (strong)
ITEM = 4
BYTES_PER_PASS = GRID * GRID * BLOCK * BLOCK * ITEM      # 4.295 GB per reduction over 64 blocks
METRIC = "fused blockwise+reduction GB/s"
RTOL_F32 = 1e-5                # north_star tolerance for fp32 reductions


def chain(x):
    import dask_array_b200 as da

    return da.sin(x) * 2 + x**2


def c2_config(world: int) -> dict:
    """The ``config`` object of the JSON line -- the SAME for this backend and for ``--impl reference``."""
    return {"workload": "c2: (sin(x)*2+x**2).mean(axis=0) and .std(), fp32, chunks 4096^2, 64 blocks (4 GiB) per GPU",
            "global_shape": [GRID * BLOCK, GRID * BLOCK * world], "chunks": [BLOCK, BLOCK],
            "placement": "block-cyclic (owner = ravel(block id) mod N)",
            "l2": "inputs (>= 512 MiB per GPU) >> 126 MB L2; no flush needed"}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def nthreads(ctx=None):
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    if ctx is not None and ctx.world > 1:
        n = max(1, n // ctx.local_world)
    return n


# ----------------------------------------------------------------------------- host data
def host_blocks(rank: int, world: int, pinned: bool, ncols_total: int, threads: int):
    """This rank's blocks of the c2 input, generated on the host like ``random/_expr.py:29-32,97-126``:
    block (i, j) draws from ``SeedSequence(0).spawn(nblocks)[ravel(i, j)]``."""
    import torch

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

    with ThreadPoolExecutor(max_workers=max(1, min(32, threads))) as ex:
        list(ex.map(fill, jobs))
    return store, index, cols


# ----------------------------------------------------------------------------- NUMA placement
def bind_near_gpu(local: int) -> dict:
    """Bind this rank (CPU affinity + preferred memory node) to the NUMA node of its GPU, so that the
    pinned staging buffers of the end-to-end leg sit next to the GPU's PCIe root.  Best effort: the
    outcome is reported, never fatal."""
    info = {}
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info["cpus"] = len(use)
        else:
            info["cpus"] = f"node cpus not in the allowed set ({len(allowed)} allowed)"
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))      # set_mempolicy(MPOL_PREFERRED)
        info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
    except Exception as e:          # noqa: BLE001  (sysfs layout / permissions differ per box)
        info["error"] = repr(e)[:120]
    return info


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled through NVML while the benchmark runs; samples carry the
    label of the timed section they fall into."""

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples = []
        self.section = None
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
                self.samples.append((self.section, sm, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()

    def summary(self, section):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        nv = self.nv
        under = [s for s in self.samples if s[0] == section]
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


class Ctx:
    """Per-process facts + the timing helpers shared by every config."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world)))
        torch.cuda.set_device(self.local)
        self.numa = bind_near_gpu(self.local) if os.environ.get("B2_NUMA", "1") == "1" else {"disabled": True}
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.torch, self.dist = torch, dist
        self.peaks = load_peaks()
        self.traffic = load_traffic()
        self.hbm_peak = float(self.peaks.get("hbm_gbs", 6650.0))
        self.peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in self.peaks else "fallback 6650 GB/s"
        self.sampler = ClockSampler(self.local)
        self.sampler.start()
        self.threads = nthreads(self)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v: int) -> int:
        t = self.torch.tensor([v], dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def all_ok(self, ok: bool, what: str):
        """Parity verdicts are collective: every rank must pass, or every rank stops."""
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        if int(t.item()) != 1:
            raise SystemExit(f"bench: PARITY FAILURE in {what}" + ("" if ok else f" (rank {self.rank})"))

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def timed(self, step, steps: int, warmup: int, section: str, graph: bool = True):
        """W warm-up replays, then K timed replays between CUDA events (max over ranks).  With
        ``graph`` the launch tape is first captured into one CUDA graph per rank."""
        from dask_array_b200 import _lib

        torch = self.torch
        n0 = _lib.launch_count()
        step.run()                                         # one plain replay: the launches of a step, counted
        per_step = _lib.launch_count() - n0
        if graph and os.environ.get("B2_BENCH_GRAPH", "1") == "1":
            for k in step.fused_launches():
                k.profile = False
            step.capture()
        for _ in range(max(warmup, 3)):
            step.run()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sampler.section = section
        e0.record()
        for _ in range(steps):
            step.run()
        e1.record()
        self.barrier()
        self.sampler.section = None
        ms = self.max_over_ranks(e0.elapsed_time(e1) / steps)
        return ms, per_step * steps                        # (a graph replay does not pass through the counter)

    def kernel_times(self, step, steps: int):
        """Second pass of K steps with a CUDA-event pair around every fused launch (not capturable in a
        graph): [(FusedLaunch, mean ms)] for the roofline of the dominant kernel."""
        g, step._graph = step._graph, None
        fused = step.fused_launches()
        for k in fused:
            k.profile = True
            k.__dict__.pop("events", None)
        for _ in range(steps):
            step.run()
        self.torch.cuda.synchronize()
        out = []
        for k in fused:
            ev = k.__dict__.pop("events", [])
            k.profile = False
            if ev:
                out.append((k, float(np.mean([a.elapsed_time(b) for a, b in ev])), len(ev)))
        step._graph = g
        return out

    def roofline(self, achieved, kernel, traffic_key, nbytes, **extra):
        out = {"bound": "hbm", "achieved": achieved, "peak": self.hbm_peak, "unit": "GB/s", "frac": achieved / self.hbm_peak,
               "traffic": self.traffic.get(traffic_key), "peak_source": self.peak_src,
               "algorithmic_bytes_per_launch": int(nbytes), "kernel": kernel}
        out.update(extra)
        return out


# ----------------------------------------------------------------------------- CPU legs
def cpu_sample_run(store, index, cols, workers):
    """The oracle (NumPy restatement of the reference's threaded compute) on block columns ``cols``
    (each 8 blocks of 4096^2 fp32): mean(axis=0) and std()."""
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
    oracle port, with all host threads, on a bounded sample of the same workload: 64 blocks per
    step (the whole workload at N=1, 1/N of it at N GPUs -- throughput does not depend on it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    store, index, cols = host_blocks(0, 1, pinned=False, ncols_total=GRID, threads=workers)
    for _ in range(args.warmup):
        cpu_sample_run(store, index, cols, workers)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample_run(store, index, cols, workers)
    dt = time.perf_counter() - t0
    nbytes = args.steps * 2 * GRID * GRID * BLOCK * BLOCK * ITEM
    value = nbytes / dt / 1e9
    sample = "64 blocks (4096^2 fp32, 4 GiB) per step, both reductions (= the N=1 workload; 1/N of the N-GPU workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": c2_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": workers, "kind": "port", "sample": sample,
                         "cpu": cpu_model(), "numpy": np.__version__},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- c2
def c2_truth(store, index, threads):
    """fp64 ground truth of this rank's blocks: per block the column sums of y = sin(x)*2 + x**2 (the
    element-wise chain in fp32 exactly as the reference evaluates it, every accumulation in fp64) and
    the block's (n, mean, M2)."""
    def one(item):
        bid, k = item
        b = store[k]
        y = np.sin(b) * 2 + b**2
        cs = y.sum(axis=0, dtype=np.float64)
        n = y.size
        mu = cs.sum() / n
        d = y.astype(np.float64)
        d -= mu
        np.multiply(d, d, out=d)
        return bid, cs, (n, mu, float(d.sum()))

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        return list(ex.map(one, index.items()))


def merge_moments(triples):
    """Chan merge of (n, mean, M2) triples in fp64 (``moment_combine``, ``_common.py:415-453``)."""
    n = sum(t[0] for t in triples)
    mu = sum(t[0] * t[1] for t in triples) / n
    m2 = sum(t[2] + t[0] * (t[1] - mu) ** 2 for t in triples)
    return n, mu, m2


def run_c2(ctx: Ctx, strong: bool, with_e2e: bool, with_cpu: bool):
    import dask_array_b200 as da

    args, W, rank = ctx.args, ctx.world, ctx.rank
    tag = "c2-strong" if strong else "c2"
    ncols_total = GRID if strong else GRID * W
    store, index, cols = host_blocks(rank, W, pinned=with_e2e, ncols_total=ncols_total, threads=ctx.threads)
    shape = (GRID * BLOCK, ncols_total * BLOCK)
    xh = da.from_host_blocks(lambda bid: store[index[bid]], shape, (BLOCK, BLOCK), np.float32, token=f"{tag}-r{rank}-w{W}")
    x = xh.persist()
    y = chain(x)
    ctx.torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = da.compile(y.mean(axis=0), y.std())           # runs once: optimise + JIT (or cubin cache) + plan + launch
    ctx.torch.cuda.synchronize()
    first_call_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    da.compute(y.mean(axis=0), y.std())                  # a plain compute() call of the same graph: nothing replayed
    compute_call_ms = (time.perf_counter() - t0) * 1e3
    total_bytes = 2 * shape[0] * shape[1] * ITEM          # both reductions, all ranks
    local_pass_bytes = len(index) * BLOCK * BLOCK * ITEM  # one reduction, this rank
    ms, launches = ctx.timed(step, args.steps, args.warmup, tag)
    launches = ctx.sum_over_ranks(launches)
    value = total_bytes / (ms * 1e-3) / 1e9
    out = {"value": value, "unit": "GB/s", "ms_per_step": ms, "gpu_launches": launches,
           "cuda_graph": step._graph is not None, "clocks": ctx.sampler.summary(tag),
           "first_call_ms": first_call_ms, "compute_call_ms": compute_call_ms,
           "pct_of_measured_hbm_peak": 100.0 * value / W / ctx.hbm_peak, "pct_of_nominal_8TBps": 100.0 * value / W / 8000.0}

    # ---- per-kernel roofline: second pass of K steps with CUDA events around every fused launch
    kern = []
    for k, dur, n in ctx.kernel_times(step, args.steps):
        if k.total_tiles < 100:
            continue
        kern.append({"kernel": "b2_fused<chain,moment,RC>" if int(k.redop) == 6 else "b2_fused<chain,sum,R>", "ms": dur,
                     "GBps": local_pass_bytes / (dur * 1e-3) / 1e9, "launches": n})
    if kern:
        dom = max(kern, key=lambda d: d["ms"])
        out["roofline"] = ctx.roofline(dom["GBps"], dom["kernel"], "c2_moment" if "moment" in dom["kernel"] else "c2_sum",
                                       local_pass_bytes, kernels=kern,
                                       timed_pass="K further steps with a CUDA-event pair around each launch (the graph "
                                                  "replay of `value` cannot hold per-launch events)")

    # ---- parity: every rank against an fp64 ground truth of its own blocks
    res = step.results()
    truth = c2_truth(store, index, ctx.threads)
    colsum = {}
    for (i, j), cs, _ in truth:
        colsum[j] = colsum.get(j, 0) + cs
    err_mean = 0.0
    for j, cs in colsum.items():
        want = cs / shape[0]
        got = res[0][j * BLOCK:(j + 1) * BLOCK]
        err_mean = max(err_mean, float(np.max(np.abs(got - want) / np.abs(want))))
    trip = [t for part in ctx.gather_objects([t for _, _, t in truth]) for t in part]
    n, mu, m2 = merge_moments(trip)
    std_truth = float(np.sqrt(m2 / n))
    err_std = abs(float(res[1]) - std_truth) / std_truth
    ctx.all_ok(err_mean <= RTOL_F32 and err_std <= RTOL_F32, f"{tag}: fp64 ground truth (mean err {err_mean:.2e}, std err {err_std:.2e})")
    parity = {"checked": True, "ranks": W, "rtol": RTOL_F32, "vs": "fp64 ground truth of every rank's own blocks",
              "mean_max_rel_err": ctx.max_over_ranks(err_mean), "std_rel_err": err_std}
    # rank 0: one block column through the CPU oracle (the reference's fp32 tree order)
    if rank == 0:
        _, _, (m, _) = cpu_sample_run(store, index, cols[:1], ctx.threads)
        j = cols[0]
        ok = np.allclose(res[0][j * BLOCK:(j + 1) * BLOCK], m, rtol=RTOL_F32, atol=0)
        parity["oracle_block_column"] = bool(ok)
    else:
        ok = True
    ctx.all_ok(ok, f"{tag}: oracle mean(axis=0) of block column {cols[0]}")
    out["parity"] = parity

    # ---- end to end: host blocks in pinned memory -> H2D -> both reductions -> D2H results
    if with_e2e:
        ye = chain(xh)
        e2e_step = da.compile(ye.mean(axis=0), ye.std())
        e2e_step.run(); e2e_step.results()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step.run()
            r2 = e2e_step.results()
        ctx.barrier()
        e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        out["e2e"] = {"value": total_bytes / e2e_s / 1e9, "unit": "GB/s",
                      "h2d_bytes_per_step": int(total_bytes // 2), "d2h_bytes_per_step": int(r2[0].nbytes + 4),
                      "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
                      "h2d_GBps_per_gpu": local_pass_bytes / e2e_s / 1e9, "numa": ctx.numa}
        del e2e_step, ye

    # ---- CPU baseline beside it (rank 0, N=1): the oracle on the whole N=1 workload
    if with_cpu and rank == 0 and W == 1:
        workers = os.cpu_count() or 1
        cpu_sample_run(store, index, cols[:1], workers)               # warm-up
        runs = [cpu_sample_run(store, index, cols, workers) for _ in range(3)]
        v, dt, (m, s) = max(runs, key=lambda r: r[0])
        out["cpu_baseline"] = {"value": v, "unit": "GB/s", "cores": workers, "kind": "port",
                               "sample": f"all 64 blocks (the full N=1 workload), both reductions, best of 3 ({dt:.2f} s each, "
                                         f"{sum(r[1] for r in runs):.1f} s of CPU work)",
                               "cpu": cpu_model(), "numpy": np.__version__}
        if not np.allclose(res[0], m, rtol=RTOL_F32, atol=0) or not np.allclose(res[1], s, rtol=RTOL_F32, atol=0):
            raise SystemExit("bench: GPU mean(axis=0) / std() disagree with the CPU oracle")
        out["parity"]["oracle_full"] = True
    del step, x, xh, y
    return out, shape


# ----------------------------------------------------------------------------- c3
def run_c3(ctx: Ctx):
    """fp64 argmax/argmin/max/min(axis=1) on (65536, 16384), chunks (8192, 16384): 8 blocks of 1 GiB dealt
    over the ranks; the reduced axis lies inside a block, so no exchange at any N."""
    import dask_array_b200 as da

    args, W, rank = ctx.args, ctx.world, ctx.rank
    R, Cc, RB = 65536, 16384, 8192
    nb = R // RB
    seeds = np.random.SeedSequence(3).spawn(nb)
    mine = [i for i in range(nb) if i % W == rank]
    host = {}

    def gen(i):
        host[i] = np.random.Generator(np.random.PCG64(seeds[i])).random((RB, Cc))

    with ThreadPoolExecutor(max_workers=max(1, min(len(mine), ctx.threads))) as ex:
        list(ex.map(gen, mine))
    x = da.from_host_blocks(lambda bid: host[bid[0]], (R, Cc), (RB, Cc), np.float64, token=f"c3-w{W}").persist()
    nbytes = R * Cc * 8
    local_bytes = len(mine) * RB * Cc * 8
    lines, results = {}, {}
    for name in ("argmax", "argmin", "max", "min"):
        step = da.compile(getattr(x, name)(axis=1))
        ms, launches = ctx.timed(step, args.steps, args.warmup, "c3")
        kt = [(k, dur) for k, dur, _ in ctx.kernel_times(step, args.steps) if k.total_tiles >= 100]
        dur = max((d for _, d in kt), default=ms)
        lines[name] = {"ms_per_step": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "kernel_ms": dur,
                       "kernel_GBps": local_bytes / (dur * 1e-3) / 1e9, "gpu_launches": ctx.sum_over_ranks(launches)}
        results[name] = step.results()[0]
        del step
    # parity: bit-exact against NumPy on every block this rank holds
    ok = True

    def check(i):
        good = True
        for name in ("argmax", "argmin", "max", "min"):
            want = getattr(np, name)(host[i], axis=1)
            good &= np.array_equal(results[name][i * RB:(i + 1) * RB], want)
        return good

    with ThreadPoolExecutor(max_workers=max(1, min(len(mine), ctx.threads))) as ex:
        ok = all(ex.map(check, mine))
    ctx.all_ok(ok, "c3: argmax/argmin/max/min(axis=1) bit-exact vs NumPy")
    worst = max(lines, key=lambda k: lines[k]["kernel_ms"])
    total_ms = sum(v["ms_per_step"] for v in lines.values())
    out = {"workload": "c3: fp64 (65536,16384) chunks (8192,16384): argmax, argmin, max, min (axis=1), four passes",
           "value": 4 * nbytes / (total_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": total_ms, "n_gpus": W, "dtype": "f64",
           "scaling": "strong", "per_reduction": lines,
           "roofline": ctx.roofline(lines[worst]["kernel_GBps"], f"b2_fused<id,{worst},C>", f"c3_{worst}", local_bytes),
           "parity": {"checked": True, "bit_exact": True, "vs": "np.argmax/argmin/max/min on every block of every rank"},
           "clocks": ctx.sampler.summary("c3"), "gpu_launches": sum(v["gpu_launches"] for v in lines.values())}
    if rank == 0 and W == 1 and not args.no_cpu_baseline:
        from oracle import reference as ref

        workers = os.cpu_count() or 1
        sample = mine[:2]
        xb = ref.Blocked({(k, 0): host[i] for k, i in enumerate(sample)}, ((RB,) * len(sample), (Cc,)))
        t0 = time.perf_counter()
        for fn in (ref.da_argmax, ref.da_argmin, ref.da_max, ref.da_min):
            fn(xb, axis=1, workers=workers)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 4 * len(sample) * RB * Cc * 8 / dt / 1e9, "unit": "GB/s", "cores": workers, "kind": "port",
                               "sample": f"{len(sample)} of 8 blocks (1 GiB each), the four reductions ({dt:.2f} s)",
                               "cpu": cpu_model(), "numpy": np.__version__}
    del x, host
    return out


# ----------------------------------------------------------------------------- c4
def run_c4(ctx: Ctx):
    """rechunk (16384,16384) (16384,256) -> (256,16384) and ``x.T + x`` (chunks 2048^2), f4 and f8, distinct
    values.  At N > 1 both cross the partition: stores into / loads from peer HBM over NVLink."""
    import dask_array_b200 as da
    from oracle import reference as ref

    args, W, rank = ctx.args, ctx.world, ctx.rank
    n = 16384

    def gen(dt, r0, nr, c0, nc):
        """value(i, j) = i * n + j (int32 bit patterns viewed as f4: distinct values)"""
        v = np.add.outer(np.arange(r0, r0 + nr, dtype=np.int64) * n, np.arange(c0, c0 + nc, dtype=np.int64))
        return v.astype(np.float64) if dt == np.float64 else v.astype(np.int32).view(np.float32)

    def source(dt, chunks):
        blk = lambda bid: gen(dt, bid[0] * chunks[0], chunks[0], bid[1] * chunks[1], chunks[1])
        return da.from_host_blocks(blk, (n, n), chunks, dt, token=f"c4-{np.dtype(dt).name}-{chunks}-w{W}").persist()

    lines = {}
    for dt in (np.float32, np.float64):
        item = np.dtype(dt).itemsize
        nm = np.dtype(dt).name
        nbytes = 2 * n * n * item
        x = source(dt, (n, 256))
        step = da.compile(x.rechunk((256, n)))
        ms, launches = ctx.timed(step, args.steps, args.warmup, "c4")
        ctx.torch.cuda.synchronize()
        ok = True
        for b in sorted(step.stores[0].blocks)[:2]:          # bit-exact: new row panels this rank owns vs the generator
            got = step.stores[0].blocks[b].to_numpy()
            ok &= np.array_equal(got.view(np.uint8), gen(dt, b[0] * 256, 256, 0, n).view(np.uint8))
        ctx.all_ok(ok, f"c4: rechunk {nm} bit-exact")
        lines[f"rechunk_{nm}"] = {"ms_per_step": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "bytes": nbytes,
                                  "gpu_launches": ctx.sum_over_ranks(launches)}
        del x, step
        sq = source(dt, (2048, 2048))
        step = da.compile(sq.T + sq)
        ms, launches = ctx.timed(step, args.steps, args.warmup, "c4")
        ctx.torch.cuda.synchronize()
        ok = True
        for b in sorted(step.stores[0].blocks)[:2]:
            got = step.stores[0].blocks[b].to_numpy()
            want = gen(dt, b[1] * 2048, 2048, b[0] * 2048, 2048).T + gen(dt, b[0] * 2048, 2048, b[1] * 2048, 2048)
            ok &= np.array_equal(got.view(np.uint8), want.view(np.uint8))
        ctx.all_ok(ok, f"c4: x.T + x {nm} bit-exact")
        lines[f"xT_plus_x_{nm}"] = {"ms_per_step": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "bytes": nbytes,
                                    "gpu_launches": ctx.sum_over_ranks(launches)}
        del sq, step
    total_ms = sum(v["ms_per_step"] for v in lines.values())
    total_bytes = sum(v["bytes"] for v in lines.values())
    worst = min(lines, key=lambda k: lines[k]["GBps"])
    out = {"workload": "c4: rechunk (16384,16384) (16384,256)->(256,16384) and x.T + x (chunks 2048^2), f4 and f8; "
                       "2*N*itemsize bytes each",
           "value": total_bytes / (total_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": total_ms, "n_gpus": W, "scaling": "strong",
           "dtype": "f32+f64", "per_op": lines,
           "parity": {"checked": True, "bit_exact": True, "vs": "the generator (two result blocks per rank per op)"},
           "clocks": ctx.sampler.summary("c4"), "gpu_launches": sum(v["gpu_launches"] for v in lines.values())}
    if W == 1:
        out["roofline"] = ctx.roofline(lines[worst]["GBps"], f"{worst} (b2_gather_bulk_kernel / b2_fused<add,ewt_sym>)",
                                       f"c4_{worst}", lines[worst]["bytes"])
    else:
        # (G-1)/G of the array crosses the partition; per-GPU traffic over NVLink
        for k, v in lines.items():
            cross = v["bytes"] / 2 * (W - 1) / W
            v["nvlink_bytes_per_step"] = cross
            v["nvlink_GBps_per_gpu"] = cross / W / (v["ms_per_step"] * 1e-3) / 1e9
        out["comm"] = os.environ.get("B2_COMM", "peer")
        out["roofline"] = {"bound": "nvlink", "achieved": lines[worst]["nvlink_GBps_per_gpu"], "peak": 900.0, "unit": "GB/s",
                           "frac": lines[worst]["nvlink_GBps_per_gpu"] / 900.0, "traffic": None, "kernel": worst,
                           "peak_source": "nominal NVLink 5 per direction per GPU"}
    if rank == 0 and W == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        xh = gen(np.float32, 0, n, 0, n)
        xb = ref.Blocked.from_array(xh, (n, 256))
        t0 = time.perf_counter()
        ref.rechunk(xb, (256, n), workers=workers)
        t1 = time.perf_counter()
        sb = ref.Blocked.from_array(xh, (2048, 2048))
        t2 = time.perf_counter()
        ref.elemwise(np.add, ref.transpose(sb), sb, workers=workers)
        t3 = time.perf_counter()
        dt_ = (t1 - t0) + (t3 - t2)
        out["cpu_baseline"] = {"value": 2 * (2 * n * n * 4) / dt_ / 1e9, "unit": "GB/s", "cores": workers, "kind": "port",
                               "sample": f"the f4 half of the workload: rechunk {t1 - t0:.2f} s + x.T + x {t3 - t2:.2f} s",
                               "cpu": cpu_model(), "numpy": np.__version__}
    return out


# ----------------------------------------------------------------------------- c5
def run_c5(ctx: Ctx):
    """``(x @ y.T).sum()`` with 32768^2 operands in 4096^2 chunks (8 x 8 x 8 block triples, 7.04e13 FLOP),
    bf16 operands and fp32 operands (bf16 x 3 split, six tensor-core products), distinct data per block."""
    import ml_dtypes
    import torch
    import dask_array_b200 as da

    args, W, rank = ctx.args, ctx.world, ctx.rank
    n, cb = 32768, 4096
    g = n // cb
    peak_b = float(ctx.peaks.get("bf16_tflops", 1650.0))
    peak_s = float(ctx.peaks.get("bf16_tflops_sustained", 1400.0))
    seeds = {"x": np.random.SeedSequence(5).spawn(g * g), "y": np.random.SeedSequence(6).spawn(g * g)}
    owned = [(i, j) for i in range(g) for j in range(g) if (i * g + j) % W == rank]
    host32 = {"x": {}, "y": {}}

    def gen(item):
        nm, bid = item
        r = np.random.Generator(np.random.PCG64(seeds[nm][bid[0] * g + bid[1]]))
        host32[nm][bid] = r.random((cb, cb), dtype=np.float32) - np.float32(0.5)

    with ThreadPoolExecutor(max_workers=max(1, ctx.threads)) as ex:
        list(ex.map(gen, [(nm, b) for nm in ("x", "y") for b in owned]))

    def as_bf16(a):
        return torch.from_numpy(a).to(torch.bfloat16).view(torch.uint16).numpy().view(ml_dtypes.bfloat16)

    flop = 2.0 * n * n * n
    lines = {}
    for label, dt in (("bf16", ml_dtypes.bfloat16), ("fp32", np.float32)):
        hx = {b: (as_bf16(v) if label == "bf16" else v) for b, v in host32["x"].items()}
        hy = {b: (as_bf16(v) if label == "bf16" else v) for b, v in host32["y"].items()}
        x = da.from_host_blocks(lambda bid: hx[bid], (n, n), (cb, cb), dt, token=f"c5x-{label}-w{W}").persist()
        y = da.from_host_blocks(lambda bid: hy[bid], (n, n), (cb, cb), dt, token=f"c5y-{label}-w{W}").persist()
        step = da.compile((x @ y.T).sum())
        reps = max(1, min(args.steps, 5 if label == "bf16" else 3))
        ms, launches = ctx.timed(step, reps, 2, "c5", graph=False)
        got = float(step.results()[0])
        # closed form of the result in fp64: sum_k (sum_i x[i,k]) * (sum_j y[j,k]); the bound of the
        # rounding error is relative to sum |x| . |y| (products of random signs cancel)
        cs = {}
        for nm, hh in (("x", hx), ("y", hy)):
            part = {}
            for (i, k), v in hh.items():
                v64 = v.astype(np.float32).astype(np.float64)
                s, a = part.setdefault(k, [0, 0])
                part[k] = [s + v64.sum(axis=0), a + np.abs(v64).sum(axis=0)]
            allp = ctx.gather_objects(part)
            tot = {}
            for p in allp:
                for k, (s, a) in p.items():
                    t = tot.setdefault(k, [0, 0])
                    tot[k] = [t[0] + s, t[1] + a]
            cs[nm] = tot
        want = sum(float(np.dot(cs["x"][k][0], cs["y"][k][0])) for k in range(g))
        bound = sum(float(np.dot(cs["x"][k][1], cs["y"][k][1])) for k in range(g))
        tol = 1e-5 if label == "fp32" else 1e-5
        err = abs(got - want) / bound
        ctx.all_ok(err <= tol, f"c5 {label}: (x @ y.T).sum() = {got!r}, fp64 closed form {want!r}, |err| / sum|x||y| = {err:.2e}")
        mma = flop * (6 if label == "fp32" else 1)
        lines[label] = {"ms_per_step": ms, "TFLOPs": flop / (ms * 1e-3) / 1e12, "tensor_pipe_TFLOPs": mma / (ms * 1e-3) / 1e12,
                        "steps": reps, "gpu_launches": ctx.sum_over_ranks(launches),
                        "parity": {"rel_err_vs_sum_abs_products": err, "tol": tol}}
        del x, y, step, hx, hy
    tp = lines["bf16"]["tensor_pipe_TFLOPs"] / W
    out = {"workload": "c5: (x @ y.T).sum(), 32768^2 operands, chunks 4096^2 (512 block GEMMs), bf16 and fp32 (bf16x3 split)",
           "metric": "blocked matmul TFLOP/s", "value": lines["bf16"]["TFLOPs"], "unit": "TFLOP/s (algorithmic 2N^3, bf16 operands)",
           "ms_per_step": lines["bf16"]["ms_per_step"], "n_gpus": W, "scaling": "strong", "dtype": "bf16 (fp32 accumulate)",
           "per_dtype": lines,
           "roofline": {"bound": "tensor", "achieved": tp, "peak": peak_s, "unit": "TFLOP/s", "frac": tp / peak_s,
                        "frac_of_burst_peak": tp / peak_b, "traffic": ctx.traffic.get("c5_gemm_bf16"),
                        "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained: the kernel runs inside a long step)",
                        "kernel": "b2_gemm_tn_batched_kernel (tcgen05 + TMA)", "flop_per_step": flop},
           "parity": {"checked": True, "vs": "fp64 closed form sum_k colsum(x)[k]*colsum(y)[k] on distinct per-block data",
                      "contract": "|got - want| <= 1e-5 * sum|x|.|y|"},
           "clocks": ctx.sampler.summary("c5"), "gpu_launches": sum(v["gpu_launches"] for v in lines.values())}
    if rank == 0 and W == 1 and not args.no_cpu_baseline:
        from oracle import reference as ref

        workers = os.cpu_count() or 1
        m = 8192
        a = np.concatenate([np.concatenate([host32["x"][(i, k)] for k in range(2)], axis=1) for i in range(2)], axis=0)
        b = np.concatenate([np.concatenate([host32["y"][(i, k)] for k in range(2)], axis=1) for i in range(2)], axis=0)
        ab, bb = ref.Blocked.from_array(a, (cb, cb)), ref.Blocked.from_array(np.ascontiguousarray(b.T), (cb, cb))
        t0 = time.perf_counter()
        ref.da_sum(ref.matmul(ab, bb, workers=workers), workers=workers)
        dt_ = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 2.0 * m**3 / dt_ / 1e12, "unit": "TFLOP/s", "cores": workers, "kind": "port",
                               "sample": f"scaled 8192^2 fp32 instance, chunks 4096^2 (8 block GEMMs + partial sums), {dt_:.2f} s "
                                         "(BASELINE.md: the reference materialises 32 GiB of partials at full size)",
                               "cpu": cpu_model(), "numpy": np.__version__}
    return out


# ----------------------------------------------------------------------------- cumulative scans
def run_cum(ctx: Ctx):
    """cumsum over fp32 32768^2 (both axes) and an fp64 vector of 2^28.  Algorithmic
    bytes: read N + write N."""
    import dask_array_b200 as da

    args, W = ctx.args, ctx.world
    if W > 1:
        return {"skipped": "single-GPU line"}
    n, cb = 32768, 4096
    g = n // cb
    seeds = np.random.SeedSequence(7).spawn(g * g)
    host = {}

    def gen(bid):
        host[bid] = np.random.Generator(np.random.PCG64(seeds[bid[0] * g + bid[1]])).random((cb, cb), dtype=np.float32)

    with ThreadPoolExecutor(max_workers=max(1, ctx.threads)) as ex:
        list(ex.map(gen, [(i, j) for i in range(g) for j in range(g)]))
    x = da.from_host_blocks(lambda bid: host[bid], (n, n), (cb, cb), np.float32, token="cum-f4").persist()
    lines = {}
    for axis in (0, 1):
        step = da.compile(x.cumsum(axis=axis))
        ms, launches = ctx.timed(step, args.steps, args.warmup, "cum")
        ctx.torch.cuda.synchronize()
        # parity sample: one line through all 8 blocks along the axis vs an fp64 cumsum
        j = 3
        if axis == 0:
            got = np.concatenate([step.stores[0].blocks[(i, j)][:, 17].to_numpy() for i in range(g)])
            want = np.cumsum(np.concatenate([host[(i, j)][:, 17] for i in range(g)]).astype(np.float64))
        else:
            got = np.concatenate([step.stores[0].blocks[(j, i)][17, :].to_numpy() for i in range(g)])
            want = np.cumsum(np.concatenate([host[(j, i)][17, :] for i in range(g)]).astype(np.float64))
        err = float(np.max(np.abs(got - want) / want))
        ctx.all_ok(err <= 1e-5, f"cum: fp32 cumsum(axis={axis}) vs fp64 (max rel err {err:.2e})")
        lines[f"f32_axis{axis}"] = {"ms_per_step": ms, "GBps": 2 * n * n * 4 / (ms * 1e-3) / 1e9, "bytes": 2 * n * n * 4,
                                    "gpu_launches": launches, "max_rel_err_vs_fp64": err}
        del step
    del x, host
    m = 1 << 28
    vs = np.random.SeedSequence(8).spawn(16)
    vh = {}

    def genv(i):
        vh[i] = np.random.Generator(np.random.PCG64(vs[i])).random(1 << 24)

    with ThreadPoolExecutor(max_workers=max(1, ctx.threads)) as ex:
        list(ex.map(genv, range(16)))
    v = da.from_host_blocks(lambda bid: vh[bid[0]], (m,), (1 << 24,), np.float64, token="cum-f8").persist()
    step = da.compile(v.cumsum())
    ms, launches = ctx.timed(step, args.steps, args.warmup, "cum")
    ctx.torch.cuda.synchronize()
    got = step.stores[0].blocks[(15,)][-1:].to_numpy()[0]
    want = float(sum(np.sum(vh[i]) for i in range(16)))
    err = abs(got - want) / want
    ctx.all_ok(err <= 1e-12, f"cum: fp64 vector cumsum total (rel err {err:.2e})")
    lines["f64_vector"] = {"ms_per_step": ms, "GBps": 2 * m * 8 / (ms * 1e-3) / 1e9, "bytes": 2 * m * 8, "gpu_launches": launches,
                           "total_rel_err": err}
    worst = min(lines, key=lambda k: lines[k]["GBps"])
    total_ms = sum(v_["ms_per_step"] for v_ in lines.values())
    total_bytes = sum(v_["bytes"] for v_ in lines.values())
    return {"workload": "cum: cumsum fp32 (32768,32768) chunks 4096^2 axis 0 and 1; fp64 vector 2^28 chunks 2^24 (2N bytes each)",
            "value": total_bytes / (total_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": total_ms, "n_gpus": 1, "per_op": lines,
            "roofline": ctx.roofline(lines[worst]["GBps"], f"b2_run_scan ({worst})", f"cum_{worst}", lines[worst]["bytes"]),
            "parity": {"checked": True, "vs": "fp64 cumsum of one line through all blocks (rtol 1e-5) / fp64 total (1e-12)"},
            "clocks": ctx.sampler.summary("cum"), "gpu_launches": sum(v_["gpu_launches"] for v_ in lines.values())}


def run_win(ctx: Ctx):
    """sliding-window (rolling) reductions over fp32 (32768,32768), chunks 4096^2 --
    ``sliding_window_view(x, w, axis).sum(axis=-1)`` for a short window on both axes and a window of 1024.
    Algorithmic bytes: read N + write N (the output is N minus a rim)."""
    import dask_array_b200 as da
    from oracle import reference as ref

    args, W = ctx.args, ctx.world
    if W > 1:
        return {"skipped": "single-GPU line"}
    n, cb = 32768, 4096
    g = n // cb
    seeds = np.random.SeedSequence(9).spawn(g * g)
    host = {}

    def gen(bid):
        host[bid] = np.random.Generator(np.random.PCG64(seeds[bid[0] * g + bid[1]])).random((cb, cb), dtype=np.float32)

    with ThreadPoolExecutor(max_workers=max(1, ctx.threads)) as ex:
        list(ex.map(gen, [(i, j) for i in range(g) for j in range(g)]))
    x = da.from_host_blocks(lambda bid: host[bid], (n, n), (cb, cb), np.float32, token="win-f4").persist()
    swv = np.lib.stride_tricks.sliding_window_view
    lines = {}
    for axis, w in ((0, 64), (1, 64), (0, 1024)):
        step = da.compile(da.sliding_window_view(x, w, axis=axis).sum(axis=-1))
        ms, launches = ctx.timed(step, args.steps, args.warmup, "win")
        ctx.torch.cuda.synchronize()
        # parity sample: one line through all blocks along the axis against NumPy's own window view in fp64
        j = 5
        if axis == 0:
            got = np.concatenate([step.stores[0].blocks[(i, j)][:, 33].to_numpy() for i in range(g)])
            line = np.concatenate([host[(i, j)][:, 33] for i in range(g)]).astype(np.float64)
        else:
            got = np.concatenate([step.stores[0].blocks[(j, i)][33, :].to_numpy() for i in range(g)])
            line = np.concatenate([host[(j, i)][33, :] for i in range(g)]).astype(np.float64)
        want = swv(line, w).sum(axis=-1)
        err = float(np.max(np.abs(got - want) / want))
        ctx.all_ok(got.shape == want.shape and err <= 1e-5, f"win: rolling sum axis={axis} w={w} vs fp64 (max rel err {err:.2e})")
        nbytes = n * n * 4 + (n - w + 1) * n * 4
        lines[f"sum_axis{axis}_w{w}"] = {"ms_per_step": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "bytes": nbytes,
                                         "gpu_launches": launches, "max_rel_err_vs_fp64": err}
        del step
    worst = min(lines, key=lambda k: lines[k]["GBps"])
    total_ms = sum(v["ms_per_step"] for v in lines.values())
    total_bytes = sum(v["bytes"] for v in lines.values())
    out = {"workload": "win: sliding_window_view(x, w, axis).sum(axis=-1), fp32 (32768,32768) chunks 4096^2: (axis 0, w 64), "
                       "(axis 1, w 64), (axis 0, w 1024); 2N bytes each",
           "value": total_bytes / (total_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": total_ms, "n_gpus": 1, "per_op": lines,
           "roofline": ctx.roofline(lines[worst]["GBps"], f"WindowHalo gather + b2_window_reduce ({worst})", f"win_{worst}",
                                    lines[worst]["bytes"],
                                    note="two launches per step: the halo gather (2N) and the window kernel (2N); algorithmic 2N"),
           "parity": {"checked": True, "vs": "NumPy sliding_window_view(...).sum(-1) in fp64 on one line through all blocks (rtol 1e-5)"},
           "clocks": ctx.sampler.summary("win"), "gpu_launches": sum(v["gpu_launches"] for v in lines.values())}
    if not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        xb = ref.Blocked({(i, 0): host[(i, 0)] for i in range(2)}, ((cb, cb), (cb,)))
        t0 = time.perf_counter()
        ref.da_sliding_window_reduce(xb, 64, 0, "sum", workers=workers)
        dt_ = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 2 * (2 * cb * cb * 4) / dt_ / 1e9, "unit": "GB/s", "cores": workers, "kind": "port",
                               "sample": f"2 of 64 blocks, rolling sum w=64 along axis 0 through NumPy's window view ({dt_:.2f} s)",
                               "cpu": cpu_model(), "numpy": np.__version__}
    del x, host
    return out


def run_c1(ctx: Ctx):
    """README example: latency only (80 kB blocks are not a roofline config)."""
    import dask_array_b200 as da
    from dask_array_b200 import _lib

    torch, args = ctx.torch, ctx.args
    x = da.ones((1000, 1000), chunks=(100, 100))
    out = {}
    for label, arr, want in (("(x + x.T)[:100, :100]", (x + x.T)[:100, :100], 2.0), ("(x + x.T).sum()", (x + x.T).sum(), 2e6)):
        t0 = time.perf_counter(); first = arr.compute(); cold = time.perf_counter() - t0
        t0 = time.perf_counter(); val = arr.compute(); full = time.perf_counter() - t0
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
        per = (_lib.launch_count() - n0) / args.steps
        step.capture()                                  # the same tape as ONE CUDA graph
        for _ in range(5):
            step.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step.run()
        torch.cuda.synchronize()
        gdt = (time.perf_counter() - t0) / args.steps
        ok = bool(np.all(np.asarray(val) == want)) and bool(np.all(np.asarray(first) == want))
        ctx.all_ok(ok, f"c1: {label}")
        out[label] = {"replay_us": dt * 1e6, "graph_replay_us": gdt * 1e6, "compute_call_ms": full * 1e3,
                      "first_call_ms": cold * 1e3, "launches_per_replay": per, "result": float(np.asarray(val).ravel()[0])}
    return {"workload": "c1: README example da.ones((1000,1000), chunks=(100,100)); latency only", "metric": "latency",
            "per_expr": out, "parity": {"checked": True, "bit_exact": True}}


CONFIGS = {"c1": run_c1, "c3": run_c3, "c4": run_c4, "c5": run_c5, "cum": run_cum, "win": run_win}


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="all", choices=["all", "c1", "c2", "c3", "c4", "c5", "cum", "win"],
                    help="all = headline c2 line carrying the other configs (default); cN = that config alone")
    ap.add_argument("--configs", default="c1,c3,c4,c5,cum,win", help="secondary configs carried by the default line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    ctx = Ctx(args)
    W, rank = ctx.world, ctx.rank
    from dask_array_b200 import _lib

    if args.config not in ("all", "c2"):
        out = CONFIGS[args.config](ctx)
        if rank == 0:
            print(json.dumps(out))
    else:
        main_line, shape = run_c2(ctx, strong=False, with_e2e=True, with_cpu=not args.no_cpu_baseline)
        cfg = c2_config(W)
        line = {"metric": METRIC, "value": main_line.pop("value"), "unit": "GB/s", "n_gpus": W, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": main_line.pop("ms_per_step"), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg}
        line.update(main_line)
        if W > 1:
            s, sshape = run_c2(ctx, strong=True, with_e2e=False, with_cpu=False)
            s["config"] = {"workload": "c2 on BASELINE's fixed array: fp32 (32768,32768) chunks 4096^2 dealt over the N GPUs",
                           "global_shape": list(sshape), "blocks_per_gpu": GRID * GRID // W}
            s["efficiency_basis"] = "value at n_gpus=1 of this bench (same 64-block workload) x N"
            line["strong"] = s
        else:
            line["strong"] = {"value": line["value"], "ms_per_step": line["ms_per_step"],
                              "note": "at N=1 the strong workload IS the headline workload (64 blocks)"}
        if args.config == "all":
            extra = {}
            for name in [c for c in args.configs.split(",") if c]:
                if W > 1 and name in ("c1", "cum", "win"):
                    continue
                try:
                    extra[name] = CONFIGS[name](ctx)
                except SystemExit:
                    raise
                except Exception as e:          # noqa: BLE001  (a secondary config must not take the headline down)
                    extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            line["configs"] = extra
        if rank == 0:
            print(json.dumps(line))
    ctx.sampler.stop()
    ctx.sampler.join(timeout=2)
    if W > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
