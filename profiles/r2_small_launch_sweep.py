"""Per-rank workload of BASELINE config 2 dealt over 8 GPUs (8 blocks of 4096^2 fp32 = 512 MiB): duration of the
two chunk kernels as a function of the tile height (B2_RPT) -- run once per value (the geometry is read at plan
time): `for r in 32 64 80 128 256; do B2_RPT=$r python profiles/r2_small_launch_sweep.py; done`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dask_array_b200 as da

nblk = int(os.environ.get("NBLK", "8"))
base = np.random.default_rng(0).random((4096, 4096), dtype=np.float32)
x = da.from_host_blocks(lambda bid: base, (4096 * nblk, 4096), (4096, 4096), np.float32, token=f"sweep{nblk}").persist()
y = da.sin(x) * 2 + x**2
out = {}
for name, expr in (("mean", y.mean(axis=0)), ("std", y.std())):
    step = da.compile(expr)
    k = max(step.fused_launches(), key=lambda f: f.total_tiles)
    for _ in range(5):
        step.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(20):
        e0.record(); k.run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[name] = (round(float(np.median(ts)) * 1e3, 1), k.total_tiles, k.spec.rpt)
print("RPT", os.environ.get("B2_RPT"), "U", os.environ.get("B2_U"), "nblk", nblk, out)
