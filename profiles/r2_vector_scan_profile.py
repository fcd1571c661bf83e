"""The fp64 vector cumsum of bench.py's `cum` config alone (2^28 elements in 2^24 chunks), for one ncu capture of
its three launches (totals, scan of the totals, scan with carry)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dask_array_b200 as da

m = 1 << 28
vh = np.random.default_rng(1).random(1 << 24)
v = da.from_host_blocks(lambda bid: vh, (m,), (1 << 24,), np.float64, token="prof-cum-f8").persist()
step = da.compile(v.cumsum())
for _ in range(3):
    step.run()
torch.cuda.synchronize()
