"""Multi-GPU parity under pytest: when the box shows >= 2 devices, launch ``tests/multigpu_check.py``
under ``torch.distributed.run`` (one rank per GPU, NCCL rendezvous on 127.0.0.1) and require it to pass:
tree reductions with the peer-memory exchange (also replayed from a CUDA graph), rechunk / transposed
reads / views / halos / scans / blocked matmul across the partition, all against the oracle or NumPy.
The log is kept under ``gpurun_out/`` so it travels back from the GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("comm", ["peer", "nccl"])
def test_multigpu_check(comm):
    n = _ndev()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    env = dict(os.environ, B2_COMM=comm)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533" if comm == "peer" else "29534",
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"multigpu_check_{comm}_n{n}.log"), "w") as f:
        f.write(p.stdout[-20000:] + "\n--- stderr ---\n" + p.stderr[-20000:])
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert f"multigpu_check ok on {n} GPUs" in p.stdout
