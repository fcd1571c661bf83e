#!/bin/bash
# round 2, pass 19 (2 GPUs): the tests added since the last full run + the multi-GPU check (top-k across ranks)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_api.py tests/test_gpu_multigpu.py tests/test_gpu_topk.py tests/test_gpu_kernels.py tests/test_gpu_cumulative.py -m gpu -q > gpurun_out/r2_pytest_pass19.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_pass19.log
