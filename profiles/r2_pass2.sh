#!/bin/bash
# round 2, pass 2 (2 GPUs): the whole -m gpu suite (incl. tests/test_gpu_multigpu.py), bench at N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -15 gpurun_out/r2_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r2_bench_n2.err
head -c 2500 gpurun_out/r2_bench_n2.json
