#!/bin/bash
# round 2, pass 1 (1 GPU): the -m gpu suite, then the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench1 rc=$?"
tail -c 800 gpurun_out/r2_bench_n1.err
head -c 3000 gpurun_out/r2_bench_n1.json
