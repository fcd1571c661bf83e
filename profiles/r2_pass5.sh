#!/bin/bash
# round 2, pass 5 (1 GPU): ncu on the GEMM variants and the scan kernels; compute-sanitizer on the kernel tests
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
C5="python bench.py --config c5 --steps 1 --no-cpu-baseline"
timeout 600 $NCU -k regex:gemm_tn_pair -s 2 -c 1 -o gpurun_out/prof_r2_c5_pair $C5 > gpurun_out/ncu_r2_c5_pair.log 2>&1; echo "ncu pair rc=$?"
B2_GEMM_2CTA=0 timeout 600 $NCU -k regex:gemm_tn_batched -s 2 -c 1 -o gpurun_out/prof_r2_c5_single $C5 > gpurun_out/ncu_r2_c5_single.log 2>&1; echo "ncu single rc=$?"
CUM="python bench.py --config cum --steps 1"
timeout 600 $NCU -k regex:b2_fused -s 4 -c 3 -o gpurun_out/prof_r2_cum $CUM > gpurun_out/ncu_r2_cum.log 2>&1; echo "ncu cum rc=$?"
# compute-sanitizer: memcheck over the kernel-level tests, racecheck + synccheck over the
# shared-memory staged kernels (mirror pair, scans); summaries kept
SAN="compute-sanitizer --error-exitcode 9 --launch-timeout 120"
timeout 900 $SAN --tool memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/san_r2_memcheck_kernels.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_r2_memcheck_kernels.log
timeout 900 $SAN --tool racecheck python -m pytest tests/test_gpu_mirror_pair.py -m gpu -q -x > gpurun_out/san_r2_racecheck_mirror.log 2>&1; echo "racecheck mirror rc=$?"; tail -4 gpurun_out/san_r2_racecheck_mirror.log
timeout 900 $SAN --tool racecheck python -m pytest tests/test_gpu_cumulative.py -m gpu -q -x -k "golden or matrix" > gpurun_out/san_r2_racecheck_cum.log 2>&1; echo "racecheck cum rc=$?"; tail -4 gpurun_out/san_r2_racecheck_cum.log
timeout 600 $SAN --tool synccheck python -m pytest tests/test_gpu_mirror_pair.py tests/test_gpu_matmul.py -m gpu -q -x > gpurun_out/san_r2_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/san_r2_synccheck.log
ls -la gpurun_out/*.ncu-rep
