#!/bin/bash
# round 2, pass 4 (1 GPU): single-pass scans (rebuilt header) and the CTA-pair GEMM, each under its own timeout
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cumulative.py -m gpu -q -x > gpurun_out/r2_pytest_pass4_cum.log 2>&1; echo "cum pytest rc=$?"; tail -3 gpurun_out/r2_pytest_pass4_cum.log
timeout 300 python bench.py --config cum --steps 20 > gpurun_out/r2_bench_cum.json 2> gpurun_out/r2_bench_cum.err; echo "cum rc=$?"; tail -c 400 gpurun_out/r2_bench_cum.err
timeout 300 python -m pytest tests/test_gpu_matmul.py tests/test_gpu_contract.py -m gpu -q -x > gpurun_out/r2_pytest_pass4_gemm.log 2>&1; echo "gemm pytest rc=$?"; tail -5 gpurun_out/r2_pytest_pass4_gemm.log
timeout 400 python bench.py --config c5 --steps 5 > gpurun_out/r2_bench_c5_pair.json 2> gpurun_out/r2_bench_c5_pair.err; echo "c5 pair rc=$?"; tail -c 400 gpurun_out/r2_bench_c5_pair.err
B2_GEMM_2CTA=0 timeout 400 python bench.py --config c5 --steps 5 > gpurun_out/r2_bench_c5_single.json 2>/dev/null; echo "c5 single rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_cum.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: (round(v["GBps"]), round(v["ms_per_step"], 3)) for k, v in d["per_op"].items()})
    except Exception as e:
        print(f, "unreadable", e)
for f in ("gpurun_out/r2_bench_c5_pair.json", "gpurun_out/r2_bench_c5_single.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: (round(v["tensor_pipe_TFLOPs"]), round(v["ms_per_step"], 2)) for k, v in d["per_dtype"].items()}, d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
