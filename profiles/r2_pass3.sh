#!/bin/bash
# round 2, pass 3 (1 GPU): tests touched since pass 2 + the single-pass scans
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_api.py tests/test_plugin.py tests/test_gpu_contract.py tests/test_gpu_cumulative.py tests/test_gpu_widening.py tests/test_gpu_chunk_protocols.py tests/test_gpu_matmul.py -m gpu -q > gpurun_out/r2_pytest_pass3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_pass3.log
tail -12 gpurun_out/r2_pytest_pass3.log
timeout 300 python bench.py --config cum --steps 20 > gpurun_out/r2_bench_cum.json 2> gpurun_out/r2_bench_cum.err; echo "cum rc=$?"; tail -c 600 gpurun_out/r2_bench_cum.err
B2_SCAN_CHAINED=0 timeout 300 python bench.py --config cum --steps 20 > gpurun_out/r2_bench_cum_3pass.json 2>/dev/null
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_cum.json", "gpurun_out/r2_bench_cum_3pass.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: (round(v["GBps"]), round(v["ms_per_step"], 3)) for k, v in d["per_op"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
