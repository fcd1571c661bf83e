#!/bin/bash
# round 2, pass 15 (1 GPU): full verification -- build check, smoke, the whole -m gpu suite, reference arm, default line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_final.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_n1_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"])
for k, v in d["configs"].items():
    print("  ", k, v.get("value"), v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), v.get("error"), (v.get("clocks") or {}).get("reasons"))
PY
