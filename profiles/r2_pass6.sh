#!/bin/bash
# round 2, pass 6 (1 GPU): scan prefetch variants
mkdir -p gpurun_out
for cfg in "4 32" "2 48" "1 64" "1 32"; do
  set -- $cfg
  B2_SCAN_SR_VEC=$1 B2_SCAN_SR_UNROLL=$2 timeout 300 python bench.py --config cum --steps 20 > gpurun_out/r2_cum_v$1_u$2.json 2>gpurun_out/r2_cum_v$1_u$2.err
  python - "$1" "$2" <<'PY'
import json, sys
f = f"gpurun_out/r2_cum_v{sys.argv[1]}_u{sys.argv[2]}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k: (round(v["GBps"]), round(v["ms_per_step"], 3)) for k, v in d["per_op"].items()})
except Exception as e:
    print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-300:])
PY
done
timeout 300 python bench.py --config cum --steps 20 > gpurun_out/r2_bench_cum.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_cum.json').read().strip().splitlines()[-1]); print('default', {k:(round(v['GBps']),round(v['ms_per_step'],3)) for k,v in d['per_op'].items()})"
timeout 600 python -m pytest tests/test_gpu_cumulative.py -m gpu -q -x 2>&1 | tail -2
