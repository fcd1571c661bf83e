#!/bin/bash
# round 2, pass 13 (1 GPU): GEMM pair-list chunking sweep; ncu traffic of the current kernels (c2, cum, win)
mkdir -p gpurun_out
for ch in 8 4 2; do
  B2_GEMM_CHUNK=$ch timeout 300 python bench.py --config c5 --steps 5 --no-cpu-baseline > gpurun_out/r2_c5_ch$ch.json 2>/dev/null
  python - $ch <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_c5_ch{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("chunk", sys.argv[1], {k: (round(v["tensor_pipe_TFLOPs"]), round(v["ms_per_step"], 2)) for k, v in d["per_dtype"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("chunk", sys.argv[1], "unreadable", e)
PY
done
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:b2_fused -s 12 -c 2 -o gpurun_out/prof_r2_c2 python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 600 $NCU -k regex:"b2_window|b2_gather" -s 6 -c 4 -o gpurun_out/prof_r2_win python bench.py --config win --steps 1 --no-cpu-baseline > gpurun_out/ncu_r2_win.log 2>&1; echo "ncu win rc=$?"
timeout 600 $NCU -k regex:b2_fused -s 4 -c 3 -o gpurun_out/prof_r2_cum python bench.py --config cum --steps 1 > gpurun_out/ncu_r2_cum.log 2>&1; echo "ncu cum rc=$?"
B2_BENCH_GRAPH=0 python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1 && B2_BENCH_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2_launch.log 2>&1; echo "launch list rc=$?"
