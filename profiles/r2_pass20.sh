#!/bin/bash
# round 2, pass 20 (2 GPUs): lane priority on / off, c2 line at N=2 and N=1; lanes test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -q -k "lanes or fused_chain" 2>&1 | tail -2
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for p in 1 0; do
  B2_LANE_PRIORITY=$p timeout 600 $RUN --nproc-per-node 2 --master-port 2960$p bench.py --gpus 2 --steps 50 --warmup 5 --config c2 > gpurun_out/r2_c2_n2_p$p.json 2>/dev/null; echo "n=2 prio=$p rc=$?"
  B2_LANE_PRIORITY=$p timeout 600 python bench.py --steps 50 --warmup 5 --config c2 > gpurun_out/r2_c2_n1_p$p.json 2>/dev/null; echo "n=1 prio=$p rc=$?"
done
python - <<'PY'
import json
for f in ("r2_c2_n1_p1", "r2_c2_n1_p0", "r2_c2_n2_p1", "r2_c2_n2_p0"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "weak", round(d["value"]), round(d["ms_per_step"], 4), "strong", round(d["strong"]["value"]), round(d["strong"]["ms_per_step"], 4))
    except Exception as e:
        print(f, "unreadable", e)
PY
