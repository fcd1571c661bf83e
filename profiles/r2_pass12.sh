#!/bin/bash
# round 2, pass 12 (2 GPUs): graph lanes -- api tests, multi-GPU check, c2 line at N=1 and N=2 with and without lanes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_multigpu.py -m gpu -q -x 2>&1 | tail -3
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for lanes in 1 0; do
  B2_GRAPH_LANES=$lanes timeout 600 python bench.py --steps 20 --warmup 5 --config c2 > gpurun_out/r2_c2_n1_l$lanes.json 2>/dev/null; echo "n=1 lanes=$lanes rc=$?"
  B2_GRAPH_LANES=$lanes timeout 600 $RUN --nproc-per-node 2 --master-port 2957$lanes bench.py --gpus 2 --steps 20 --warmup 5 --config c2 > gpurun_out/r2_c2_n2_l$lanes.json 2>/dev/null; echo "n=2 lanes=$lanes rc=$?"
done
python - <<'PY'
import json
for f in ("r2_c2_n1_l1", "r2_c2_n1_l0", "r2_c2_n2_l1", "r2_c2_n2_l0"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "weak", round(d["value"]), round(d["ms_per_step"], 4), "strong", round(d["strong"]["value"]), round(d["strong"]["ms_per_step"], 4), d["parity"]["checked"])
    except Exception as e:
        print(f, "unreadable", e)
PY
