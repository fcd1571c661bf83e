#!/bin/bash
# round 2, pass 16 (8 GPUs): the c2 line at N = 8, 4 after the batched two-stage folds (N=1 reference on the same box)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
  timeout 600 $RUN --nproc-per-node $n --master-port 2959$n bench.py --gpus $n --steps 20 --warmup 5 --config c2 > gpurun_out/r2_fold_n$n.json 2> gpurun_out/r2_fold_n$n.err; echo "n=$n rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --config c2 > gpurun_out/r2_fold_n1.json 2>/dev/null; echo "n=1 rc=$?"
python - <<'PY'
import json
base = None
for n in (1, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/r2_fold_n{n}.json").read().strip().splitlines()[-1])
        s = d["strong"]
        if n == 1:
            base = d["ms_per_step"]
        ks = [(k["kernel"][-10:], round(k["ms"], 4)) for k in s.get("roofline", {}).get("kernels", [])]
        print(n, "weak", round(d["value"]), round(d["ms_per_step"], 4), "eff", round(base / d["ms_per_step"], 3),
              "| strong", round(s["value"]), round(s["ms_per_step"], 4), "eff", round(base / n / s["ms_per_step"], 3), ks)
    except Exception as e:
        print(n, "unreadable", e)
PY
