#!/bin/bash
# round 2, pass 14 (8 GPUs): final numbers -- the default line at N=8, the c2 line at N=4, 2, 1 on the same box
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $RUN --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_final_n8.json 2> gpurun_out/r2_final_n8.err; echo "n=8 rc=$?"; tail -c 400 gpurun_out/r2_final_n8.err | grep -v OMP
for n in 4 2; do
  timeout 600 $RUN --nproc-per-node $n --master-port 2958$n bench.py --gpus $n --steps 20 --warmup 5 --config c2 > gpurun_out/r2_final_n$n.json 2> gpurun_out/r2_final_n$n.err; echo "n=$n rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 5 --config c2 > gpurun_out/r2_final_n1_box8.json 2> gpurun_out/r2_final_n1_box8.err; echo "n=1 rc=$?"
python - <<'PY'
import json
base = None
for n, f in ((1, "r2_final_n1_box8"), (2, "r2_final_n2"), (4, "r2_final_n4"), (8, "r2_final_n8")):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        s = d["strong"]
        if n == 1:
            base = d["ms_per_step"]
        print(n, "weak", round(d["value"]), round(d["ms_per_step"], 4), "eff", round(base / d["ms_per_step"], 3),
              "| strong", round(s["value"]), round(s["ms_per_step"], 4), "eff", round(base / n / s["ms_per_step"], 3),
              "e2e", round(d["e2e"]["ms_per_step"], 1), d["parity"]["checked"], s.get("parity", {}).get("checked"))
        for k, v in d.get("configs", {}).items():
            print("   ", k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    except Exception as e:
        print(n, "unreadable", e)
PY
