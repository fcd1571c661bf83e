#!/bin/bash
# round 2, pass 7 (8 GPUs): multi-GPU parity check, then the default bench line at N=8 and N=4
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $RUN --nproc-per-node 8 --master-port 29551 tests/multigpu_check.py > gpurun_out/r2_multigpu_check_peer_n8.log 2>&1; echo "mg8 rc=$?"; tail -2 gpurun_out/r2_multigpu_check_peer_n8.log
timeout 900 $RUN --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc=$?"; tail -c 800 gpurun_out/r2_bench_n8.err
timeout 900 $RUN --nproc-per-node 4 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 5 --configs c4 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench4 rc=$?"; tail -c 400 gpurun_out/r2_bench_n4.err
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1; lscpu | head -25 >> gpurun_out/r2_topo_n8.txt
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_n8.json", "gpurun_out/r2_bench_n4.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "weak", round(d["value"]), round(d["ms_per_step"], 4), "strong", round(d["strong"]["value"]), round(d["strong"]["ms_per_step"], 4),
              "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1), d["e2e"].get("numa"))
        for k, v in d.get("configs", {}).items():
            print("  ", k, v.get("value"), v.get("ms_per_step"), v.get("error"))
    except Exception as e:
        print(f, "unreadable", e)
PY
