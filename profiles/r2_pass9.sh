#!/bin/bash
# round 2, pass 9 (1 GPU): sliding-window bench line, then the default line
mkdir -p gpurun_out
timeout 300 python bench.py --config win --steps 20 > gpurun_out/r2_bench_win.json 2> gpurun_out/r2_bench_win.err; echo "win rc=$?"; tail -c 500 gpurun_out/r2_bench_win.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_win.json').read().strip().splitlines()[-1]); print({k:(round(v['GBps']),round(v['ms_per_step'],3)) for k,v in d['per_op'].items()}, d.get('cpu_baseline',{}).get('value'))"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench1 rc=$?"; tail -c 500 gpurun_out/r2_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['first_call_ms'], d['compute_call_ms'], {k:(v.get('value'), v.get('error')) for k,v in d['configs'].items()})"
