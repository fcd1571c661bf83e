#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --config win --steps 20 > gpurun_out/r2_bench_win.json 2> gpurun_out/r2_bench_win.err; echo "win rc=$?"; tail -c 500 gpurun_out/r2_bench_win.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_win.json').read().strip().splitlines()[-1]); print({k:(round(v['GBps']),round(v['ms_per_step'],3)) for k,v in d['per_op'].items()}, d.get('cpu_baseline',{}).get('value'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_win_launches.csv python bench.py --config win --steps 1 --no-cpu-baseline > /dev/null 2>&1; grep -E "b2_window|b2_gather" gpurun_out/r2_win_launches.csv | tail -8 | cut -c1-220
