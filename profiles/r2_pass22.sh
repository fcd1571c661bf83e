#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:b2_window_cols -s 1 -c 1 -o gpurun_out/prof_r2_win_cols python bench.py --config win --steps 1 --no-cpu-baseline > gpurun_out/ncu_r2_win_cols.log 2>&1; echo "ncu win cols rc=$?"
timeout 200 $NCU -k regex:b2_fused -s 3 -c 4 -o gpurun_out/prof_r2_cum_vector python profiles/r2_vector_scan_profile.py > gpurun_out/ncu_r2_cum_vector.log 2>&1; echo "ncu vector rc=$?"
