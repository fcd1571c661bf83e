#!/bin/bash
for u in 4 8 16; do for r in 96 128 192; do B2_U=$u B2_RPT=$r python profiles/r2_small_launch_sweep.py 2>&1 | tail -1; done; done
