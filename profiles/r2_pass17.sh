#!/bin/bash
for r in 16 32 64 80 128 256; do B2_RPT=$r python profiles/r2_small_launch_sweep.py 2>&1 | tail -1; done
for r in 64 256; do NBLK=64 B2_RPT=$r python profiles/r2_small_launch_sweep.py 2>&1 | tail -1; done
NBLK=1 python profiles/r2_small_launch_sweep.py 2>&1 | tail -1
