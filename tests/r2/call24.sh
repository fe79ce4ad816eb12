#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_factor_flow -c 1 -o $O/c24_ncu_wide python tests/prof_run.py sphere500 1 > $O/c24_ncu.log 2>&1
tail -2 $O/c24_ncu.log
ncu -i $O/c24_ncu_wide.ncu-rep --page raw --csv > $O/c24_raw.csv 2>/dev/null
ls -la $O/c24_ncu_wide.ncu-rep
