#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "wide_tile" 2>&1 | tail -15
for w in sphere200 sphere300; do G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing.so timeout 300 python tests/chol_timing.py $w 2>&1 | grep -v "^chunk\|^last\|^whole"; done 2>&1 | tee $O/c16_flow_timing_wide.txt
