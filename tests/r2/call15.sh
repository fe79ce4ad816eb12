#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for w in sphere200 sphere300; do G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing.so timeout 300 python tests/chol_timing.py $w; done 2>&1 | tee $O/c15_flow_timing.txt
