#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for v in p1t1 p0t1 p1t0; do echo "=== variant $v (p = priority barrier, t = TRSM preload)"; G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing_$v.so timeout 600 python tests/chain_timing.py ba10k | grep -v "^ba10k"; done 2>&1 | tee $O/c12_chain_timing_variants.txt
