#!/bin/bash
# round 2, GPU call 8: chain backward with packed records; inner-loop variants of the rank-6 update
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tail_chain or bundle" 2>&1 | tail -3
for v in 0 1 2; do echo "=== update variant $v"; G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing_v$v.so timeout 600 python tests/chain_timing.py ba10k; done 2>&1 | tee $O/c8_chain_timing_variants.txt
G2O_B200_LIB=openslam_g2o_b200/libg2o_b200_timing_v0.so timeout 600 python tests/chain_timing.py venice 2>&1 | tee -a $O/c8_chain_timing_variants.txt
