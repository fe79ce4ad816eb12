#!/bin/bash
# round 2, GPU call 11: ncu --set full of the chain kernels (source-level stall reasons) on the Venice-shaped graph
O=gpurun_out/r2; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chol_chain -c 3 -o $O/c11_ncu_chain python tests/prof_run.py venice 1 > $O/c11_ncu.log 2>&1
tail -5 $O/c11_ncu.log; ls -la $O/c11_ncu_chain.ncu-rep
