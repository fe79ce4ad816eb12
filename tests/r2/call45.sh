#!/bin/bash
# round 2, session 3, call 6: PCG inside the Level-2/3 solver + the refactored Level-1 PCG
out=gpurun_out/r2b
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "pcg" 2>&1 | tail -40 > $out/c45_pytest.txt
tail -25 $out/c45_pytest.txt
