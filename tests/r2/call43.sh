#!/bin/bash
# round 2, session 3, call 4: wide-landmark Schur path (forced by G2O_B200_SR_WIDE in the parametrised test) + BA parity tests
out=gpurun_out/r2b
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_widening_expmap.py tests/test_robust.py -q -m gpu -k "bundle or ba or BA or venice or expmap or robust" 2>&1 | tail -40 > $out/c43_pytest.txt
tail -15 $out/c43_pytest.txt
