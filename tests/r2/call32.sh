#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "pcg or pieces or wide" 2>&1 | tail -15
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee $O/c32_pytest_gpu.txt
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "pcg or pieces or wide_tile_tensor_path_matches_dense" > $O/c32_sanitizer_memcheck.txt 2>&1; tail -4 $O/c32_sanitizer_memcheck.txt
