#!/bin/bash
# round 2, session 3, call 1: new GPU tests (sparse-inverse marginals, landmark SLAM) without -x, then the full GPU suite
out=gpurun_out/r2b
mkdir -p $out
timeout 600 python -m pytest tests/test_widening_landmark_slam.py "tests/test_gpu_parity.py::test_marginals_match_dense_inverse_of_the_oracle_hessian" -q -m gpu 2>&1 | tail -60 > $out/c40_new_tests.txt
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > $out/c40_pytest_gpu.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/c40_bench_venice.json 2> $out/c40_bench_venice.err
tail -5 $out/c40_new_tests.txt; tail -5 $out/c40_pytest_gpu.txt; cut -c1-600 $out/c40_bench_venice.json
