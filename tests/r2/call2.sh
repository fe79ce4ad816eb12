#!/bin/bash
# round 2, GPU call 2: FP64 pipe microbenchmark, full GPU suite (incl. the full-size parity gate), new bench line (N = 1)
O=gpurun_out/r2; mkdir -p $O
tests/csrc/fp64_pipes > $O/fp64_pipes.json 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > $O/c2_pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/c2_bench_n1.json 2> $O/c2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/c2_bench_ref.json 2> $O/c2_bench_ref.err
cat $O/fp64_pipes.json; cat $O/c2_pytest_gpu.txt; tail -c 600 $O/c2_bench_n1.err; tail -c 400 $O/c2_bench_ref.json
