#!/bin/bash
# round 2, GPU call 3 (2 GPUs): sharded parity tests (native NCCL + host callback), bench at N = 2 (Venice + config 4)
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_distributed.py -q -m gpu -x 2>&1 | tail -15 > $O/c3_pytest_dist.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/c3_bench_n2.json 2> $O/c3_bench_n2.err
cat $O/c3_pytest_dist.txt; tail -c 1500 $O/c3_bench_n2.err; head -c 600 $O/c3_bench_n2.json
