#!/bin/bash
# round 2, session 3 (2 GPUs): sharded parity tests with the final kernels, Venice bench at N = 2 (chi2 trajectory equal to N = 1)
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_distributed.py -q -m gpu 2>&1 | tail -15 > $O/c48_pytest_dist.txt
cat $O/c48_pytest_dist.txt | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/c48_bench_n2.json 2> $O/c48_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b/c48_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "chi2", d.get("chi2_first_run"))
    print("config4", {k: (v.get("value"), v.get("ms_per_step"), v.get("chi2_first_run")) for k, v in (d.get("configs") or {}).items()})
except Exception as e:
    print("parse failed", e)
PY
tail -c 600 $O/c48_bench_n2.err
