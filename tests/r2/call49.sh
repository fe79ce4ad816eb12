#!/bin/bash
# round 2, session 3: per-edge robust kernels - full GPU suite with the changed kernel signatures + short bench (no regression)
out=gpurun_out/r2b
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $out/c49_pytest_gpu.txt
tail -4 $out/c49_pytest_gpu.txt
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $out/c49_bench.json 2> $out/c49_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b/c49_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1))
print({k: round(v["ms_total"] / 10, 4) for k, v in d["kernel_groups_ms_per_10_iterations"].items()})
PY
