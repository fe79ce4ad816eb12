#!/bin/bash
out=gpurun_out/r2b
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "marginals or full_size_venice_matches" 2>&1 | tail -3
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > $out/c53_bench.json 2> $out/c53_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b/c53_bench.json").read().strip().splitlines()[-1])
k = d["kernel_groups_ms_per_10_iterations"]
print("value", round(d["value"], 1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1), "chain", round(k["chol_chain"]["ms_total"] / 10, 4), "factor", round(k["factor"]["ms_total"] / 10, 4))
PY
