#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
for asap in 0 1; do
G2O_B200_GROUPS_ASAP=$asap timeout 600 python tests/config5_probe.py 300 500 2>&1 | tee $O/c21_config5_asap$asap.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('asap', $asap, d['poses'], d.get('iteration_s'), d.get('phases_ms', {}).get('chol_factor_flow'), d.get('error'))
"
done
for wl in venice sphere2500; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c21_${wl}.json 2> $O/c21_${wl}.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c21_${wl}.json") if l.startswith("{")][-1])
print("$wl value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1])
PY
done
