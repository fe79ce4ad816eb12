#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for w in 0 1; do
  for wl in sphere2500 venice; do
    G2O_B200_SPLIT_LATE=$w timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c29_${wl}_s$w.json 2> $O/c29_${wl}_s$w.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c29_${wl}_s$w.json") if l.startswith("{")][-1])
print("$wl split_late=$w value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1], {k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items() if "chol" in k})
PY
  done
  G2O_B200_SPLIT_LATE=$w timeout 600 python tests/config5_probe.py 200 500 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('split_late $w', d['poses'], d.get('iteration_s'), d.get('phases_ms', {}).get('chol_factor_flow'), d['factor']['flow_tasks'], d.get('error'))
"
done 2>&1 | tee $O/c29_split_late.txt
