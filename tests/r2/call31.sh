#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for cfg in "G2O_B200_RELAX=1" "G2O_B200_RELAX=0" "G2O_B200_RELAX_FRAC=0.1" "G2O_B200_RELAX_FRAC=0.5"; do
    env $cfg timeout 300 python bench.py --workload venice --steps 40 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c31_venice.json 2> $O/c31_venice.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c31_venice.json") if l.startswith("{")][-1])
print("venice $cfg value", d["value"], "ms", d["ms_per_step"], {k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items() if "chol" in k}, d["factor"]["chain_links"], d["factor"]["factor_flops"], d["factor"]["chain_flops"])
PY
done 2>&1 | tee $O/c31_relax.txt
for gi in 16 32 64; do
G2O_B200_GROUP_ITEMS=$gi timeout 600 python tests/config5_probe.py 300 500 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('group_items $gi', d['poses'], d.get('iteration_s'), d['factor']['flow_tasks'], d['factor']['split_tile_slots'], d.get('error'))
"
done 2>&1 | tee $O/c31_group_items.txt
