#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
for w in 0 1; do
    G2O_B200_WIDE_TILES=$w timeout 300 python bench.py --workload sphere2500 --steps 40 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c27_sphere_w$w.json 2> $O/c27_sphere_w$w.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c27_sphere_w$w.json") if l.startswith("{")][-1])
print("sphere2500 wide=$w value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1])
print({k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items()})
PY
done
