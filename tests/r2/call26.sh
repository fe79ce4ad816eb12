#!/bin/bash
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee $O/c26_pytest_gpu.txt
for wl in venice; do
    timeout 300 python bench.py --workload $wl --steps 40 --warmup 5 --no-extras --no-cpu-baseline --no-parallel-ordering > $O/c26_${wl}.json 2> $O/c26_${wl}.err
    python - <<PY
import json
d=json.loads([l for l in open("$O/c26_${wl}.json") if l.startswith("{")][-1])
print("$wl value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "chi2", d["chi2_first_run"][-1])
print({k: round(v["ms_total"],3) for k,v in d[[k for k in d if k.startswith("kernel_groups")][0]].items()})
PY
done
timeout 300 ncu --set full --clock-control none -k regex:schur_range -c 1 -o $O/c26_ncu_schur python tests/prof_run.py venice 1 > $O/c26_ncu.log 2>&1
ncu -i $O/c26_ncu_schur.ncu-rep --page raw --csv > $O/c26_schur_raw.csv 2>/dev/null
