#!/bin/bash
# round 2, session 3, call 3: final-state evidence: full GPU suite, default bench line (extras + widening + CPU leg), reference arm,
# ncu launch lists of the two headline workloads, ncu --set full of the lane-per-observation linearisation
out=gpurun_out/r2b
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $out/c42_pytest_gpu.txt
tail -3 $out/c42_pytest_gpu.txt
timeout 900 python bench.py > $out/c42_bench_default.json 2> $out/c42_bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b/c42_bench_default.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1))
    print("hbm", {k: (round(v["avg_launch_ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["hbm_kernels"].items()})
    print("widening", json.dumps(d.get("widening"))[:1500])
    print("configs", {k: (v.get("value"), v.get("ms_per_step")) if isinstance(v, dict) else v for k, v in (d.get("configs") or {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
for wl in venice sphere2500; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/c42_launches_${wl}.csv \
      python tests/prof_run.py $wl 3 > $out/c42_ncu_${wl}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_linearize_packets -c 1 -o $out/c42_ncu_linearize python tests/prof_run.py venice 1 > $out/c42_ncu.log 2>&1
ncu -i $out/c42_ncu_linearize.ncu-rep --page raw --csv > $out/c42_ncu_linearize_raw.csv 2>/dev/null
wc -l $out/c42_launches_*.csv
