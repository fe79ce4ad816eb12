#!/bin/bash
# round 2, session 3, final evidence: full GPU suite, smoke, default bench line (extras + widening + CPU leg), ncu launch lists
out=gpurun_out/r2b
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $out/c47_pytest_gpu.txt
tail -3 $out/c47_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/c47_smoke.txt 2>&1; tail -6 $out/c47_smoke.txt
timeout 900 python bench.py > $out/c47_bench_default.json 2> $out/c47_bench_default.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b/c47_bench_default.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
    print("hbm", {k: (round(v["avg_launch_ms"], 4), round(v["frac"], 3)) for k, v in d["roofline"]["hbm_kernels"].items()})
    w = d.get("widening") or {}
    print("marginals", w.get("marginals_sphere2500"))
    print("pcg", json.dumps(w.get("venice_pcg_vs_cholesky"))[:900])
    print("slam", {k: (round(v["lm_iterations_per_s"], 1), v["chi2_rel"]) for k, v in w.items() if k.startswith("slam")})
    print("configs", {k: (v.get("value"), v.get("ms_per_step")) if isinstance(v, dict) else v for k, v in (d.get("configs") or {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
for wl in venice sphere2500; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/c47_launches_${wl}.csv \
      python tests/prof_run.py $wl 3 > $out/c47_ncu_${wl}.log 2>&1
done
wc -l $out/c47_launches_*.csv
